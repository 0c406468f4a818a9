"""Slab partition of the assembly across GPUs and the slab-distributed operator.

Assembly shards with no communication: rank r owns a contiguous range of rows of the FIRST tensor
axis, which is a contiguous block ``data[mu0_a:mu0_b, :, :]`` of the multi-level banded tensor
(SURVEY §8e).  Only the matvec of the distributed operator exchanges data (a halo of p planes with
the two neighbouring ranks), see :class:`SlabOperator`.
"""
import numpy as np


def partition_rows(dev, nparts, balance='entries'):
    """Split the rows of axis 0 into `nparts` contiguous slabs.  Returns a list of (row_begin, row_end); slabs may
    be empty only if there are fewer rows than parts.

    balance='entries': equal numbers of band entries (what a matvec or CG on the slabs costs; boundary rows have
    shorter bands).  balance='assembly': equal ASSEMBLY work — a slab pays for every span its rows see in stage 1
    (p more than it owns unless it starts at the boundary) and, in stages 2 + 3, for the band entries it computes:
    for the symmetric forms the upper ones plus the lower ones whose partner row lies in another slab."""
    rs = np.asarray(dev.row_start0(), dtype=np.int64)
    n = len(rs) - 1
    nparts = int(nparts)
    if balance == 'assembly' and 1 < nparts <= n:
        cuts = _partition_by_work(dev, nparts)
        if cuts is not None:
            return [(cuts[i], cuts[i + 1]) for i in range(nparts)]
    cuts = [0]
    for r in range(1, nparts):
        target = rs[-1] * r / nparts
        c = int(np.searchsorted(rs, target, side='left'))
        if c > 0 and abs(rs[c - 1] - target) <= abs(rs[min(c, n)] - target):
            c -= 1
        cuts.append(min(max(c, cuts[-1] + 1 if cuts[-1] + 1 <= n else n), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nparts) if cuts[i + 1] > cuts[i]]


def _partition_by_work(dev, nparts):
    """contiguous partition of the axis-0 rows minimising the largest slab cost
        cost(ra, rb) = ALPHA * spans(ra, rb) + entries(ra, rb)
    with spans = mesh spans in the supports of rows [ra, rb) and entries = band entries of axis 0 the slab computes.
    ALPHA = 0.8 (p + 1): ratio of the measured stage-1 time per span to the stage-2+3 time per band entry of the
    fused kernels (B200, p = 3: 28 us per span, 8.9 us per entry at n = 128; mass 12 / 3.8 us)."""
    from . import _lib
    kv = dev.kvs[1][0]
    supp = np.asarray(kv.mesh_support_idx_all(), dtype=np.int64)
    bidx = np.asarray(dev.structure.bidx[0], dtype=np.int64)
    I, J = bidx[:, 0], bidx[:, 1]
    n = supp.shape[0]
    symmetric = dev.same_space and dev.form in (_lib.FORM_MASS, _lib.FORM_STIFFNESS)
    alpha = 0.8 * (kv.p + 1)
    key = (kv.kv.tobytes(), int(kv.p), int(nparts), bool(symmetric))
    if key in _work_partitions:
        return list(_work_partitions[key])
    rs = np.searchsorted(I, np.arange(n + 1))
    upper = np.concatenate(([0], np.cumsum(np.bincount(I[J >= I], minlength=n)))) if symmetric else rs
    # plain Python numbers: the search below evaluates the cost several thousand times
    s_lo, s_hi = supp[:, 0].tolist(), supp[:, 1].tolist()
    upper = [int(u) for u in upper]
    if symmetric:
        # lower[ra][k]: band entries of the rows [ra, ra + k) whose column lies below ra (k <= p + 1: rows further
        # down do not reach below ra) — the lower entries a slab starting at ra computes itself because their
        # partner row belongs to another slab
        lower = []
        for ra in range(n):
            acc, row = 0, [0]
            for k in range(1, kv.p + 2):
                if ra + k <= n:
                    acc += int(np.count_nonzero(J[rs[ra + k - 1]:rs[ra + k]] < ra))
                row.append(acc)
            lower.append(row)
    pmax = kv.p + 1

    def cost(ra, rb):
        ent = upper[rb] - upper[ra]
        if symmetric:
            ent += lower[ra][min(rb - ra, pmax)]
        return alpha * (s_hi[rb - 1] - s_lo[ra]) + ent

    def cuts_for(limit):
        cuts, ra = [0], 0
        for _ in range(nparts):
            rb = ra + 1
            if rb > n or cost(ra, rb) > limit:
                return None
            while rb < n and cost(ra, rb + 1) <= limit:
                rb += 1
            cuts.append(rb)
            ra = rb
            if ra == n:
                break
        return cuts if cuts[-1] == n else None

    lo, hi = 0.0, float(cost(0, n))
    best = None
    for _ in range(48):
        mid = 0.5 * (lo + hi)
        c = cuts_for(mid)
        if c is None:
            lo = mid
        else:
            best, hi = c, mid
    if best is None or len(best) - 1 > nparts:
        return None
    # the greedy fill may need fewer slabs than ranks: split the largest ones until every rank has rows
    while len(best) - 1 < nparts:
        k = max(range(len(best) - 1), key=lambda i: best[i + 1] - best[i])
        if best[k + 1] - best[k] < 2:
            return None
        best.insert(k + 1, (best[k] + best[k + 1]) // 2)
    if len(_work_partitions) > 64:
        _work_partitions.clear()
    _work_partitions[key] = tuple(best)
    return best


_work_partitions = {}       # (knots, degree, parts, symmetric) -> cuts: every assembler of a rank asks again


class SlabAssembly:
    """One rank's share of a slab-sharded assembly (public multi-GPU entry point).

    Every rank builds the same (replicated, KB-sized) tables, evaluates the coefficient fields only
    on the Gauss planes its rows see, and assembles its row slab independently — no collective.
    """

    def __init__(self, kvs, geo, form, rank=0, world=1, nqp=None, balance='assembly'):
        """`balance`: 'assembly' (equal assembly work per rank) or 'entries' (equal band entries: the choice when
        the slabs are then used by the distributed matvec / CG), see :func:`partition_rows`"""
        from . import _lib, assemblers
        n0 = tuple(kvs)[0].numdofs
        if world > n0:
            raise ValueError('cannot shard %d rows of the first tensor axis over %d ranks' % (n0, world))
        self.form = {'mass': _lib.FORM_MASS, 'stiffness': _lib.FORM_STIFFNESS}[form]
        self.geo = geo
        self.dev = assemblers.DeviceAssembler(tuple(kvs), None, self.form, nqp=nqp)
        slabs = partition_rows(self.dev, world, balance=balance)
        self.slabs = slabs
        self.rows = slabs[rank] if rank < len(slabs) else None
        self.rank, self.world = rank, world

    @property
    def local_nnz(self):
        return 0 if self.rows is None else self.dev.slab_size(self.rows)

    def assemble_mlb(self, out=None, workspace=None, tabulate=False):
        """K1 (optional) + K2 on the slab's planes + K3 on the slab; returns the device buffer."""
        if self.rows is None:
            return None
        if tabulate:
            self.dev.tabulate()
        self.dev.compute_fields(self.geo, rows=self.rows)
        return self.dev.assemble_mlb(rows=self.rows, out=out, workspace=workspace)

    def assemble_csr_device(self, mlb=None):
        """Device CSR arrays (indptr, indices, values) of the local rows; indptr is slab-relative."""
        if self.rows is None:
            return None
        if mlb is None:
            mlb = self.assemble_mlb()
        return self.dev.device_structure.csr_arrays(mlb, row0=self.rows)

    def assemble_csr(self):
        """scipy CSR with the local rows (shape: local rows x all columns)."""
        if self.rows is None:
            return None
        return self.dev.device_structure.to_csr(self.assemble_mlb(), row0=self.rows)

    def csr_sizes(self):
        """(local rows, local nnz, index dtype) of the CSR arrays of this slab"""
        from ._hostcsr import csr_sizes
        nrows, nnz, _ = csr_sizes(self.dev, self.rows)
        return nrows, nnz, (np.int32 if nnz < 2 ** 31 else np.int64)

    def assemble_csr_host(self, host=None, nchunks=8, workspace=None, pattern='host', pattern_threads=None):
        """Assemble the local rows and deliver the CSR arrays (indptr, indices, data) in host memory,
        overlapping the device->host copy of one row chunk with the assembly of the next (CUDA
        backend only; see :mod:`pyiga_b200._hostcsr`).  `host` may hold three preallocated pinned torch
        tensors; returns them.

        The integer arrays are a closed form of the band tables: with ``pattern='host'`` (default)
        host threads write them straight into `host[0:2]` while the GPU computes and ships the values,
        so only 8 of the 12 B/nnz cross PCIe; ``pattern='device'`` produces and copies them from
        the GPU as well."""
        from ._hostcsr import assemble_csr_host
        self.dev.compute_fields(self.geo, rows=self.rows)
        self.last_timings = {}
        tensors, _ = assemble_csr_host(self.dev, self.rows, host=host, nchunks=nchunks, workspace=workspace, pattern=pattern,
                                       pattern_threads=pattern_threads, timings=self.last_timings)
        return tensors


# ---------------------------------------------------------------------------------------------
# slab-distributed operator: y_r = A_r x with a halo exchange, and CG on top of it
# ---------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


class SlabOperator:
    """The rows [ra, rb) of the first tensor axis of a multi-level banded matrix, applied to a vector
    that is distributed by the same slabs.  The band of the slab's rows reaches
    `p` planes into the neighbouring slabs: those planes of x are received from their owners with
    point-to-point messages (NCCL send/recv on NVLink) while the rows that need no halo are already
    being multiplied.  With one rank it is a plain device matvec (``pyiga/mlmatrix_cy.pyx:295-325``)."""

    def __init__(self, dev, mlb, rows=None, slabs=None, rank=0):
        self.dev, self.mlb = dev, mlb
        self.be = dev.be
        n0 = dev.ndofs_test[0]
        self.slabs = slabs if slabs is not None else [(0, n0)]
        self.rank = rank
        self.rows = rows if rows is not None else self.slabs[rank]
        ra, rb = self.rows
        b0 = dev.structure.bidx[0].astype(np.int64)
        sel = (b0[:, 0] >= ra) & (b0[:, 0] < rb)
        self.ha, self.hb = int(b0[sel, 1].min()), int(b0[sel, 1].max()) + 1     # trial planes touched
        self.plane = int(np.prod(dev.ndofs_trial[1:], dtype=np.int64))
        self.plane_rows = int(np.prod(dev.ndofs_test[1:], dtype=np.int64))
        # rows whose band stays inside the slab can be multiplied before the halo arrives
        rmin = np.full(n0, n0, dtype=np.int64)
        rmax = np.zeros(n0, dtype=np.int64)
        np.minimum.at(rmin, b0[:, 0], b0[:, 1])
        np.maximum.at(rmax, b0[:, 0], b0[:, 1])
        inner = [i for i in range(ra, rb) if rmin[i] >= ra and rmax[i] < rb]
        self.inner = (inner[0], inner[-1] + 1) if inner else None
        # message plan: (peer, first plane, last plane) to receive / to send
        self.recv, self.send = [], []
        for s, (sa, sb) in enumerate(self.slabs):
            if s == rank:
                continue
            lo, hi = max(sa, self.ha), min(sb, self.hb)
            if lo < hi:
                self.recv.append((s, lo, hi))
            sel_s = (b0[:, 0] >= sa) & (b0[:, 0] < sb)
            if sel_s.any():
                pa, pb = int(b0[sel_s, 1].min()), int(b0[sel_s, 1].max()) + 1
                lo, hi = max(ra, pa), min(rb, pb)
                if lo < hi:
                    self.send.append((s, lo, hi))
        self.x_ext = self.be.zeros((self.hb - self.ha) * self.plane)
        self.local_size = (rb - ra) * self.plane_rows
        self.halo_bytes = 8 * self.plane * sum(hi - lo for _, lo, hi in self.recv)

    def _t(self, buf):
        """torch view of a backend buffer (communication goes through torch.distributed)"""
        import torch
        return buf if isinstance(buf, torch.Tensor) else torch.from_numpy(np.asarray(buf))

    def matvec(self, x_local, y_local=None):
        be, dev = self.be, self.dev
        ra, rb = self.rows
        if y_local is None:
            y_local = be.empty(self.local_size)
        xe = self._t(self.x_ext)
        xl = self._t(x_local)
        xe[(ra - self.ha) * self.plane:(rb - self.ha) * self.plane].copy_(xl)
        dist = _dist()
        works = []
        if dist is not None and (self.recv or self.send):
            ops = []
            for peer, lo, hi in self.recv:
                ops.append(dist.P2POp(dist.irecv, xe[(lo - self.ha) * self.plane:(hi - self.ha) * self.plane], peer))
            for peer, lo, hi in self.send:
                ops.append(dist.P2POp(dist.isend, xl[(lo - ra) * self.plane:(hi - ra) * self.plane], peer))
            works = dist.batch_isend_irecv(ops)
        ds = dev.device_structure
        yt = self._t(y_local)
        if self.inner and works:
            ia, ib = self.inner
            ds.matvec_device(self._slab_ptr(ia), self.x_ext, yt[(ia - ra) * self.plane_rows:(ib - ra) * self.plane_rows],
                             row0=(ia, ib), x_j0=self.ha)
            for w in works:
                w.wait()
            for (a, b) in ((ra, ia), (ib, rb)):
                if a < b:
                    ds.matvec_device(self._slab_ptr(a), self.x_ext, yt[(a - ra) * self.plane_rows:(b - ra) * self.plane_rows],
                                     row0=(a, b), x_j0=self.ha)
        else:
            for w in works:
                w.wait()
            ds.matvec_device(self.mlb, self.x_ext, y_local, row0=(ra, rb), x_j0=self.ha)
        return y_local

    def _slab_ptr(self, row):
        """view of the MLB slab starting at row `row` of axis 0"""
        rs = self.dev.row_start0()
        inner = int(np.prod(self.dev.nband[1:], dtype=np.int64))
        off = int(rs[row] - rs[self.rows[0]]) * inner
        return self._t(self.mlb)[off:]


def _allsum(t):
    dist = _dist()
    if dist is not None:
        dist.all_reduce(t)
    return t


def cg(op, b, M=None, x0=None, rtol=1e-10, maxiter=200):
    """Preconditioned conjugate gradients on slab-distributed torch vectors (same recurrences as
    scipy.sparse.linalg.cg, which the reference uses, ``pyiga/approx.py:92-93``).  `op(x)` and
    `M(r)` map local slabs to local slabs; dot products are all-reduced.  Returns
    (x, iterations, [relative residual norms])."""
    import torch
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b - op(x) if x0 is not None else b.clone()
    bnorm = float(torch.sqrt(_allsum(torch.dot(b, b).reshape(1)))[0])
    if bnorm == 0.0:
        return x, 0, [0.0]
    z = M(r) if M is not None else r
    p = z.clone()
    rz = _allsum(torch.dot(r, z).reshape(1))
    hist = []
    it = 0
    for it in range(1, maxiter + 1):
        Ap = op(p)
        alpha = rz / _allsum(torch.dot(p, Ap).reshape(1))
        x += alpha * p
        r -= alpha * Ap
        res = float(torch.sqrt(_allsum(torch.dot(r, r).reshape(1)))[0]) / bnorm
        hist.append(res)
        if res <= rtol:
            break
        z = M(r) if M is not None else r
        rz_new = _allsum(torch.dot(r, z).reshape(1))
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, hist


class GatheredKronecker:
    """Kronecker-product preconditioner on a slab-distributed vector: the vector is a few MB, so it
    is all-gathered, the mode products run redundantly on every rank, and the local slab is kept."""

    def __init__(self, kron, slabs, rank, plane):
        self.kron, self.slabs, self.rank, self.plane = kron, slabs, rank, plane

    def __call__(self, r_local):
        import torch
        dist = _dist()
        if dist is None or len(self.slabs) == 1:
            full = r_local
        else:
            sizes = [(b - a) * self.plane for a, b in self.slabs]
            parts = [torch.empty(s, dtype=r_local.dtype, device=r_local.device) for s in sizes]
            dist.all_gather(parts, r_local) if len(set(sizes)) == 1 else self._gather_uneven(parts, r_local)
            full = torch.cat(parts)
        out = self.kron.matvec_device(full)
        out = out if isinstance(out, torch.Tensor) else torch.from_numpy(np.asarray(out))
        a, b = self.slabs[self.rank]
        return out[a * self.plane:b * self.plane].clone()

    def _gather_uneven(self, parts, r_local):
        import torch
        dist = _dist()
        n = max(p.numel() for p in parts)
        pad = torch.zeros(n, dtype=r_local.dtype, device=r_local.device)
        pad[:r_local.numel()] = r_local
        bufs = [torch.empty_like(pad) for _ in parts]
        dist.all_gather(bufs, pad)
        for p, b in zip(parts, bufs):
            p.copy_(b[:p.numel()])
