"""Slab-distributed, device-resident preconditioned CG (``csrc/distcg.cuh``).

The counterpart of ``scipy.sparse.linalg.cg(M, b, M=KroneckerOperator(*Minvs))`` as the reference uses
it in ``project_L2`` (``pyiga/approx.py:82-96``) for a multi-level banded matrix whose rows of the
first tensor axis are sharded over the GPUs of one node.  ``torch.distributed`` is used once, to pass
the 64-byte window handles around; the iterations themselves exchange halo planes, dot products and
the preconditioner's gathered input through peer-mapped device memory inside the kernels.
"""
import ctypes as C

import numpy as np

from . import _device


def band_halo(structure):
    """number of planes the band of the first axis reaches into a neighbouring row slab"""
    b0 = structure.bidx[0].astype(np.int64)
    return int(np.abs(b0[:, 0] - b0[:, 1]).max())


class DistributedCG:
    """One rank's handle on the solver.

    dstruct: :class:`~pyiga_b200._mlb.DeviceStructure` of the whole matrix; mlb: device buffer with the
    value tensor of the local row slab; slabs: list of (row_begin, row_end) per rank; Ainv: the three
    dense inverse factors of the Kronecker preconditioner (numpy, N_k x N_k)."""

    def __init__(self, dstruct, mlb, slabs, rank, Ainv, group=None):
        self.be = be = _device.backend()
        self.dstruct, self.mlb = dstruct, mlb
        S = dstruct.structure
        world = len(slabs)
        assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])), 'slabs must be contiguous'
        self.rank, self.world, self.slabs = rank, world, slabs
        cuts = (C.c_int * (world + 1))(*([s[0] for s in slabs] + [slabs[-1][1]]))
        halo = band_halo(S)
        self.plane = int(np.prod([b[1] for b in S.bs[1:]], dtype=np.int64))
        self.nloc = (slabs[rank][1] - slabs[rank][0]) * self.plane
        self.comm = None
        if world > 1:
            import torch
            import torch.distributed as dist
            nbytes = C.c_size_t()
            _device.check(be.lib.pb200_cg_window_bytes(dstruct.handle, world, cuts, halo, C.byref(nbytes)))
            comm = C.c_void_p()
            _device.check(be.lib.pb200_comm_create(be.device_index, rank, world, nbytes.value, C.byref(comm)))
            self.comm = comm
            mine = np.zeros(64, dtype=np.uint8)
            _device.check(be.lib.pb200_comm_handle(comm, mine.ctypes.data))
            on_gpu = dist.get_backend(group) == 'nccl'
            t = torch.from_numpy(mine)
            t = t.cuda() if on_gpu else t
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t, group=group)
            allh = np.concatenate([p.cpu().numpy() for p in parts])
            _device.check(be.lib.pb200_comm_open_peers(comm, allh.ctypes.data))
            dist.barrier(group=group)
        self._Ainv = [be.from_host(np.ascontiguousarray(A, dtype=np.float64).ravel()) for A in Ainv]
        ptrs = (C.c_void_p * 3)(*[be.ptr(a) for a in self._Ainv])
        h = C.c_void_p()
        _device.check(be.lib.pb200_cg_create(dstruct.handle, rank, world, cuts, halo, be.ptr(mlb), ptrs, self.comm, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, 'handle', None) is not None:
            self.be.lib.pb200_cg_destroy(self.handle)
            self.handle = None
        if getattr(self, 'comm', None) is not None:
            self.be.lib.pb200_comm_destroy(self.comm)
            self.comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, b_local, rtol=1e-10, maxiter=200, check_every=10, out=None):
        """(x_local, iterations, relative residual); `b_local`: device buffer with the local slab of b"""
        be = self.be
        x = be.empty(self.nloc) if out is None else out
        it, res = C.c_int(), C.c_double()
        _device.check(be.lib.pb200_cg_solve(self.handle, be.ptr(b_local), be.ptr(x), float(rtol), int(maxiter), int(check_every),
                                            C.byref(it), C.byref(res), be.stream()))
        return x, it.value, res.value

    def matvec(self, p_local, out=None):
        be = self.be
        y = be.empty(self.nloc) if out is None else out
        _device.check(be.lib.pb200_cg_matvec(self.handle, be.ptr(p_local), be.ptr(y), be.stream()))
        return y
