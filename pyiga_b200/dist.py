"""Slab partition of the assembly across GPUs and the slab-distributed operator.

Assembly shards with no communication: rank r owns a contiguous range of rows of the FIRST tensor
axis, which is a contiguous block ``data[mu0_a:mu0_b, :, :]`` of the multi-level banded tensor
(SURVEY §8e).  Only the matvec of the distributed operator exchanges data (a halo of p planes with
the two neighbouring ranks), see :class:`SlabOperator`.
"""
import numpy as np


def partition_rows(dev, nparts):
    """Split the rows of axis 0 into `nparts` contiguous slabs with balanced band counts
    (boundary rows have shorter bands).  Returns a list of (row_begin, row_end); slabs may be
    empty only if there are fewer rows than parts."""
    rs = np.asarray(dev.row_start0(), dtype=np.int64)
    n = len(rs) - 1
    nparts = int(nparts)
    cuts = [0]
    for r in range(1, nparts):
        target = rs[-1] * r / nparts
        c = int(np.searchsorted(rs, target, side='left'))
        if c > 0 and abs(rs[c - 1] - target) <= abs(rs[min(c, n)] - target):
            c -= 1
        cuts.append(min(max(c, cuts[-1] + 1 if cuts[-1] + 1 <= n else n), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nparts) if cuts[i + 1] > cuts[i]]


class SlabAssembly:
    """One rank's share of a slab-sharded assembly (public multi-GPU entry point).

    Every rank builds the same (replicated, KB-sized) tables, evaluates the coefficient fields only
    on the Gauss planes its rows see, and assembles its row slab independently — no collective.
    """

    def __init__(self, kvs, geo, form, rank=0, world=1, nqp=None):
        from . import _lib, assemblers
        self.form = {'mass': _lib.FORM_MASS, 'stiffness': _lib.FORM_STIFFNESS}[form]
        self.geo = geo
        self.dev = assemblers.DeviceAssembler(tuple(kvs), None, self.form, nqp=nqp)
        slabs = partition_rows(self.dev, world)
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
