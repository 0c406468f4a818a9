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
