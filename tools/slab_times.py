"""Per-rank step times of an N-way slab partition, measured one slab after the other on ONE GPU:
    python tools/slab_times.py [form] [p] [n] [world]
The slabs of the multi-GPU run are independent (no collective), so the slowest of them is the multi-GPU step
time up to launch overheads; this shows where the scaling loss sits without N GPUs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import _device, bspline, geometry
from pyiga_b200.dist import SlabAssembly


def main():
    form = sys.argv[1] if len(sys.argv) > 1 else 'stiffness'
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    world = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    res = []
    for rank in range(world):
        sa = SlabAssembly(kvs, geo, form, rank=rank, world=world)
        out = be.empty(sa.local_nnz)
        ws = be.empty(sa.dev.workspace_bytes(sa.rows), np.uint8)
        for _ in range(3):
            sa.assemble_mlb(out=out, workspace=ws, tabulate=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sa.assemble_mlb(out=out, workspace=ws, tabulate=True)
        e1.record()
        torch.cuda.synchronize()
        sa.dev.set_timing(True)
        sa.assemble_mlb(out=out, workspace=ws, tabulate=True)
        torch.cuda.synchronize()
        st = dict(sa.dev.stage_times())
        res.append({'rank': rank, 'rows': list(sa.rows), 'nnz': sa.local_nnz, 'ms': e0.elapsed_time(e1) / 10,
                    'stages': {k: round(v, 4) for k, v in st.items()}})
        del sa, out, ws
    total = sum(r['nnz'] for r in res)
    worst = max(r['ms'] for r in res)
    print(json.dumps({'form': form, 'p': p, 'n': n, 'world': world, 'max_ms': worst, 'sum_ms': sum(r['ms'] for r in res),
                      'nnz_per_s_at_max': total / (worst * 1e-3), 'ranks': res}))


if __name__ == '__main__':
    main()
