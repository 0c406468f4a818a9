"""Runs a few assembly steps of one configuration (for ncu captures):
    python tools/step_once.py [form] [p] [n] [steps]"""
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
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    sa = SlabAssembly(kvs, geo, form)
    out = be.empty(sa.local_nnz)
    ws = be.empty(sa.dev.workspace_bytes(sa.rows), np.uint8)
    for _ in range(steps):
        sa.assemble_mlb(out=out, workspace=ws, tabulate=True)
    torch.cuda.synchronize()
    print('done', float(out[:1000].abs().max()))


if __name__ == '__main__':
    main()
