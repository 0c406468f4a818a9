"""A/B timing of an assembler option on one configuration:
    python tools/ab_option.py <option> [form] [p] [n]
prints the stage times of a few steps with the option set to 1 and to 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import _device, bspline, geometry
from pyiga_b200.dist import SlabAssembly


def main():
    opt = sys.argv[1]
    form = sys.argv[2] if len(sys.argv) > 2 else 'stiffness'
    p = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 128
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    sa = SlabAssembly(kvs, geo, form)
    out = be.empty(sa.local_nnz)
    ws = be.empty(sa.dev.workspace_bytes(sa.rows), np.uint8)
    ref = None
    for val in (1, 0, 1, 0):
        sa.dev.set_option(opt, val)
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
        sa.dev.set_timing(False)
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        print('%s=%d  %.4f ms/step  %s  diff vs first %.2e' % (opt, val, e0.elapsed_time(e1) / 10, {k: round(v, 4) for k, v in st.items()}, err))


if __name__ == '__main__':
    main()
