"""Stage times of the conv-diff vform (config 5, p=3 n=96) on one GPU: python tools/convdiff_stages.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import _device, assemble, bspline, geometry


def main():
    be = _device.backend()
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 96),)
    geo = geometry.twisted_box()
    form = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'
    asm = assemble.instantiate_assembler(form, kvs, {'geo': geo, 'diff_coeff': lambda x, y, z: 1.0 + x * y}, None)
    dev = asm.dev
    res = be.empty(dev.nnz)
    ws = be.empty(dev.workspace_bytes(), np.uint8)
    for _ in range(3):
        dev.assemble_mlb(out=res, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dev.assemble_mlb(out=res, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    dev.set_timing(True)
    dev.assemble_mlb(out=res, workspace=ws)
    torch.cuda.synchronize()
    print('%.4f ms/step' % (e0.elapsed_time(e1) / 10), [(k, round(v, 4)) for k, v in dev.stage_times()])


if __name__ == '__main__':
    main()
