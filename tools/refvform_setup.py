"""Set-up time of a REFERENCE VForm object on the device backend: where the coefficient arrays of the
finalized form are evaluated (pyiga_b200/refvform.py).  Builds pyiga's own conv-diff form (parse_vf of the
reference, needs oracle/_ref importable), instantiates the device assembler at p=3 with n spans per axis,
once with the interpreter on device arrays (default on CUDA) and once forced to the host numpy path, and
assembles the MLB tensor.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
import numpy as np
import torch

from pyiga_b200 import _device, bspline, geometry, refvform, vform


def run(n, host):
    from pyiga import bspline as rbs, vform as rvf
    kvs = 3 * (rbs.make_knots(3, 0.0, 1.0, n),)
    geo = geometry.twisted_box()
    dc = lambda x, y, z: 1.0 + np.sin(x) * y
    rf = rvf.parse_vf('(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx', kvs,
                      args={'diff_coeff': dc})
    cls = vform.compile_vform(rf)
    saved = refvform._tensor_device
    if host:
        refvform._tensor_device = lambda: None
    try:
        out = {}
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            asm = cls(kvs, geo=geo, diff_coeff=dc)
            torch.cuda.synchronize()
            out['setup_ms_call%d' % it] = 1e3 * (time.perf_counter() - t0)
            t0 = time.perf_counter()
            M = asm.dev.assemble_mlb()
            torch.cuda.synchronize()
            out['assemble_ms_call%d' % it] = 1e3 * (time.perf_counter() - t0)
            out['abs_sum'] = float(M.abs().sum())
            del asm, M
        return out
    finally:
        refvform._tensor_device = saved


def main():
    _device.backend()
    res = {'form': "reference parse_vf: (inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx, p=3",
           'runs': []}
    for n, host in ((48, True), (48, False), (96, False)):
        r = run(n, host)
        r.update(n=n, interpreter='host numpy' if host else 'device arrays')
        res['runs'].append(r)
    print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
