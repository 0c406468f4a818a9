"""Where the set-up time of a generic form goes (conv-diff p=3 n=96): cProfile of instantiate_assembler on the GPU.
    python tools/setup_profile.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pyiga_b200 import assemble, bspline, geometry


def main():
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 96),)
    geo = geometry.twisted_box()
    form = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'
    args = {'geo': geo, 'diff_coeff': lambda x, y, z: 1.0 + x * y}
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        asm = assemble.instantiate_assembler(form, kvs, args, None)
        torch.cuda.synchronize()
        print('setup %d: %.1f ms' % (it, 1e3 * (time.perf_counter() - t0)))
        del asm
    pr = cProfile.Profile()
    pr.enable()
    asm = assemble.instantiate_assembler(form, kvs, args, None)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)


if __name__ == '__main__':
    main()
