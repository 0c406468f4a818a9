"""First-contact script for a GPU box: environment, FP64 probe, stage timings of the pipeline."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import _device, assemblers, bspline, geometry


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:
        return 'ERR %s' % e


def main():
    out = {}
    out['cpu'] = sh("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Flags' | head -4 | cut -c1-400")
    out['gpu'] = sh('nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv')
    be = _device.backend()
    g = C.c_double()
    _device.check(be.lib.pb200_probe_fp64(0, 200000, C.byref(g)))
    out['fp64_gflops'] = g.value
    print(json.dumps(out, indent=1), flush=True)

    cases = [(3, 32), (3, 64), (3, 128), (2, 64), (4, 48)]
    if len(sys.argv) > 1:
        cases = [tuple(int(t) for t in a.split(',')) for a in sys.argv[1:]]
    for form in ('Stiffness', 'Mass'):
        for p, n in cases:
            kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
            geo = geometry.twisted_nurbs_box()
            t0 = time.time()
            asm = getattr(assemblers, form + 'Assembler3D')(kvs, geo)
            dev = asm.dev
            be.synchronize()
            t_setup = time.time() - t0
            ws_bytes = dev.workspace_bytes()
            ws = be.empty(ws_bytes, np.uint8)
            outbuf = be.empty(dev.nnz)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            dev.set_timing(True)
            res = []
            for it in range(4):
                e[0].record()
                dev.tabulate()
                e[1].record()
                dev.compute_fields(geo)
                e[2].record()
                dev.assemble_mlb(out=outbuf, workspace=ws)
                e[3].record()
                torch.cuda.synchronize()
                res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
            st = dev.stage_times()
            k1, k2, k3 = res[-1]
            tot = k1 + k2 + k3
            print(json.dumps({'form': form, 'p': p, 'n': n, 'nnz': dev.nnz, 'setup_s': round(t_setup, 3),
                              'ws_GB': round(ws_bytes / 1e9, 2), 'k1_ms': round(k1, 3), 'k2_ms': round(k2, 3),
                              'k3_ms': round(k3, 3), 'nnz_per_s': dev.nnz / (tot * 1e-3),
                              'stages': [(a, round(b, 3)) for a, b in st]}), flush=True)
            del ws, outbuf, asm, dev
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
