"""Where the end-to-end time goes: constructor, fields, chunk pipeline, host pattern threads.
    python tools/e2e_breakdown.py [n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyiga_b200 import bspline, geometry
from pyiga_b200.dist import SlabAssembly


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    out = {'n': n, 'cpus': os.cpu_count()}
    sl = SlabAssembly(kvs, geo, 'stiffness')
    nrows, nnz, idt = sl.csr_sizes()
    tdt = torch.int32 if idt == np.int32 else torch.int64
    pinned = [torch.empty(nrows + 1, dtype=tdt, pin_memory=True), torch.empty(nnz, dtype=tdt, pin_memory=True),
              torch.empty(nnz, dtype=torch.float64, pin_memory=True)]
    for t in pinned:
        t.zero_()
    ds = sl.dev.device_structure
    for thr in (1, 4, 8, 15):
        if thr > (os.cpu_count() or 1):
            break
        t0 = time.perf_counter()
        ds.csr_pattern_host(pinned[0], pinned[1], nthreads=thr)
        dt = time.perf_counter() - t0
        out['pattern_host_%dthr_ms' % thr] = 1e3 * dt
    ws = None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sl = SlabAssembly(kvs, geo, 'stiffness')
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sl.assemble_csr_host(host=pinned, pattern='host')
        t2 = time.perf_counter()
        sl.assemble_csr_host(host=pinned, pattern='device')
        t3 = time.perf_counter()
    out['phases_default_threads'] = dict(sl.last_timings) if getattr(sl, 'last_timings', None) else None
    sl.assemble_csr_host(host=pinned, pattern='host')
    out['phases_default_threads'] = dict(sl.last_timings)
    for thr in (1, 4, 15):
        for nch in (8,):
            ts = []
            for rep in range(3):
                torch.cuda.synchronize()
                t4 = time.perf_counter()
                sl.assemble_csr_host(host=pinned, pattern='host', pattern_threads=thr, nchunks=nch)
                ts.append(time.perf_counter() - t4)
            out['csr_host_%dthr_%dchunks_ms' % (thr, nch)] = 1e3 * min(ts)
    out['constructor_ms'] = 1e3 * (t1 - t0)
    out['assemble_csr_host_pattern_host_ms'] = 1e3 * (t2 - t1)
    out['assemble_csr_host_pattern_device_ms'] = 1e3 * (t3 - t2)
    # raw D2H of the values alone
    v = torch.empty(nnz, dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pinned[2].copy_(v, non_blocking=True)
    torch.cuda.synchronize()
    out['d2h_values_only_ms'] = 1e3 * (time.perf_counter() - t0)
    out['d2h_GBps'] = 8 * nnz / (out['d2h_values_only_ms'] * 1e6)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
