"""The drop-in call at config 4b: pyiga_b200.assemble.stiffness(kvs, geo) for 3D p=4 n=192 on one GPU
(5.3e9 nonzeros: int64 CSR indices, 42 GB of values, 85 GB of host arrays).  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pyiga_b200 import assemble, bspline, geometry


def main():
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    form = sys.argv[3] if len(sys.argv) > 3 else 'stiffness'
    mem = {ln.split(':')[0]: ln.split()[1] for ln in open('/proc/meminfo') if ln.startswith(('MemTotal', 'MemAvailable'))}
    kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
    geo = geometry.twisted_nurbs_box()
    t0 = time.perf_counter()
    A = getattr(assemble, form)(kvs, geo)
    dt = time.perf_counter() - t0
    ones = np.ones(A.shape[1])
    out = {'call': 'pyiga_b200.assemble.%s(kvs, geo)' % form, 'p': p, 'n': n, 'nnz': int(A.nnz), 'shape': list(A.shape),
           'index_dtype': str(A.indices.dtype), 'seconds': dt, 'nnz_per_s': A.nnz / dt,
           'host_kB': mem}
    if form == 'stiffness':     # constants are in the kernel of the stiffness form: row sums vanish
        out['max_abs_row_sum_over_max_abs_entry'] = float(np.abs(A @ ones).max() / np.abs(A.data).max())
    out['sorted_indices'] = bool(A.has_sorted_indices)
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
