#!/usr/bin/env python
"""Benchmark of the assembly hot path: assembled nonzeros/sec, 3D p=3 stiffness.

    python bench.py --gpus N --steps K --warmup W            (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference ...                      (the reference's CPU path, rank 0 only)

A step is one pass of the hot path over the whole workload: K1 (1D basis tables) + K2 (geometry
Jacobian and coefficient fields) + K3 (sum-factorised contraction into the multi-level banded value
tensor).  With N GPUs the rows of the first tensor axis are slab-sharded (no collective) and the
total work is fixed ("strong" scaling).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'assembled nonzeros/sec (3D p=3 stiffness)'
UNIT = 'nnz/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--p', type=int, default=3)
    ap.add_argument('--n', type=int, default=128)
    ap.add_argument('--form', default='stiffness', choices=['stiffness', 'mass'])
    ap.add_argument('--geo', default='nurbs', choices=['nurbs', 'bspline'])
    ap.add_argument('--ref-n', type=int, default=64, help='spans per axis of the bounded CPU sample (the reference arm)')
    ap.add_argument('--ref-quick', action='store_true', help='reference arm: only the --ref-n size (no n=32/48, no single thread)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary configurations (extra_configs)')
    ap.add_argument('--no-cg', action='store_true', help='skip the distributed CG check on the mass matrix (config 5, `cg` key)')
    return ap.parse_args()


def workload_name(a, n=None):
    geo = 'twisted NURBS box' if a.geo == 'nurbs' else 'twisted_box B-spline'
    return '3D %s p=%d n=%d, %s, nqp=%d' % (a.form, a.p, a.n if n is None else n, geo, a.p + 1)


def algorithmic_flops(p, n, form):
    """F of SURVEY.md §8d: row-wise sum factorisation with 1D support sparsity, full band."""
    q = p + 1
    N = n + p
    S1, P1, M1 = n * (p + 1) * q, n * (p + 1) ** 2 * q, N * (2 * p + 1) - p * (p + 1)
    na, c0 = (3, 18) if form == 'stiffness' else (1, 2)
    return 2.0 * na * (S1 * S1 * P1 + S1 * P1 * M1 + P1 * M1 * M1) + c0 * float(S1) ** 3, M1 ** 3


# ---------------------------------------------------------------------------------------------
# reference arm (CPU)
# ---------------------------------------------------------------------------------------------
def _reference_setup(a, n=None):
    """Returns (assemble_fn, kind, cores): the reference's own implementation from oracle/_ref when it
    loads, else the C port of the oracle; `n` spans per axis."""
    import numpy as np
    n = a.ref_n if n is None else n
    cores = os.cpu_count() or 1
    probe = subprocess.run([sys.executable, '-c',
                            'import sys; sys.path.insert(0, %r); import pyiga.assemblers' % os.path.join(ROOT, 'oracle', '_ref')],
                           capture_output=True)
    if probe.returncode == 0:
        if os.path.join(ROOT, 'oracle', '_ref') not in sys.path:
            sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        import pyiga
        from pyiga import assemble, bspline, geometry
        pyiga.set_max_threads(cores)
        kvs = 3 * (bspline.make_knots(a.p, 0.0, 1.0, n),)
        G = geometry.twisted_box()
        if a.geo == 'nurbs':
            i, j, k = np.meshgrid(np.arange(2), np.arange(4), np.arange(2), indexing='ij')
            G = geometry.NurbsFunc(G.kvs, G.coeffs.copy(), 1.0 + 0.25 * ((i + 2 * j + 3 * k) % 3))
        fn = assemble.stiffness if a.form == 'stiffness' else assemble.mass
        return (lambda: fn(kvs, G)), 'reference', cores
    from oracle import c_oracle, pyiga_oracle as orc
    from pyiga_b200 import geometry
    G = geometry.twisted_nurbs_box() if a.geo == 'nurbs' else geometry.twisted_box()
    kv = orc.make_knots(a.p, 0.0, 1.0, n)

    def run():
        prob = orc.Problem([kv] * 3, [a.p] * 3, [k.kv for k in G.kvs], [k.p for k in G.kvs], G.coeffs, G._rational)
        return c_oracle.assemble_csr(prob, a.form, nthreads=cores)
    return run, 'port', cores


def _time_reference(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    best, nnz = None, 0
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        A = fn()
        times.append(time.perf_counter() - t0)
        nnz = A.nnz
        del A
    return sum(times) / len(times), min(times), nnz


def run_reference(a, quiet=False):
    """The reference's own CPU implementation of the path (pyiga.assemble.stiffness / mass end to end:
    setup + multi_entries + CSR) on a BOUNDED sample of the workload: n = --ref-n spans per axis
    instead of the GPU arm's n (throughput in nnz/s is flat in n; p=3 n=128
    needs > 30 GB and minutes per step).  The line says which size was timed; `same_config` is false."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    steps = max(1, min(a.steps, 3))             # bounded: a step at n=64 takes ~10 s on 16 cores
    warmup = min(a.warmup, 1)
    fn, kind, cores = _reference_setup(a)
    mean_s, best_s, nnz = _time_reference(fn, warmup, steps)
    ms = 1e3 * mean_s
    value = nnz / mean_s
    what = 'pyiga.assemble.%s' % a.form if kind == 'reference' else 'oracle C port'
    sample = ('%s: %d nnz per step, end-to-end %s (setup + multi_entries + CSR) on %d host threads, mean of %d steps'
              % (workload_name(a, a.ref_n), nnz, what, cores, steps))
    base = {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample, 'ref_n': a.ref_n,
            'same_config': a.ref_n == a.n, 'extrapolated': a.ref_n != a.n, 'best_of_steps': nnz / best_s}
    if not a.ref_quick and kind == 'reference':
        # p=3 at n in {32, 48, 64}, and the single-thread figure
        import pyiga
        sizes = {}
        for n in (32, 48):
            if n >= a.ref_n:
                continue
            f2, _, _ = _reference_setup(a, n)
            m, b, z = _time_reference(f2, 0, 1)
            sizes['n=%d' % n] = {'nnz': z, 'nnz_per_s': z / b}
        sizes['n=%d' % a.ref_n] = {'nnz': nnz, 'nnz_per_s': nnz / best_s}
        base['sizes'] = sizes
        f1, _, _ = _reference_setup(a, 32)
        pyiga.set_max_threads(1)
        m, b, z = _time_reference(f1, 0, 1)
        pyiga.set_max_threads(cores)
        base['single_thread'] = {'value': z / b, 'unit': UNIT, 'cores': 1, 'sample': workload_name(a, 32)}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': steps,
        'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(a, a.ref_n), 'sample': sample,
                   'note': 'bounded sample of the GPU arm\'s workload (%s): same degree, geometry and form, n=%d instead of n=%d'
                           % (workload_name(a), a.ref_n, a.n)},
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if not quiet:
        emit(line)
    return line


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', os.environ.get('PB200_BENCH_SMI_MS', '20')], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                t = [x.strip() for x in ln.split(',')]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1]))
                    mx.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), t[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                   'samples': len(sm)}
        return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def pipeline_bytes(dev, rows, form, p, mirror=True):
    """Algorithmic HBM bytes of every pipeline kernel for the row slab `rows` (each input element a
    kernel needs is read once, each output element written once).  With mirroring the
    final stage computes only the upper half of each symmetric pair of lines, and the earlier
    stages only the terms those lines need (slab filter modes 1/2/3 of csrc/walk.cuh)."""
    import numpy as np
    ra, rb = rows
    G, M = dev.nnodes, dev.nband
    S = dev.structure
    b0 = S.bidx[0].astype(np.int64)
    b1 = S.bidx[1].astype(np.int64)
    i0, j0 = b0[:, 0], b0[:, 1]
    ini, inj = (i0 >= ra) & (i0 < rb), (j0 >= ra) & (j0 < rb)
    c1 = int(ini.sum())
    c2 = int((ini | inj).sum())
    c3 = int((ini & ((i0 <= j0) | ~inj)).sum()) if mirror else c1
    upper1 = int((b1[:, 0] <= b1[:, 1]).sum())
    if mirror:
        lines = int((ini & inj & (i0 < j0)).sum()) * M[1] + int((ini & (i0 == j0)).sum()) * upper1 \
            + int((ini & ~inj).sum()) * M[1]
    else:
        lines = c1 * M[1]
    msi = np.asarray(dev.kvs[0][0].mesh_support_idx_all())
    planes = int(msi[rb - 1, 1] - msi[ra, 0]) * (p + 1)
    plane = G[1] * G[2]
    d = 8.0
    out = {'k2_fields': d * dev.nfields * planes * plane}
    # fused stage 1 (geometry + fields in registers): nothing is read, the X1 terms are written
    out['s1f'] = d * (3 * c3 + 3 * c2) * plane
    out['s1f_mass'] = d * c3 * plane
    if form == 'stiffness':
        out['s1a'] = d * (3 * planes * plane + (c3 + 2 * c2) * plane)
        out['s1b'] = d * (3 * planes * plane + (2 * c3 + c2) * plane)
        out['s2a_final4'] = d * (4 * c3 * plane + c3 * M[1] * G[2])
        out['s2b'] = d * ((2 * c2 + c3) * plane + (c2 + c3) * M[1] * G[2])
        out['s3_final4'] = d * (4 * lines * G[2] + c1 * M[1] * M[2])
    else:
        out['s1_copy'] = d * (planes * plane + c3 * plane)
        out['s2_copy'] = d * (c3 * plane + c3 * M[1] * G[2])
        out['s3_copy'] = d * (lines * G[2] + c1 * M[1] * M[2])
    return out


def slab_counts(dev, rows):
    """band entries of axis 0 by slab filter: c1 rows of the slab, c2 rows or columns in the slab, c3 the
    "upper" entries the symmetric pipeline computes (the rest is mirrored)"""
    import numpy as np
    ra, rb = rows
    b0 = dev.structure.bidx[0].astype(np.int64)
    i0, j0 = b0[:, 0], b0[:, 1]
    ini, inj = (i0 >= ra) & (i0 < rb), (j0 >= ra) & (j0 < rb)
    return int(ini.sum()), int((ini | inj).sum()), int((ini & ((i0 <= j0) | ~inj)).sum())


def own_fp64_ops(dev, rows, form, p):
    """FP64 multiply / fused-multiply-add instructions per thread-lane that the pairwise sum
    factorisation of THIS implementation executes for the slab (fused pipeline: stage 1 with symmetric
    windows, stages 2+3 on the upper band entries of axis 0 only, 32-span batches overlapping by p).
    Geometry evaluation is not counted (SURVEY 8d excludes K2 from F as well).  flops = 2 x this."""
    import numpy as np
    P1 = p + 1
    G, M = dev.nnodes, dev.nband
    c1, c2, c3 = slab_counts(dev, rows)
    msi = np.asarray(dev.kvs[0][0].mesh_support_idx_all())
    planes = int(msi[rows[1] - 1, 1] - msi[rows[0], 0]) * P1
    pts = planes * G[1] * G[2]
    sym, non = P1 + P1 * (P1 + 1) // 2, P1 + P1 * P1
    n2 = G[2] // P1
    nbatch = 1 + max(0, -(-(n2 + p - 32) // (32 - p)))
    # lanes a band entry of axis 0 occupies: full batches of 32, and the last batch shares its warp with the
    # tails of other entries when it is narrow enough (fused23.cuh)
    tail_w = min(32, n2 + p - (nbatch - 1) * (32 - p))
    tail_k = min(4, 32 // tail_w) if nbatch >= 2 else 1
    lanes = (nbatch - 1) * 32.0 + 32.0 / max(tail_k, 1)
    overlap = lanes / n2 if n2 >= 32 else 1.0
    if form == 'stiffness':
        s1 = pts * (4 * sym + 2 * non)
        # phase A: terms T0 (4 inputs), T1, T2 (2 inputs each) with full blocks, T3 (1 input) with a symmetric one
        a = c3 * G[1] * G[2] * overlap * (9 * P1 + 4 * P1 * P1 + P1 * (P1 + 1) // 2)
        b = c3 * G[1] * M[2] * (4 * P1 + 2 * P1 * P1)
    else:
        s1 = pts * sym
        a = c3 * G[1] * G[2] * overlap * sym
        b = c3 * G[1] * M[2] * non
    return {'stage1': float(s1), 'stage23_phaseA': float(a), 'stage23_phaseB': float(b)}


def fused_bytes(dev, rows, form):
    """algorithmic HBM bytes of the fused pipeline: X1 written once by stage 1 and read once by stages
    2+3 (terms read at mu0 and at the transposed mu0 are two reads), the matrix written once"""
    G, M = dev.nnodes, dev.nband
    c1, c2, c3 = slab_counts(dev, rows)
    plane = 8.0 * G[1] * G[2]
    if form == 'stiffness':
        return {'s1f': (4 * c3 + 2 * c2) * plane, 's23': 8 * c3 * plane + 8.0 * c1 * M[1] * M[2]}
    return {'s1f_mass': c3 * plane, 's23_mass': c3 * plane + 8.0 * c1 * M[1] * M[2]}


def parity_against_fixture(a, dev, rows, d_out):
    """Compare the matrix that was just timed with the sampled entries the REAL reference computed for
    this workload (tests/golden/large_*.npz, generated by tests/golden/make_golden_large.py with
    pyiga's multi_entries).  Returns (max abs error, entries compared, scale) or None without a fixture."""
    import numpy as np
    name = '%s_p%d_n%d' % ('stiff' if a.form == 'stiffness' else 'mass', a.p, a.n)
    path = os.path.join(ROOT, 'tests', 'golden', 'large_%s.npz' % name)
    if a.geo != 'nurbs' or not os.path.exists(path):
        return None
    z = np.load(path)
    ij = z['ij'].astype(np.int64)
    pos = dev.structure.positions(ij[:, 0], ij[:, 1])
    nout = int(z['nout'])
    if not (np.all(pos[len(pos) - nout:] == -1) and np.all(pos[:len(pos) - nout] >= 0)):
        return (float('inf'), 0, 1.0, name)
    plane_rows = int(np.prod(dev.ndofs_test[1:], dtype=np.int64))
    i0 = ij[:, 0] // plane_rows
    sel = (pos >= 0) & (i0 >= rows[0]) & (i0 < rows[1])
    scale = float(z['full_maxabs']) if 'full_maxabs' in z else float(z['sample_maxabs'])
    if not sel.any():
        return (0.0, 0, scale, name)
    inner = int(np.prod(dev.nband[1:], dtype=np.int64))
    local = pos[sel] - int(dev.row_start0()[rows[0]]) * inner
    got = dev.be.to_host(d_out[dev.be.from_host(local)])
    return (float(np.abs(got - z['val'][sel]).max()), int(sel.sum()), scale, name)


def run_distributed_cg(a, kvs, geo, rank, world, barrier):
    """Config 5's solver stage on this run's mesh: preconditioned CG on the geometry mass matrix (the system of
    pyiga/approx.py:82-96), slab-distributed, device-resident (pyiga_b200.distcg: halo and dot products through
    peer windows over NVLink, CUDA graph per batch of iterations).  Collective: every rank calls it.  Not part of
    the timed assembly step; reported under the `cg` key."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pyiga_b200 import _device, assemble
    from pyiga_b200.dist import SlabAssembly
    from pyiga_b200.distcg import DistributedCG
    be = _device.backend()
    ok, cg, err_msg = 1.0, None, None
    try:
        sa = SlabAssembly(kvs, geo, 'mass', rank=rank, world=world, balance='entries')
        mlb = sa.assemble_mlb()
        Ainv = [np.linalg.inv(assemble.bsp_mass_1d(kv).toarray()) for kv in kvs]
        cg = DistributedCG(sa.dev.device_structure, mlb, sa.slabs, rank, Ainv)
    except Exception as exc:        # all ranks must agree before anyone waits on a peer
        ok, err_msg = 0.0, repr(exc)[:200]
    if world > 1:
        t = torch.tensor([ok], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = float(t[0])
    if not ok:
        if cg is not None:
            cg.close()
        return {'error': err_msg or 'setup failed on another rank'}
    nloc = cg.nloc
    ones = torch.ones(nloc, dtype=torch.float64, device='cuda')
    b = cg.matvec(ones).clone()
    y, x = be.empty(nloc), be.empty(nloc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        cg.matvec(ones, y)
    barrier()
    e0.record()
    for _ in range(20):
        cg.matvec(ones, y)
    e1.record()
    barrier()
    mv_ms = e0.elapsed_time(e1) / 20
    cg.solve(b, rtol=1e-10, maxiter=10, check_every=10, out=x)         # warm-up: graph capture
    times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        _, it, res = cg.solve(b, rtol=1e-10, maxiter=100, check_every=10, out=x)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times), mv_ms, float((x - 1.0).abs().max())], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {'workload': 'CG on the 3D mass matrix p=%d n=%d (%s), Kronecker preconditioner, rtol 1e-10' % (a.p, a.n, a.geo),
           'nnz': sa.dev.nnz, 'local_nnz_rank0': sa.local_nnz, 'cg_ms': 1e3 * float(t[0]), 'cg_iterations': int(it),
           'cg_rel_residual': float(res), 'max_err_vs_known_solution': float(t[2]),
           'matvec_ms': float(t[1]), 'matvec_GBps_local': 8.0 * sa.local_nnz / (float(t[1]) * 1e-3) / 1e9,
           'timing': 'cg_ms: host clock around pb200_cg_solve (min of 3, max over ranks); matvec_ms: CUDA events, max over ranks'}
    cg.close()
    return out


def run_extra_configs(be, fp64_peak, hbm_peak):
    """Secondary configurations of BASELINE.json on one GPU, device-resident like `value`: a few steps
    each, nnz/s and the fraction of the SURVEY 8d path bound max(F / FP64 peak, 8 nnz / HBM)."""
    import numpy as np
    import torch
    from pyiga_b200 import assemble, bspline, geometry
    from pyiga_b200.dist import SlabAssembly
    out = []

    def timed(step, reps=3):
        step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    for form, p, n in (('mass', 3, 128), ('stiffness', 3, 64), ('stiffness', 4, 192), ('mass', 4, 192)):
        try:
            kvs = 3 * (bspline.make_knots(p, 0.0, 1.0, n),)
            geo = geometry.twisted_nurbs_box()
            sa = SlabAssembly(kvs, geo, form)
            dev = sa.dev
            res = be.empty(dev.nnz)
            dev.compute_fields(geo)
            if not dev.uses_fused_fields():
                dev.fields              # the unfused pipeline keeps the field array (allocated on first use)
            budget = max(be.free_bytes() - (8 << 30), 1 << 30)
            chunks = dev.row_chunks(sa.rows, budget)
            ws = be.empty(max(dev.workspace_bytes(c) for c in chunks), np.uint8)
            ms = timed(lambda: sa.assemble_mlb(out=res, workspace=ws, tabulate=True))
            F, nnz = algorithmic_flops(p, n, form)
            t_bound = max(F / (fp64_peak * 1e12), 8.0 * nnz / (hbm_peak * 1e9))
            out.append({'workload': '3D %s p=%d n=%d, twisted NURBS box' % (form, p, n), 'nnz': int(nnz), 'ms_per_step': ms,
                        'value': nnz / (ms * 1e-3), 'unit': UNIT, 'row_chunks': len(chunks),
                        'fused': bool(dev.uses_fused_fields()),
                        'path_frac': t_bound / (ms * 1e-3), 'path_bound': 'fp64' if F / (fp64_peak * 1e12) > 8.0 * nnz / (hbm_peak * 1e9) else 'hbm'})
            del res, ws, sa, dev
            torch.cuda.empty_cache()
        except Exception as exc:
            out.append({'workload': '3D %s p=%d n=%d' % (form, p, n), 'error': repr(exc)[:200]})
    # config 5: convection-diffusion vform p=3 n=96 (B-spline twisted box), assembled to the MLB tensor
    try:
        form = '(inner(diff_coeff * grad(u), grad(v)) + inner((x[1], -x[0], 1.0), grad(u)) * v) * dx'
        kvs = 3 * (bspline.make_knots(3, 0.0, 1.0, 96),)
        geo = geometry.twisted_box()
        setups = []
        asm = None
        for _ in range(2):      # the first call pays for cudaMalloc of the 4.5 GB field buffer (the allocator cache is empty here)
            del asm
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            asm = assemble.instantiate_assembler(form, kvs, {'geo': geo, 'diff_coeff': lambda x, y, z: 1.0 + x * y}, None)
            torch.cuda.synchronize()
            setups.append(1e3 * (time.perf_counter() - t0))
        setup_ms = setups[-1]
        dev = asm.dev
        res = be.empty(dev.nnz)
        ws = be.empty(dev.workspace_bytes(), np.uint8)
        ms = timed(lambda: dev.assemble_mlb(out=res, workspace=ws))
        nnz = dev.nnz
        F = 2.33e11      # n_alpha = 3, n_beta = 4, c0 = 25
        t_bound = max(F / (fp64_peak * 1e12), 8.0 * nnz / (hbm_peak * 1e9))
        out.append({'workload': '3D convection-diffusion vform p=3 n=96, twisted_box B-spline', 'nnz': int(nnz), 'ms_per_step': ms,
                    'value': nnz / (ms * 1e-3), 'unit': UNIT, 'setup_ms_host_coefficients_and_fields': setup_ms,
                    'setup_ms_first_call': setups[0],
                    'path_frac': t_bound / (ms * 1e-3), 'path_bound': 'fp64'})
    except Exception as exc:
        out.append({'workload': '3D convection-diffusion vform p=3 n=96', 'error': repr(exc)[:200]})
    return out


def run_ours(a):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    from pyiga_b200 import _device, bspline, geometry
    from pyiga_b200.dist import SlabAssembly
    be = _device.backend()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kvs = 3 * (bspline.make_knots(a.p, 0.0, 1.0, a.n),)
    geo = geometry.twisted_nurbs_box() if a.geo == 'nurbs' else geometry.twisted_box()
    sa = SlabAssembly(kvs, geo, a.form, rank=rank, world=world)
    dev = sa.dev
    total_nnz = dev.nnz
    assert dev.fast_path, 'no sum-factorised kernels for this configuration'

    # ---- device-resident throughput ---------------------------------------------------------
    out = be.empty(max(sa.local_nnz, 1))
    budget = max(be.free_bytes() - (6 << 30), 1 << 30)
    chunks = dev.row_chunks(sa.rows, budget) if sa.rows else []
    ws_bytes = max([dev.workspace_bytes(c) for c in chunks] + [0])
    ws = be.empty(max(ws_bytes, 1), np.uint8)

    def step():
        if sa.rows is not None:
            sa.assemble_mlb(out=out, workspace=ws, tabulate=True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        step()
    launches0 = be.lib.pb200_launch_count()
    step()
    launches_per_step = be.lib.pb200_launch_count() - launches0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps

    # ---- per-kernel durations (second pass over the same steps, events inside the library) ----
    dev.set_timing(True)
    acc = {}
    fields_ms = []
    ef0, ef1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.steps):
        if sa.rows is None:
            break
        dev.tabulate()
        ef0.record()
        dev.compute_fields(geo, rows=sa.rows)
        ef1.record()
        dev.assemble_mlb(rows=sa.rows, out=out, workspace=ws)
        torch.cuda.synchronize()
        fields_ms.append(ef0.elapsed_time(ef1))
        for name, t in dev.stage_times():
            acc.setdefault(name, []).append(t)
    dev.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    stages = {k: sum(v) / len(v) for k, v in acc.items()}
    if fields_ms:
        stages['k1_geo_tables' if dev.uses_fused_fields() else 'k2_fields'] = sum(fields_ms) / len(fields_ms)

    # ---- parity: the matrix that was timed, against the reference's sampled entries -----------------
    par = parity_against_fixture(a, dev, sa.rows, out) if sa.rows is not None else None
    par_err, par_cnt, par_scale, par_name = par if par else (0.0, 0, 1.0, None)

    # ---- end to end through the public API: host descriptors in, scipy CSR matrix on the host out -----
    from pyiga_b200 import assemble as pb_assemble
    e2e_times, h2d, d2h, host_idx = [], 0, 0, 0
    pinned = None
    e2e_fn = getattr(pb_assemble, a.form)
    del ws
    torch.cuda.empty_cache()
    for it in range(a.e2e_steps + 1 if a.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            # the drop-in call of the reference arm: assemble.stiffness(kvs, geo) -> scipy.sparse.csr_matrix
            A = e2e_fn(kvs, geo)
            last = float(A.data[-1])                                      # the result is on the host
            d2h = A.data.nbytes
            host_idx = A.indices.nbytes + A.indptr.nbytes
            h2d = sum(8 * (kv.kv.size + 2 * g.size) for kv, g in zip(kvs, dev.gaussgrid)) + geo.coeffs.nbytes \
                + sum(8 * kv.kv.size for kv in geo.kvs)
            assert A.shape[0] == int(np.prod(dev.ndofs_test)) and A.nnz == total_nnz
            del A
        else:
            sl = SlabAssembly(kvs, geo, a.form, rank=rank, world=world)      # uploads knots, nodes, control net
            if sl.rows is not None:
                if pinned is None:
                    nrows_l, nnz_l, idt = sl.csr_sizes()
                    tdt = torch.int32 if idt == np.int32 else torch.int64
                    pinned = [torch.empty(nrows_l + 1, dtype=tdt, pin_memory=True), torch.empty(nnz_l, dtype=tdt, pin_memory=True),
                              torch.empty(nnz_l, dtype=torch.float64, pin_memory=True)]
                sl.assemble_csr_host(host=pinned)                           # chunked: D2H overlaps the next chunk
                d2h = pinned[2].numel() * pinned[2].element_size()          # values only; see e2e.what
                host_idx = sum(t.numel() * t.element_size() for t in pinned[:2])
                h2d = sum(8 * (kv.kv.size + 2 * g.size) for kv, g in zip(kvs, sl.dev.gaussgrid)) + geo.coeffs.nbytes \
                    + sum(8 * kv.kv.size for kv in geo.kvs)
            del sl
        t_call = time.perf_counter() - t0
        barrier()
        if it > 0:
            e2e_times.append(time.perf_counter() - t0)
        if os.environ.get('PB200_BENCH_DEBUG'):
            sys.stderr.write('[e2e rank %d it %d] call %.1f ms, with barrier %.1f ms\n'
                             % (rank, it, 1e3 * t_call, 1e3 * (time.perf_counter() - t0)))
    e2e_ms = 1e3 * sum(e2e_times) / len(e2e_times) if e2e_times else None

    # ---- config 5's solver stage (all ranks; outside the timed step) --------------------------
    cg_info = None
    if not a.no_cg:
        try:
            torch.cuda.empty_cache()
            cg_info = run_distributed_cg(a, kvs, geo, rank, world, barrier)
        except Exception as exc:
            cg_info = {'error': repr(exc)[:200]}

    # ---- reduce over ranks -------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, e2e_ms or 0.0, par_err], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, par_err = float(t[0]), (float(t[1]) or None), float(t[2])
        b = torch.tensor([float(h2d), float(d2h), float(launches_per_step), float(par_cnt), float(host_idx)], device='cuda',
                         dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        h2d, d2h, launches_per_step, par_cnt, host_idx = int(b[0]), int(b[1]), int(b[2]), int(b[3]), int(b[4])

    if rank == 0:
        g = C.c_double()
        _device.check(be.lib.pb200_probe_fp64(local, 100000, C.byref(g)))
        fp64_peak = g.value / 1e3                                         # TFLOP/s, measured DFMA probe
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        hbm_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback'
        value = total_nnz / (ms * 1e-3)
        F, _ = algorithmic_flops(a.p, a.n, a.form)
        t_bound = max(F / (fp64_peak * 1e12), 8.0 * total_nnz / (hbm_peak * 1e9)) / world
        # the algorithm's own FP64 work (rank 0's slab x world: the slabs are balanced)
        own = own_fp64_ops(dev, sa.rows, a.form, a.p)
        own_flops = 2.0 * sum(own.values()) * world
        path = {'bound': 'fp64' if F / (fp64_peak * 1e12) > 8.0 * total_nnz / (hbm_peak * 1e9) else 'hbm',
                'algorithmic_flops': F, 'flops_per_nnz': F / total_nnz, 'achieved': F / (ms * 1e-3) / 1e12,
                'peak': fp64_peak * world, 'unit': 'TFLOP/s', 'frac': t_bound / (ms * 1e-3),
                'peak_source': 'measured DFMA probe (pb200_probe_fp64) x %d GPUs' % world,
                'bound_nnz_per_s': total_nnz / t_bound,
                # SURVEY 8d's F counts ROW-WISE sum factorisation; the pairwise scheme here shares stages 1
                # and 2 between rows and computes one triangle of the symmetric form, so it needs fewer
                # flops: `frac` can exceed 1.  `frac_own` is this implementation's own FP64 work against the
                # same peak and cannot.
                'own_flops': own_flops, 'own_flops_parts_rank0': {k: 2.0 * v for k, v in own.items()},
                'own_flops_note': 'FMA/MUL count x2 of the fused pipeline (symmetric stage-1 windows, upper band entries of axis 0 only, 32-span batches with p overlap); geometry evaluation not counted',
                'own_achieved': own_flops / (ms * 1e-3) / 1e12, 'frac_own': own_flops / (ms * 1e-3) / 1e12 / (fp64_peak * world)}
        # dominant kernel of rank 0: the slower of its FP64 work at the measured DFMA peak and its
        # algorithmic bytes at the measured copy bandwidth bounds it
        roof = None
        if stages:
            dom = max(stages, key=stages.get)
            fused = dev.uses_fused_fields()
            pbytes = dict(pipeline_bytes(dev, sa.rows, a.form, a.p))
            if fused:
                pbytes.update(fused_bytes(dev, sa.rows, a.form))
            kflops = {'s1f': 2.0 * own['stage1'], 's1f_mass': 2.0 * own['stage1'],
                      's23': 2.0 * (own['stage23_phaseA'] + own['stage23_phaseB']),
                      's23_mass': 2.0 * (own['stage23_phaseA'] + own['stage23_phaseB'])}
            byts = pbytes.get(dom)
            kernel_gbs = {k: round(pbytes[k] / (stages[k] * 1e-3) / 1e9, 1) for k in stages if k in pbytes}
            kernel_tfs = {k: round(kflops[k] / (stages[k] * 1e-3) / 1e12, 2) for k in stages if k in kflops}
            traffic, step_traffic, stale = None, None, None
            try:    # DRAM bytes per launch from the committed ncu capture of the same workload
                t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
                if t.get('workload') == workload_name(a) and world == 1:
                    traffic = t['dram_bytes_per_launch'].get(dom)
                    if all(k in t['dram_bytes_per_launch'] for k in stages if not k.startswith('k1')):
                        step_traffic = sum(t['dram_bytes_per_launch'][k] for k in stages if k in t['dram_bytes_per_launch'])
                    stale = t.get('source_hash') != source_hash()
            except Exception:
                pass
            if byts:
                t_hbm = byts / (hbm_peak * 1e9)
                t_fp = kflops.get(dom, 0.0) / (fp64_peak * 1e12)
                hbm_bound = t_hbm >= t_fp
                ach = byts / (stages[dom] * 1e-3) / 1e9 if hbm_bound else kflops[dom] / (stages[dom] * 1e-3) / 1e12
                peak = hbm_peak if hbm_bound else fp64_peak
                roof = {'kernel': dom, 'bound': 'hbm' if hbm_bound else 'fp64', 'achieved': ach, 'peak': peak,
                        'unit': 'GB/s' if hbm_bound else 'TFLOP/s', 'frac': ach / peak, 'traffic': traffic,
                        'traffic_stale': stale, 'bytes_per_launch': byts, 'flops_per_launch': kflops.get(dom),
                        'hbm_frac': byts / (stages[dom] * 1e-3) / 1e9 / hbm_peak,
                        'fp64_frac': (kflops[dom] / (stages[dom] * 1e-3) / 1e12 / fp64_peak) if dom in kflops else None,
                        'ms_per_launch': stages[dom], 'peak_source': hbm_src if hbm_bound else 'measured DFMA probe (pb200_probe_fp64)',
                        'share_of_step': stages[dom] / sum(stages.values()),
                        'all_kernels_GBps': kernel_gbs, 'all_kernels_own_TFLOPs': kernel_tfs,
                        'step_bytes': sum(pbytes[k] for k in stages if k in pbytes),
                        'step_GBps': sum(pbytes[k] for k in stages if k in pbytes) / (ms * 1e-3) / 1e9,
                        'step_traffic': step_traffic,
                        'step_traffic_frac': (step_traffic / (ms * 1e-3) / 1e9 / hbm_peak) if step_traffic else None}
        tol = 1e-12
        parity = {'fixture': ('tests/golden/large_%s.npz (sampled pyiga multi_entries)' % par_name) if par_name else None,
                  'n': int(par_cnt), 'max_rel_err': (par_err / par_scale) if par_name else None, 'tol': tol,
                  'ok': (par_err <= tol * par_scale) if par_name else None}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': workload_name(a), 'nnz': total_nnz, 'ndofs': int(np.prod(dev.ndofs_test)),
                       'gauss_points': dev.npoints, 'parallelism': 'row slabs of the first tensor axis x%d, no collective' % world,
                       'cache': 'intermediate (%.1f GB) and output (%.1f GB) exceed the 126 MB L2; no flush needed'
                                % (1e-9 * dev.workspace_bytes(sa.rows), 8e-9 * total_nnz),
                       'timed_region': 'K1 basis tables + geometry/fields + contraction (fused kernels where eligible), MLB tensor resident in HBM'},
            'gpu_launches': int(launches_per_step) * a.steps,
            'launches_per_step': int(launches_per_step),
            'kernel_ms': {k: round(v, 4) for k, v in sorted(stages.items())},
            'parity_check': parity,
            'roofline': roof, 'path_roofline': path, 'clocks': clocks,
            'e2e': {'value': total_nnz / (e2e_ms * 1e-3) if e2e_ms else None, 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'host_index_bytes_per_step': int(host_idx),
                    'what': ('pyiga_b200.assemble.%s(kvs, geo) -> scipy.sparse.csr_matrix on the host: the call the reference arm times '
                             '(pyiga/assemble.py:1017-1049).  ' % a.form if world == 1 else
                             'SlabAssembly(kvs, geo, rank, world).assemble_csr_host() per rank (the multi-GPU entry point).  ')
                            + 'Tables H2D, K1 + fused assembly in row chunks, CSR value permutation, D2H of the values into pinned host '
                              'memory overlapped chunk by chunk; indptr/indices (closed form of the band tables) are written into the '
                              'result arrays by host threads meanwhile (pb200_csr_pattern_host)'},
        }
        if cg_info is not None:
            line['cg'] = cg_info
        if world == 1 and not a.no_extras:
            try:
                del out
                torch.cuda.empty_cache()
                line['extra_configs'] = run_extra_configs(be, fp64_peak, hbm_peak)
            except Exception as exc:
                line['extra_configs'] = [{'error': repr(exc)[:200]}]
        if world == 1 and not a.no_cpu_baseline:
            try:
                child = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2',
                                        '--warmup', '1', '--p', str(a.p), '--n', str(a.n), '--form', a.form, '--geo', a.geo,
                                        '--ref-n', str(a.ref_n)], capture_output=True, text=True, timeout=1200,
                                       env=dict(os.environ, RANK='0', WORLD_SIZE='1'))
                ref_line = json.loads(child.stdout.strip().splitlines()[-1])
                line['cpu_baseline'] = ref_line['cpu_baseline']
            except Exception as exc:    # the baseline is informational; never lose the GPU line over it
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'unavailable',
                                        'sample': 'failed: %r' % (exc,)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if par_name and not (par_err <= 1e-12 * par_scale):
        sys.stderr.write('PARITY FAILURE: max abs error %.3e > 1e-12 * %.3e\n' % (par_err, par_scale))
        sys.exit(3)


def source_hash():
    """hash of the CUDA sources: a committed ncu traffic figure is only current for the code it was taken from"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'pyiga_b200', 'csrc')
    for name in sorted(os.listdir(d)):
        if name.endswith(('.cu', '.cuh')):
            h.update(open(os.path.join(d, name), 'rb').read())
    return h.hexdigest()[:16]


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the process's real stdout"""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse_args()
    # Libraries write banners to the C-level stdout (NCCL prints its version there when NCCL_DEBUG is
    # set in the environment): route file descriptor 1 to stderr for the whole run and keep the real
    # stdout for the result line only.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
