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
    ap.add_argument('--ref-n', type=int, default=32, help='spans per axis of the bounded CPU sample')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
def _reference_setup(a):
    """Returns (assemble_fn, kind, cores, sample_nnz): the reference's own implementation from
    oracle/_ref when it loads, else the C port of the oracle."""
    import numpy as np
    cores = os.cpu_count() or 1
    probe = subprocess.run([sys.executable, '-c',
                            'import sys; sys.path.insert(0, %r); import pyiga.assemblers' % os.path.join(ROOT, 'oracle', '_ref')],
                           capture_output=True)
    if probe.returncode == 0:
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        import pyiga
        from pyiga import assemble, bspline, geometry
        pyiga.set_max_threads(cores)
        kvs = 3 * (bspline.make_knots(a.p, 0.0, 1.0, a.ref_n),)
        G = geometry.twisted_box()
        if a.geo == 'nurbs':
            i, j, k = np.meshgrid(np.arange(2), np.arange(4), np.arange(2), indexing='ij')
            G = geometry.NurbsFunc(G.kvs, G.coeffs.copy(), 1.0 + 0.25 * ((i + 2 * j + 3 * k) % 3))
        fn = assemble.stiffness if a.form == 'stiffness' else assemble.mass
        return (lambda: fn(kvs, G)), 'reference', cores
    from oracle import c_oracle, pyiga_oracle as orc
    from pyiga_b200 import geometry
    G = geometry.twisted_nurbs_box() if a.geo == 'nurbs' else geometry.twisted_box()
    kv = orc.make_knots(a.p, 0.0, 1.0, a.ref_n)

    def run():
        prob = orc.Problem([kv] * 3, [a.p] * 3, [k.kv for k in G.kvs], [k.p for k in G.kvs], G.coeffs, G._rational)
        return c_oracle.assemble_csr(prob, a.form, nthreads=cores)
    return run, 'port', cores


def run_reference(a, quiet=False):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    fn, kind, cores = _reference_setup(a)
    for _ in range(a.warmup):
        fn()
    times = []
    nnz = 0
    for _ in range(max(a.steps, 1)):
        t0 = time.perf_counter()
        A = fn()
        times.append(time.perf_counter() - t0)
        nnz = A.nnz
    ms = 1e3 * sum(times) / len(times)
    value = nnz / (ms * 1e-3)
    sample = ('%s: %d nnz per step, end-to-end %s (setup + multi_entries + CSR) on %d host threads'
              % (workload_name(a, a.ref_n), nnz, 'pyiga.assemble.%s' % a.form if kind == 'reference' else 'oracle C port',
                 cores))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
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
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=f,
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

    # ---- end to end through the public API: host descriptors in, scipy-layout CSR on the host out
    rs = dev.row_start0()
    e2e_times, h2d, d2h = [], 0, 0
    pinned = None
    for it in range(a.e2e_steps + 1 if a.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        sl = SlabAssembly(kvs, geo, a.form, rank=rank, world=world)      # uploads knots, nodes, control net
        if sl.rows is not None:
            if pinned is None:
                nrows_l, nnz_l, idt = sl.csr_sizes()
                tdt = torch.int32 if idt == np.int32 else torch.int64
                pinned = [torch.empty(nrows_l + 1, dtype=tdt, pin_memory=True), torch.empty(nnz_l, dtype=tdt, pin_memory=True),
                          torch.empty(nnz_l, dtype=torch.float64, pin_memory=True)]
            sl.assemble_csr_host(host=pinned, workspace=ws)               # chunked: D2H overlaps the next chunk
            d2h = pinned[2].numel() * pinned[2].element_size()              # values only; see e2e.what
            h2d = sum(8 * (kv.kv.size + 2 * g.size) for kv, g in zip(kvs, sl.dev.gaussgrid)) + geo.coeffs.nbytes \
                + sum(8 * kv.kv.size for kv in geo.kvs)
        t_call = time.perf_counter() - t0
        barrier()
        if it > 0:
            e2e_times.append(time.perf_counter() - t0)
        if os.environ.get('PB200_BENCH_DEBUG'):
            sys.stderr.write('[e2e rank %d it %d] call %.1f ms, with barrier %.1f ms, phases %s\n'
                             % (rank, it, 1e3 * t_call, 1e3 * (time.perf_counter() - t0), getattr(sl, 'last_timings', None)))
        del sl
    e2e_ms = 1e3 * sum(e2e_times) / len(e2e_times) if e2e_times else None

    # ---- reduce over ranks -------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, e2e_ms or 0.0], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), (float(t[1]) or None)
        b = torch.tensor([float(h2d), float(d2h), float(launches_per_step)], device='cuda', dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        h2d, d2h, launches_per_step = int(b[0]), int(b[1]), int(b[2])

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
        path = {'bound': 'fp64' if F / (fp64_peak * 1e12) > 8.0 * total_nnz / (hbm_peak * 1e9) else 'hbm',
                'algorithmic_flops': F, 'flops_per_nnz': F / total_nnz, 'achieved': F / (ms * 1e-3) / 1e12,
                'peak': fp64_peak * world, 'unit': 'TFLOP/s', 'frac': t_bound / (ms * 1e-3),
                'peak_source': 'measured DFMA probe (pb200_probe_fp64) x %d GPUs' % world,
                'bound_nnz_per_s': total_nnz / t_bound}
        # dominant kernel of rank 0
        roof = None
        if stages:
            dom = max(stages, key=stages.get)
            pbytes = pipeline_bytes(dev, sa.rows, a.form, a.p)
            byts = pbytes.get(dom)
            kernel_gbs = {k: round(pbytes[k] / (stages[k] * 1e-3) / 1e9, 1) for k in stages if k in pbytes}
            traffic, step_traffic = None, None
            try:    # DRAM bytes per launch from the committed ncu capture of the same workload
                t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
                if t.get('workload') == workload_name(a) and world == 1:
                    traffic = t['dram_bytes_per_launch'].get(dom)
                    if all(k in t['dram_bytes_per_launch'] for k in stages):
                        step_traffic = sum(t['dram_bytes_per_launch'][k] for k in stages)
            except Exception:
                pass
            if byts:
                ach = byts / (stages[dom] * 1e-3) / 1e9
                roof = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                        'frac': ach / hbm_peak, 'traffic': traffic, 'bytes_per_launch': byts,
                        'ms_per_launch': stages[dom], 'peak_source': hbm_src,
                        'share_of_step': stages[dom] / sum(stages.values()),
                        'all_kernels_GBps': kernel_gbs, 'step_bytes': sum(pbytes.values()),
                        'step_GBps': sum(pbytes.values()) / (ms * 1e-3) / 1e9,
                        # all kernels of the step: DRAM traffic by ncu against the measured copy bandwidth
                        'step_traffic': step_traffic,
                        'step_traffic_frac': (step_traffic / (ms * 1e-3) / 1e9 / hbm_peak) if step_traffic else None}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': max(a.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': workload_name(a), 'nnz': total_nnz, 'ndofs': int(np.prod(dev.ndofs_test)),
                       'gauss_points': dev.npoints, 'parallelism': 'row slabs of the first tensor axis x%d, no collective' % world,
                       'cache': 'inputs (%.1f GB fields) and outputs (%.1f GB) exceed the 126 MB L2; no flush needed'
                                % (8e-9 * dev.nfields * dev.npoints, 8e-9 * total_nnz),
                       'timed_region': 'K1 basis tables + K2 geometry/fields + K3 contraction, MLB tensor resident in HBM'},
            'gpu_launches': int(launches_per_step) * a.steps,
            'launches_per_step': int(launches_per_step),
            'kernel_ms': {k: round(v, 4) for k, v in sorted(stages.items())},
            'roofline': roof, 'path_roofline': path, 'clocks': clocks,
            'e2e': {'value': total_nnz / (e2e_ms * 1e-3) if e2e_ms else None, 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'host_index_bytes_per_step': int(sum(t.numel() * t.element_size() for t in pinned[:2])) if pinned else 0,
                    'what': 'SlabAssembly(kvs, geo).assemble_csr_host(): tables H2D, K1+K2+K3 in row chunks, CSR value permutation, D2H of the values into pinned host memory overlapped chunk by chunk; indptr/indices (closed form of the band tables) are written into the pinned host arrays by host threads meanwhile (pb200_csr_pattern_host)'},
        }
        if world == 1 and not a.no_cpu_baseline:
            try:
                child = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '3',
                                        '--warmup', '1', '--p', str(a.p), '--form', a.form, '--geo', a.geo,
                                        '--ref-n', str(a.ref_n)], capture_output=True, text=True, timeout=900,
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
