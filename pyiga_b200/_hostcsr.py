"""Delivery of an assembled matrix as host CSR arrays — the last step of ``assemble.stiffness`` /
``mass`` / ``assemble(..., format='csr')`` (reference: COO -> CSR + mirror, ``pyiga/assemble.py:745-754``).

The rows of the first tensor axis are assembled in chunks; while chunk k+1 is being assembled the
CSR-ordered values of chunk k travel to the host over a copy stream, and host threads write
``indptr`` / ``indices`` — a closed form of the per-axis band tables — straight into the result
arrays (``pb200_csr_pattern_host``), so only 8 of the 12 B/nnz cross PCIe.

The result arrays live in pinned host memory taken from a small pool; they belong to the returned
matrix and go back to the pool when the matrix is garbage-collected (cudaHostAlloc of several GB
costs about a second, a recycled buffer nothing).
"""
import threading
import weakref

import numpy as np

_POOL_LIMIT_BYTES = 48 << 30        # largest result held in pinned memory; above: pageable numpy arrays
_pool = {}                          # (numel, dtype) -> [free pinned tensors]
_pool_lock = threading.Lock()
_POOL_KEEP = 2


def _release(t):
    with _pool_lock:
        free = _pool.setdefault((t.numel(), t.dtype), [])
        if len(free) < _POOL_KEEP:
            free.append(t)


def _pinned(torch, n, dtype):
    """pinned tensor of exactly n elements and its numpy view; the tensor returns to the pool when the
    view (and every array derived from it) is gone"""
    with _pool_lock:
        free = _pool.get((int(n), dtype))
        t = free.pop() if free else None
    if t is None:
        t = torch.empty(int(n), dtype=dtype, pin_memory=True)
    arr = t.numpy()
    weakref.finalize(arr, _release, t)
    return t, arr


def clear_pool():
    with _pool_lock:
        _pool.clear()


def csr_sizes(dev, rows=None):
    """(rows, nnz, index dtype) of the CSR arrays of the row slab `rows` of the first axis"""
    ra, rb = (0, dev.ndofs_test[0]) if rows is None else rows
    nrows = (rb - ra) * int(np.prod(dev.ndofs_test[1:], dtype=np.int64))
    nnz = dev.slab_size((ra, rb))
    ncols = int(np.prod(dev.ndofs_trial, dtype=np.int64))
    return nrows, nnz, (np.int32 if max(nnz, ncols) < 2 ** 31 else np.int64)


def host_cores():
    """cores this process may run on (the affinity mask where the platform has one, else the machine's count)"""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 2


def pattern_split(rs, ra, rb, nthr, inner_r, inner_b):
    """Rows [ra, rm) of the first axis for `nthr` pattern threads, rows [rm, rb) — about 1 / (nthr + 1) of the
    band entries — for the calling thread.  Returns (rm, None) or (rm, (rm, rb, first local row of the second
    part, first local entry of the second part))."""
    rm = rb
    if rb - ra >= 2:
        target = rs[ra] + (rs[rb] - rs[ra]) * nthr / (nthr + 1.0)
        rm = int(np.clip(np.searchsorted(rs, target), ra + 1, rb))
    if rm < rb:
        return rm, (rm, rb, (rm - ra) * inner_r, int(rs[rm] - rs[ra]) * inner_b)
    return rm, None


def assemble_csr_host(dev, rows=None, host=None, nchunks=8, workspace=None, pattern='host', pattern_threads=None,
                      timings=None):
    """Assemble the rows `rows` of the first axis (default: all) and deliver (indptr, indices, data) in
    host memory.  `host` may hold three preallocated (pinned) torch tensors; otherwise they come from
    the pool.  Returns (tensors, numpy views).  CUDA backend only."""
    import time
    tm = timings if timings is not None else {}
    t_start = time.perf_counter()
    be = dev.be
    torch = be.torch
    ra, rb = (0, dev.ndofs_test[0]) if rows is None else rows
    nrows, nnz, idt = csr_sizes(dev, (ra, rb))
    tdt = torch.int32 if idt == np.int32 else torch.int64
    views = None
    if host is None:
        total = nnz * (8 + np.dtype(idt).itemsize) + (nrows + 1) * np.dtype(idt).itemsize
        if total <= _POOL_LIMIT_BYTES:
            pairs = [_pinned(torch, nrows + 1, tdt), _pinned(torch, nnz, tdt), _pinned(torch, nnz, torch.float64)]
            host, views = [p[0] for p in pairs], [p[1] for p in pairs]
        else:       # very large results: pageable arrays (the driver stages the copies)
            views = [np.empty(nrows + 1, dtype=idt), np.empty(nnz, dtype=idt), np.empty(nnz, dtype=np.float64)]
            host = [torch.from_numpy(v) for v in views]
    rs = dev.row_start0()
    inner_b = int(np.prod(dev.nband[1:], dtype=np.int64))
    inner_r = int(np.prod(dev.ndofs_test[1:], dtype=np.int64))
    # chunks of rows balanced by band count, small enough for the device memory that is free
    nchunks = max(1, min(nchunks, rb - ra))
    while True:
        targets = np.linspace(rs[ra], rs[rb], nchunks + 1)
        cuts = sorted(set(int(np.clip(np.searchsorted(rs, t), ra, rb)) for t in targets) | {ra, rb})
        chunks = [(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]
        cmax_nnz = max(int(rs[b] - rs[a]) * inner_b for a, b in chunks)
        ws_need = max(dev.workspace_bytes(c) for c in chunks)
        need = ws_need * (workspace is None) + cmax_nnz * 8 * 3 + (64 << 20)
        if need <= be.free_bytes() or nchunks >= rb - ra:
            break
        nchunks = min(rb - ra, nchunks * 2)
    if workspace is None:
        workspace = be.empty(max(ws_need, 1), np.uint8)
    if not dev.uses_fused_fields():
        dev.need_fields((ra, rb))
    mlb = be.empty(cmax_nnz)
    stage = [be.empty(cmax_nnz) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=be.device)
    main = torch.cuda.current_stream(be.device)
    freed = [None, None]
    nnz_off = 0
    worker, err = None, []
    tm['setup_ms'] = 1e3 * (time.perf_counter() - t_start)
    ds = dev.device_structure
    own_share = None
    if pattern == 'host':
        # host cores of this rank (torchrun sets LOCAL_WORLD_SIZE): all but one write the pattern from the start;
        # the calling thread joins them with the last rows once every chunk is enqueued, instead of idling in the
        # stream synchronisation (with 8 ranks on 16 cores that doubles the writers of a rank)
        import os
        nthr = pattern_threads or max(1, host_cores() // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))) - 1)
        rm, own_share = pattern_split(rs, ra, rb, nthr, inner_r, inner_b)

        def fill():
            try:
                ds.csr_pattern_host(host[0], host[1], row0=(ra, rm), nthreads=nthr)
            except Exception as exc:        # surfaced after the join
                err.append(exc)
        worker = threading.Thread(target=fill)
        worker.start()
    else:
        cmax_rows = max((b - a) * inner_r for a, b in chunks)
        istage = [(be.empty(cmax_rows + 1, idt), be.empty(cmax_nnz, idt)) for _ in range(2)]
    row_off = 0
    for k, (a, b) in enumerate(chunks):
        cn = int(rs[b] - rs[a]) * inner_b
        cr = (b - a) * inner_r
        if freed[k % 2] is not None:
            main.wait_event(freed[k % 2])           # the staging buffers are free again
        dev.assemble_mlb(rows=(a, b), out=mlb, workspace=workspace)
        if worker is not None:
            vv = ds.csr_values(mlb, row0=(a, b), out=stage[k % 2])
        else:
            ip, ix, vv = ds.csr_arrays(mlb, row0=(a, b), out=istage[k % 2] + (stage[k % 2],), idt=idt)
            if nnz_off:
                ip += nnz_off
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            if worker is None:
                host[0][row_off:row_off + cr + 1].copy_(ip, non_blocking=True)
                host[1][nnz_off:nnz_off + cn].copy_(ix, non_blocking=True)
            host[2][nnz_off:nnz_off + cn].copy_(vv, non_blocking=True)
            freed[k % 2] = torch.cuda.Event()
            freed[k % 2].record(copy_stream)
        row_off += cr
        nnz_off += cn
    tm['enqueue_ms'] = 1e3 * (time.perf_counter() - t_start) - tm['setup_ms']
    if own_share is not None:
        rm, _, row_m, nnz_m = own_share
        ds.csr_pattern_host(host[0][row_m:], host[1][nnz_m:], row0=(rm, rb), indptr_offset=nnz_m, nthreads=1)
        tm['own_pattern_ms'] = 1e3 * (time.perf_counter() - t_start) - tm['setup_ms'] - tm['enqueue_ms']
    copy_stream.synchronize()
    main.synchronize()
    tm['device_done_ms'] = 1e3 * (time.perf_counter() - t_start)
    if worker is not None:
        worker.join()
        if err:
            raise err[0]
    tm['total_ms'] = 1e3 * (time.perf_counter() - t_start)
    return host, views


def assemble_csr_matrix(dev, nchunks=8):
    """The whole matrix as ``scipy.sparse.csr_matrix`` (float64 data, int32 indices unless
    nnz >= 2^31, sorted, canonical) through the pipelined delivery."""
    import scipy.sparse
    host, views = assemble_csr_host(dev, None, nchunks=nchunks)
    nrows, nnz, idt = csr_sizes(dev)
    ncols = int(np.prod(dev.ndofs_trial, dtype=np.int64))
    indptr, indices, data = views
    A = scipy.sparse.csr_matrix((data, indices, indptr), shape=(nrows, ncols), copy=False)
    A.has_sorted_indices = True
    A.has_canonical_format = True
    return A
