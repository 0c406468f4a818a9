// Host side of the slab-distributed CG (distcg.cuh): windows shared between the ranks of a node and
// the solver object.  Included at the end of api.cu.
#include "distcg.cuh"
#ifdef PB_EMULATE
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#endif

struct pb200_comm {
    int device = 0, rank = 0, world = 1;
    size_t bytes = 0;
    char* base[PB_CG_MAXPEERS] = {nullptr};
#ifdef PB_EMULATE
    char name[64] = {0};
#endif
};

// A window is device memory allocated with cudaMalloc (not by the caller's allocator) so that it can be
// exported through CUDA IPC; the emulation build uses POSIX shared memory, which gives the CPU tests the
// same multi-process semantics.
extern "C" int pb200_comm_create(int device, int rank, int world, size_t bytes, pb200_comm** out) {
    if (!out || world < 1 || world > PB_CG_MAXPEERS || rank < 0 || rank >= world) return fail(PB200_EINVAL, "invalid communicator arguments");
    std::unique_ptr<pb200_comm> c(new pb200_comm);
    c->device = device; c->rank = rank; c->world = world; c->bytes = bytes;
#ifdef PB_EMULATE
    static int counter = 0;
    snprintf(c->name, sizeof c->name, "/pb200_%d_%d", (int)getpid(), counter++);
    int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) return fail(PB200_ENOMEM, "shared memory window of %zu bytes failed", bytes);
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(PB200_ENOMEM, "mmap of the window failed");
    memset(p, 0, bytes);
    c->base[rank] = (char*)p;
#else
    CK(cudaSetDevice(device));
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemset(p, 0, bytes));
    CK(cudaDeviceSynchronize());
    c->base[rank] = (char*)p;
#endif
    *out = c.release();
    return 0;
}

// 64-byte handle of the own window, to be passed to the other ranks (any transport)
extern "C" int pb200_comm_handle(pb200_comm* c, void* handle64) {
    if (!c || !handle64) return fail(PB200_EINVAL, "null argument");
    memset(handle64, 0, 64);
#ifdef PB_EMULATE
    memcpy(handle64, c->name, sizeof c->name);
#else
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle does not fit");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->base[c->rank]));
    memcpy(handle64, &h, sizeof h);
#endif
    return 0;
}

// map the windows of all ranks (handles: world x 64 bytes, in rank order)
extern "C" int pb200_comm_open_peers(pb200_comm* c, const void* handles) {
    if (!c || !handles) return fail(PB200_EINVAL, "null argument");
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) continue;
        const char* h = (const char*)handles + 64 * q;
#ifdef PB_EMULATE
        int fd = shm_open(h, O_RDWR, 0600);
        if (fd < 0) return fail(PB200_EINVAL, "cannot open the window of rank %d", q);
        void* p = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return fail(PB200_ENOMEM, "mmap of the window of rank %d failed", q);
        c->base[q] = (char*)p;
#else
        CK(cudaSetDevice(c->device));
        cudaIpcMemHandle_t ih;
        memcpy(&ih, h, sizeof ih);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        c->base[q] = (char*)p;
#endif
    }
    return 0;
}

extern "C" int pb200_comm_destroy(pb200_comm* c) {
    if (!c) return 0;
    for (int q = 0; q < c->world; ++q) {
        if (!c->base[q]) continue;
#ifdef PB_EMULATE
        munmap(c->base[q], c->bytes);
#else
        if (q == c->rank) cudaFree(c->base[q]); else cudaIpcCloseMemHandle(c->base[q]);
#endif
    }
#ifdef PB_EMULATE
    shm_unlink(c->name);
#endif
    delete c;
    return 0;
}

struct pb200_cg {
    PbCgDev d;
    PbMlbParams mp;
    pb200_comm* comm = nullptr;
    bool own_comm = false;
    void* mem = nullptr;        // device-local buffers
    int nblocks = 0, nb_lo = 0, nb_hi = 0;
    int grid0x = 0, grid0y = 0;
#ifndef PB_EMULATE
    cudaGraphExec_t graph = nullptr;
    int graph_iters = 0;
    const double* graph_x = nullptr;
    cudaStream_t own_stream = nullptr;      // the solve runs on a stream of its own (the caller's may be the legacy
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;     // default stream, which cannot be captured), ordered by events
#endif
};

// window bytes a solver with these slabs needs (identical on all ranks)
static size_t cg_window_layout(PbCgDev& d) {
    size_t o = 0;
    auto take = [&](size_t b) { size_t at = (o + 255) & ~size_t(255); o = at + b; return at; };
    const size_t pbytes = (size_t)(2 * d.hmax + d.lmax) * d.plane * sizeof(double);
    for (int b = 0; b < 2; ++b) d.off_pext[b] = take(pbytes);
    for (int b = 0; b < 2; ++b) d.off_gath[b] = take((size_t)d.N0 * d.plane * sizeof(double));
    d.off_hflag = take(2 * sizeof(long long));
    d.off_gflag = take((size_t)d.world * sizeof(long long));
    d.off_rstamp = take((size_t)2 * d.world * sizeof(long long));
    d.off_rval = take((size_t)2 * d.world * PB_CG_NRED * sizeof(double));
    return (o + 255) & ~size_t(255);
}

extern "C" int pb200_cg_window_bytes(const pb200_mlstruct* S, int world, const int* cuts, int halo, size_t* bytes) {
    if (!S || !cuts || !bytes || world < 1 || world > PB_CG_MAXPEERS) return fail(PB200_EINVAL, "invalid argument");
    PbCgDev d;
    memset(&d, 0, sizeof d);
    d.world = world; d.hmax = halo; d.N0 = S->Nv[0];
    d.plane = 1;
    for (int k = 1; k < S->dim; ++k) d.plane *= S->Nu[k];
    for (int q = 0; q < world; ++q) d.lmax = std::max(d.lmax, cuts[q + 1] - cuts[q]);
    *bytes = cg_window_layout(d);
    return 0;
}

// S: structure of the (square, same test and trial space) matrix; d_mlb: the slab of rows
// [cuts[rank], cuts[rank+1]); d_Ainv: dense inverses (N_k x N_k, row-major) of the Kronecker preconditioner's
// factors on the device; comm: window of pb200_cg_window_bytes bytes with the peers opened (NULL for world 1:
// the solver allocates a private one); halo: number of planes the band reaches into the neighbours.
extern "C" int pb200_cg_create(const pb200_mlstruct* S, int rank, int world, const int* cuts, int halo, const double* d_mlb,
                               const double* const* d_Ainv, pb200_comm* comm, pb200_cg** out) {
    if (!S || !cuts || !d_mlb || !d_Ainv || !out) return fail(PB200_EINVAL, "null argument");
    if (S->dim != 3) return fail(PB200_EUNSUPPORTED, "the distributed solver is implemented for 3D tensor-product matrices");
    for (int k = 0; k < 3; ++k) if (S->Nv[k] != S->Nu[k]) return fail(PB200_EINVAL, "square matrix (same test and trial space) expected");
    if (world < 1 || world > PB_CG_MAXPEERS || rank < 0 || rank >= world) return fail(PB200_EINVAL, "invalid rank / world");
    for (int q = 0; q < world; ++q)
        if (cuts[q + 1] - cuts[q] < std::max(halo, 1)) return fail(PB200_EINVAL, "slab %d is thinner than the band (%d rows)", q, halo);
    if (cuts[0] != 0 || cuts[world] != S->Nv[0]) return fail(PB200_EINVAL, "the slabs do not cover the rows of the first axis");
    CK(pbSetDevice(S->device));
    std::unique_ptr<pb200_cg> g(new pb200_cg);
    PbCgDev& d = g->d;
    memset(&d, 0, sizeof d);
    d.rank = rank; d.world = world; d.ra = cuts[rank]; d.rb = cuts[rank + 1];
    d.hmax = halo; d.N0 = S->Nv[0];
    for (int q = 0; q <= world; ++q) d.cuts[q] = cuts[q];
    for (int q = 0; q < world; ++q) d.lmax = std::max(d.lmax, cuts[q + 1] - cuts[q]);
    for (int k = 0; k < 3; ++k) { d.N[k] = S->Nv[k]; d.Ainv[k] = d_Ainv[k]; }
    d.plane = (long long)S->Nu[1] * S->Nu[2];
    d.nloc = (long long)(d.rb - d.ra) * d.plane;
    const size_t wbytes = cg_window_layout(d);
    if (!comm) {
        if (world != 1) return fail(PB200_EINVAL, "a window shared with the peers is needed for world > 1");
        int rc = pb200_comm_create(S->device, 0, 1, wbytes, &comm);
        if (rc) return rc;
        g->own_comm = true;
    }
    if (comm->bytes < wbytes || comm->world != world || comm->rank != rank) return fail(PB200_EINVAL, "window too small or of another communicator");
    g->comm = comm;
    for (int q = 0; q < world; ++q) {
        if (!comm->base[q]) return fail(PB200_EINVAL, "the window of rank %d has not been opened", q);
        d.win[q] = comm->base[q];
    }
    int rc = fill_mlb_params(S, d.ra, d.rb, d_mlb, g->mp);
    if (rc) return rc;
    g->nblocks = (int)((d.nloc + 255) / 256);
    // row blocks that touch halo planes (scheduled last by the matvec kernel)
    const long long hrows = (long long)d.hmax * d.plane;
    g->nb_lo = rank > 0 ? (int)((hrows + 255) / 256) : 0;
    g->nb_hi = rank + 1 < world ? (int)((hrows + 255) / 256) : 0;
    // local buffers: x is the caller's; r z Ap t1 | part | scal | ctl
    const size_t nl = (size_t)d.nloc;
    // grid of the tiled mode-0 kernel: (plane / 128) x (slab rows / PB_CG_TI)
    g->grid0x = (int)((d.plane + 127) / 128);
    g->grid0y = (d.rb - d.ra + PB_CG_TI - 1) / PB_CG_TI;
    d.nblocks2 = g->grid0x * g->grid0y;
    const size_t n2sq = (size_t)d.N[2] * d.N[2];
    size_t bytes = (4 * nl + (size_t)PB_CG_NRED * g->nblocks + d.nblocks2 + n2sq + 16) * sizeof(double) + 8 * sizeof(long long) + 1024;
    CK(pbMalloc(&g->mem, bytes));
#ifdef PB_EMULATE
    memset(g->mem, 0, bytes);
#else
    CK(cudaMemset(g->mem, 0, bytes));
#endif
    double* b = (double*)g->mem;
    d.r = b; d.z = b + nl; d.Ap = b + 2 * nl; d.t1 = b + 3 * nl;
    d.part = b + 4 * nl;
    d.part2 = d.part + (size_t)PB_CG_NRED * g->nblocks;
    double* at2 = d.part2 + d.nblocks2;
    d.AinvT2 = at2;
    d.scal = at2 + n2sq;
    d.ctl = reinterpret_cast<long long*>(d.scal + 16);
#ifndef PB_EMULATE
    {   // transposed copy of the last factor (once)
        std::vector<double> hA(n2sq), hT(n2sq);
        CK(cudaMemcpy(hA.data(), d_Ainv[2], n2sq * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < d.N[2]; ++i)
            for (int j = 0; j < d.N[2]; ++j) hT[(size_t)j * d.N[2] + i] = hA[(size_t)i * d.N[2] + j];
        CK(cudaMemcpy(at2, hT.data(), n2sq * sizeof(double), cudaMemcpyHostToDevice));
    }
#endif
    *out = g.release();
    return 0;
}

extern "C" int pb200_cg_destroy(pb200_cg* g) {
    if (!g) return 0;
#ifndef PB_EMULATE
    if (g->graph) cudaGraphExecDestroy(g->graph);
    if (g->own_stream) cudaStreamDestroy(g->own_stream);
    if (g->ev_in) cudaEventDestroy(g->ev_in);
    if (g->ev_out) cudaEventDestroy(g->ev_out);
#endif
    if (g->mem) pbFree(g->mem);
    if (g->own_comm) pb200_comm_destroy(g->comm);
    delete g;
    return 0;
}

#ifdef PB_EMULATE
template <class F> static double cg_emu_blocks(const PbCgDev& d, int nblocks, double* part, F&& body, int shift = 0) {
    for (int v = 0; v < nblocks; ++v) {
        const int bk = (v + shift) % nblocks;
        double s = 0.0;
        for (int i = 0; i < 256; ++i) {
            const long long t = (long long)bk * 256 + i;
            if (t < d.nloc) s += body(t, bk);
        }
        if (part) part[bk] = s;
    }
    return 0.0;
}
#endif

// kernels of one phase of the algorithm, enqueued on `st`
static void cg_allreduce(pb200_cg* g, int kind, pbStream st) {
    PbCgDev& d = g->d;
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    if (d.ctl[0] && kind >= 2) return;
    double loc[PB_CG_NRED];
    for (int k = 0; k < PB_CG_NRED; ++k) {
        loc[k] = 0.0;
        for (int i = 0; i < g->nblocks; ++i) loc[k] += d.part[(long long)k * g->nblocks + i];
    }
    pb_cg_allreduce_finish(d, kind, loc);
#else
    pb_cg_allreduce_kernel<<<1, 256, 0, st>>>(d, kind, g->nblocks);
#endif
}

static void cg_precond(pb200_cg* g, int second, pbStream st) {
    PbCgDev& d = g->d;
    g_launches += 3;
#ifdef PB_EMULATE
    (void)st;
    if (d.ctl[0]) return;
    for (long long t = 0; t < d.nloc; ++t) pb_cg_mode2_elem(d, t);
    for (long long t = 0; t < d.nloc; ++t) pb_cg_mode1_elem(d, t);
    pb_cg_mode1_signal(d);
    if (d.world > 1)
        for (int q = 0; q < d.world; ++q) pb_cg_wait(pb_cg_flag(d, d.rank, d.off_gflag) + q, d.ctl[2]);
    cg_emu_blocks(d, g->nblocks, d.part + (second ? g->nblocks : 0), [&](long long t, int) { return pb_cg_mode0_elem(d, t); });
#else
    {
        const long long rows2 = d.nloc / d.N[2];
        const size_t sm2 = (size_t)PB_CG_TI * d.N[2] * sizeof(double);
        pb_cg_mode2_tiled_kernel<<<(unsigned)((rows2 + PB_CG_TI - 1) / PB_CG_TI), 128, sm2, st>>>(d);
        const dim3 g1((unsigned)((d.N[2] + 127) / 128), (unsigned)((d.N[1] + PB_CG_TI - 1) / PB_CG_TI), (unsigned)(d.rb - d.ra));
        pb_cg_modek_tiled_kernel<1><<<g1, 128, (size_t)PB_CG_TI * d.N[1] * sizeof(double), st>>>(d, second);
        const dim3 g0((unsigned)g->grid0x, (unsigned)g->grid0y, 1);
        pb_cg_modek_tiled_kernel<0><<<g0, 128, (size_t)PB_CG_TI * d.N0 * sizeof(double), st>>>(d, second);
    }
#endif
}

static void cg_direction(pb200_cg* g, int first, pbStream st) {
    PbCgDev& d = g->d;
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    if (d.ctl[0]) return;
    for (long long t = 0; t < d.nloc; ++t) pb_cg_direction_elem(d, t, first != 0);
    pb_cg_direction_signal(d);
#else
    pb_cg_direction_kernel<<<g->nblocks, 256, 0, st>>>(d, first);
#endif
}

static void cg_iteration(pb200_cg* g, pbStream st) {
    PbCgDev& d = g->d;
    g_launches += 2;
#ifdef PB_EMULATE
    if (d.ctl[0]) return;
    for (int v = 0; v < g->nblocks; ++v) {      // same block order as the kernel: halo blocks last
        const int bk = (v + g->nb_lo) % g->nblocks;
        const long long r0 = (long long)bk * 256, r1 = std::min<long long>(r0 + 256, d.nloc);
        bool lo, hi;
        pb_cg_block_needs(d, r0, r1, lo, hi);
        if (lo) pb_cg_wait(pb_cg_flag(d, d.rank, d.off_hflag) + 0, d.ctl[2]);
        if (hi) pb_cg_wait(pb_cg_flag(d, d.rank, d.off_hflag) + 1, d.ctl[2]);
        double s = 0.0;
        for (long long r = r0; r < r1; ++r) s += pb_cg_matvec_row(d, g->mp, r);
        d.part[bk] = s;
    }
#else
    pb_cg_matvec_kernel<<<g->nblocks, 256, 0, st>>>(d, g->mp, g->nb_lo, g->nb_hi);
#endif
    cg_allreduce(g, 2, st);
#ifdef PB_EMULATE
    if (!d.ctl[0]) cg_emu_blocks(d, g->nblocks, d.part, [&](long long t, int) { return pb_cg_update_elem(d, t); });
#else
    pb_cg_update_kernel<<<g->nblocks, 256, 0, st>>>(d);
#endif
    cg_precond(g, 1, st);
    cg_allreduce(g, 3, st);
    cg_direction(g, 0, st);
}

// Solve A x = b for the local slabs (device pointers, nloc entries each; x0 = 0).  The host reads the
// convergence flag every `check_every` iterations (one CUDA graph launch per batch).  Returns the number of
// iterations and the relative residual ||r|| / ||b|| of the last one.
extern "C" int pb200_cg_solve(pb200_cg* g, const double* d_b, double* d_x, double rtol, int maxiter, int check_every,
                              int* iters, double* relres, void* stream) {
    if (!g || !d_b || !d_x) return fail(PB200_EINVAL, "null argument");
    if (maxiter < 1 || check_every < 1) return fail(PB200_EINVAL, "maxiter and check_every must be positive");
    PbCgDev& d = g->d;
    d.x = d_x;
    pbStream st = (pbStream)stream;
#ifdef PB_EMULATE
    d.ctl[0] = 0; d.ctl[1] = 0; d.scal[6] = rtol;
    cg_emu_blocks(d, g->nblocks, d.part, [&](long long t, int) { return pb_cg_init_elem(d, d_b, t); });
    for (int i = 0; i < g->nblocks; ++i) d.part[g->nblocks + i] = 0.0;
    ++g_launches;
    cg_allreduce(g, 0, st);
    d.ctl[2] += 1;                      // fresh stamps for this solve
    cg_precond(g, 0, st);
    cg_allreduce(g, 1, st);
    cg_direction(g, 1, st);
    while (!d.ctl[0] && d.ctl[1] < maxiter) cg_iteration(g, st);
    if (iters) *iters = (int)d.ctl[1];
    if (relres) *relres = d.scal[5] > 0 ? std::sqrt(d.scal[4] / d.scal[5]) : 0.0;
    return 0;
#else
    CK(pbSetDevice(g->comm->device));
    if (!g->own_stream) {
        CK(cudaStreamCreateWithFlags(&g->own_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&g->ev_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g->ev_out, cudaEventDisableTiming));
    }
    cudaStream_t caller = st;
    CK(cudaEventRecord(g->ev_in, caller));
    st = g->own_stream;
    CK(cudaStreamWaitEvent(st, g->ev_in, 0));
    // control words: done = 0, iterations = 0, tolerance; the epoch advances by one (fresh stamps)
    long long h_ctl[2] = {0, 0};
    CK(cudaMemcpyAsync(d.ctl, h_ctl, sizeof h_ctl, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d.scal + 6, &rtol, sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));      // h_ctl / rtol are stack variables
    pb_cg_init_kernel<<<g->nblocks, 256, 0, st>>>(d, d_b);
    ++g_launches;
    cg_allreduce(g, 0, st);
    pb_cg_bump_epoch_kernel<<<1, 1, 0, st>>>(d);
    cg_precond(g, 0, st);
    cg_allreduce(g, 1, st);
    cg_direction(g, 1, st);
    CK(pbLastError());
    // a batch of `check_every` iterations as one graph (re-captured when the batch size or x changes)
    const double*& graph_x = g->graph_x;
    if (!g->graph || g->graph_iters != check_every || graph_x != d_x) {
        if (g->graph) { cudaGraphExecDestroy(g->graph); g->graph = nullptr; }
        cudaGraph_t gr = nullptr;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < check_every; ++i) cg_iteration(g, st);
        CK(cudaStreamEndCapture(st, &gr));
        CK(cudaGraphInstantiate(&g->graph, gr, 0));
        cudaGraphDestroy(gr);
        g->graph_iters = check_every; graph_x = d_x;
    }
    long long h[2] = {0, 0};
    double h_s[6];
    for (int done_iters = 0; done_iters < maxiter; done_iters += check_every) {
        CK(cudaGraphLaunch(g->graph, st));
        CK(cudaMemcpyAsync(h, d.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h[0]) break;
    }
    CK(cudaMemcpyAsync(h, d.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_s, d.scal, sizeof h_s, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(g->ev_out, st));
    CK(cudaStreamWaitEvent(caller, g->ev_out, 0));
    if (iters) *iters = (int)h[1];
    if (relres) *relres = h_s[5] > 0 ? std::sqrt(h_s[4] / h_s[5]) : 0.0;
    int timed_out = 0;
    CK(cudaMemcpyFromSymbol(&timed_out, pb_cg_timed_out, sizeof(int)));
    if (timed_out) return fail(PB200_ECUDA, "distributed CG: a peer rank did not answer within %.0f s (halo / all-reduce flag)", PB_CG_WAIT_NS * 1e-9);
    return 0;
#endif
}

// y_local = A_slab p for a slab-distributed vector (halo exchange through the window, no CG): for tests
// and bandwidth measurements.  d_p: nloc entries.
extern "C" int pb200_cg_matvec(pb200_cg* g, const double* d_p, double* d_y, void* stream) {
    if (!g || !d_p || !d_y) return fail(PB200_EINVAL, "null argument");
    PbCgDev& d = g->d;
    pbStream st = (pbStream)stream;
    // load p as the direction of a new epoch: z := p, beta unused
    double* keep_z = d.z;
    double* keep_Ap = d.Ap;
    d.z = const_cast<double*>(d_p);
    d.Ap = d_y;
#ifdef PB_EMULATE
    d.ctl[0] = 0;
    d.ctl[2] += 1;
    cg_direction(g, 1, st);
    for (int v = 0; v < g->nblocks; ++v) {
        const int bk = (v + g->nb_lo) % g->nblocks;
        const long long r0 = (long long)bk * 256, r1 = std::min<long long>(r0 + 256, d.nloc);
        bool lo, hi;
        pb_cg_block_needs(d, r0, r1, lo, hi);
        if (lo) pb_cg_wait(pb_cg_flag(d, d.rank, d.off_hflag) + 0, d.ctl[2]);
        if (hi) pb_cg_wait(pb_cg_flag(d, d.rank, d.off_hflag) + 1, d.ctl[2]);
        for (long long r = r0; r < r1; ++r) pb_cg_matvec_row(d, g->mp, r);
    }
#else
    CK(pbSetDevice(g->comm->device));
    CK(cudaMemsetAsync(d.ctl, 0, sizeof(long long), st));       // done = 0
    pb_cg_bump_epoch_kernel<<<1, 1, 0, st>>>(d);
    pb_cg_direction_kernel<<<g->nblocks, 256, 0, st>>>(d, 1);
    pb_cg_matvec_kernel<<<g->nblocks, 256, 0, st>>>(d, g->mp, g->nb_lo, g->nb_hi);
    CK(pbLastError());
#endif
    d.z = keep_z;
    d.Ap = keep_Ap;
    return 0;
}
