// Slab-distributed, device-resident preconditioned conjugate gradients on a multi-level banded matrix.
//
// Reference semantics: `scipy.sparse.linalg.cg(M, b, M=KroneckerOperator(*Minvs))` on the geometry mass
// matrix as in `project_L2` (pyiga/approx.py:82-96), with `ml_matvec_3d` as the operator
// (pyiga/mlmatrix_cy.pyx:295-325) and `apply_kronecker` as the preconditioner (pyiga/kronecker.py:15-34).
//
// Every rank owns the rows [ra, rb) of the first tensor axis.  All communication goes through
// a WINDOW of device memory that the ranks of one node map into each other's address space (CUDA IPC:
// loads and stores travel over NVLink / NVSwitch); no host takes part in an iteration:
//   * halo of the matvec: the kernel that produces the new search direction p stores its boundary planes
//     directly into the halo regions of the neighbours' p buffers and then raises a flag there; the matvec
//     blocks that need the halo are scheduled last and spin on that flag while interior rows are already
//     being multiplied;
//   * dot products: every rank stores its partial sum into a slot of every peer's window, waits for all
//     slots and adds them in rank order (bitwise identical on all ranks, no atomics);
//   * Kronecker preconditioner A0^-1 (x) A1^-1 (x) A2^-1: the mode products of axes 1 and 2 are local to
//     the slab; their result is stored into a gather buffer of every peer, and only the (dense) mode-0
//     product reads the other slabs.
// Scalars (alpha, beta, residual), the iteration counter and the convergence flag live in device memory;
// kernels of later iterations return immediately once the flag is set, so a fixed batch of iterations can
// be captured in a CUDA graph and the host only looks at the flag between batches.
#pragma once
#include "mlb.cuh"

#define PB_CG_MAXPEERS 16
#define PB_CG_NRED 2            // values per all-reduce
#define PB_CG_TI 8              // outputs per thread of the tiled mode products

struct PbCgDev {
    int rank, world;
    int ra, rb;                 // my rows of axis 0
    int hmax;                   // halo planes on each side of the p buffers
    int N0, lmax;               // rows of axis 0, tallest slab
    long long plane;            // entries per plane of axis 0 (trial = test space)
    long long nloc;             // (rb - ra) * plane
    int cuts[PB_CG_MAXPEERS + 1];
    // device-local state
    double* x; double* r; double* z; double* Ap; double* t1;
    double* part;               // per-block partial sums [PB_CG_NRED][nblocks]
    double* scal;               // 0 rz, 1 pAp, 2 alpha, 3 beta, 4 rr, 5 bnorm2, 6 rtol
    long long* ctl;             // 0 done, 1 iterations, 2 epoch of the current p, 3 all-reduce calls
    // windows: win[q] = base of rank q's window as mapped here; the layout is the same on all ranks
    char* win[PB_CG_MAXPEERS];
    size_t off_pext[2];         // p buffers: (hmax + lmax + hmax) planes each
    size_t off_gath[2];         // gather buffers of the preconditioner: N0 planes each
    size_t off_hflag;           // long long [2]: stamp of the halo from below / above
    size_t off_gflag;           // long long [world]
    size_t off_rstamp;          // long long [2][world]
    size_t off_rval;            // double [2][world][PB_CG_NRED]
    const double* Ainv[3];      // dense inverses of the 1D factors
    const double* AinvT2;       // transposed copy of Ainv[2] (coalesced reads in the mode-2 product)
    double* part2;              // per-block partial sums of <r, z> (the tiled mode-0 kernel has its own grid)
    int nblocks2;
    int N[3];
};

PB_HD double* pb_cg_pext(const PbCgDev& c, int q, int b) { return reinterpret_cast<double*>(c.win[q] + c.off_pext[b]); }
PB_HD double* pb_cg_gath(const PbCgDev& c, int q, int b) { return reinterpret_cast<double*>(c.win[q] + c.off_gath[b]); }
PB_HD volatile long long* pb_cg_flag(const PbCgDev& c, int q, size_t off) { return reinterpret_cast<volatile long long*>(c.win[q] + off); }

#include <atomic>
PB_HD void pb_cg_fence() {
#if defined(__CUDA_ARCH__)
    __threadfence_system();
#else
    std::atomic_thread_fence(std::memory_order_seq_cst);
#endif
}
#define PB_CG_FENCE() pb_cg_fence()

// A peer that never raises its flag (a rank that died, a window mapped wrongly) must not hang the GPU: the
// device-side wait gives up after PB_CG_WAIT_NS and records it; the host entry points turn that into an error.
#define PB_CG_WAIT_NS 5000000000ULL
#if defined(__CUDACC__)
__device__ int pb_cg_timed_out = 0;
#endif
PB_HD void pb_cg_wait(volatile long long* f, long long stamp) {
#if defined(__CUDA_ARCH__)
    if (*f < stamp) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f < stamp) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > PB_CG_WAIT_NS || *(volatile int*)&pb_cg_timed_out) { pb_cg_timed_out = 1; break; }
        }
    }
#else
    while (*f < stamp) {}
#endif
    PB_CG_FENCE();
}

// ---- element-wise bodies (thread tid of nloc) -------------------------------------------------------
// r = b, x = 0, value for <b, b>
PB_HD double pb_cg_init_elem(const PbCgDev& c, const double* b, long long t) {
    const double v = b[t];
    c.r[t] = v;
    c.x[t] = 0.0;
    return v * v;
}
// x += alpha p, r -= alpha Ap, value for <r, r>
PB_HD double pb_cg_update_elem(const PbCgDev& c, long long t) {
    const long long e = c.ctl[2];
    const double alpha = c.scal[2];
    const double* p = pb_cg_pext(c, c.rank, (int)(e & 1)) + (long long)c.hmax * c.plane;
    c.x[t] = fma(alpha, p[t], c.x[t]);
    const double rv = fma(-alpha, c.Ap[t], c.r[t]);
    c.r[t] = rv;
    return rv * rv;
}
// p_new = z + beta p_old into the buffer of the next epoch; boundary planes also go to the neighbours
PB_HD void pb_cg_direction_elem(const PbCgDev& c, long long t, bool first) {
    const long long e_new = first ? c.ctl[2] : c.ctl[2];       // R2 / the init all-reduce already advanced the epoch
    const int bn = (int)(e_new & 1);
    double v = c.z[t];
    if (!first) {
        const double* pold = pb_cg_pext(c, c.rank, bn ^ 1) + (long long)c.hmax * c.plane;
        v = fma(c.scal[3], pold[t], v);
    }
    const long long off = (long long)c.hmax * c.plane + t;
    pb_cg_pext(c, c.rank, bn)[off] = v;
    const long long pl = t / c.plane, L = c.rb - c.ra;
    if (c.rank > 0 && pl < c.hmax) {                     // my first planes are the upper halo of rank-1
        const long long Lq = c.cuts[c.rank] - c.cuts[c.rank - 1];
        pb_cg_pext(c, c.rank - 1, bn)[((long long)c.hmax + Lq) * c.plane + t] = v;
    }
    if (c.rank + 1 < c.world && pl >= L - c.hmax) {      // my last planes are the lower halo of rank+1
        pb_cg_pext(c, c.rank + 1, bn)[t - (L - c.hmax) * c.plane] = v;
    }
}
// after all stores of the direction kernel: stamp the neighbours' halo flags
PB_HD void pb_cg_direction_signal(const PbCgDev& c) {
    const long long e_new = c.ctl[2];
    PB_CG_FENCE();
    if (c.rank > 0) pb_cg_flag(c, c.rank - 1, c.off_hflag)[1] = e_new;
    if (c.rank + 1 < c.world) pb_cg_flag(c, c.rank + 1, c.off_hflag)[0] = e_new;
}

// one row of y = A p with the value for <p, Ap>
PB_HD double pb_cg_matvec_row(const PbCgDev& c, const PbMlbParams& mp, long long r) {
    const long long e = c.ctl[2];
    const double* pe = pb_cg_pext(c, c.rank, (int)(e & 1));
    pb_mlb_matvec_row(mp, pe, c.ra - c.hmax, c.Ap, r);
    return c.Ap[r] * pe[(long long)c.hmax * c.plane + r];
}
// does the row block [r0, r1) read halo planes from below / above?
PB_HD void pb_cg_block_needs(const PbCgDev& c, long long r0, long long r1, bool& lo, bool& hi) {
    const long long L = c.rb - c.ra;
    lo = c.rank > 0 && r0 / c.plane < c.hmax;
    hi = c.rank + 1 < c.world && (r1 - 1) / c.plane >= L - c.hmax;
}

// mode products of the preconditioner on the local planes: y[a, i, c] = sum_j A[i, j] x[a, j, c]
//   stage 1: t1 = mode 2 (last axis) of r;  stage 2: mode 1 of t1 -> gather buffers of every rank
PB_HD void pb_cg_mode2_elem(const PbCgDev& c, long long t) {
    pb_modek_elem(c.Ainv[2], c.N[2], c.N[2], c.r, 1, c.t1, t);
}
PB_HD void pb_cg_mode1_elem(const PbCgDev& c, long long t) {
    const long long e = c.ctl[2];
    const int b = (int)(e & 1);
    const long long inner = c.N[2];
    const long long cc = t % inner, tt = t / inner;
    const int i = (int)(tt % c.N[1]);
    const long long a = tt / c.N[1];
    const double* xr = c.t1 + a * c.N[1] * inner + cc;
    const double* Ar = c.Ainv[1] + (long long)i * c.N[1];
    double acc = 0.0;
    for (int j = 0; j < c.N[1]; ++j) acc = fma(Ar[j], xr[(long long)j * inner], acc);
    const long long g = (long long)c.ra * c.plane + t;
    for (int q = 0; q < c.world; ++q) pb_cg_gath(c, q, b)[g] = acc;
}
PB_HD void pb_cg_mode1_signal(const PbCgDev& c) {
    const long long e = c.ctl[2];
    PB_CG_FENCE();
    for (int q = 0; q < c.world; ++q) pb_cg_flag(c, q, c.off_gflag)[c.rank] = e;
}
// z = mode 0 of the gathered vector, rows of my slab; value for <r, z>
PB_HD double pb_cg_mode0_elem(const PbCgDev& c, long long t) {
    const long long e = c.ctl[2];
    const double* g = pb_cg_gath(c, c.rank, (int)(e & 1));
    const long long cc = t % c.plane;
    const int i = (int)(t / c.plane) + c.ra;
    const double* Ar = c.Ainv[0] + (long long)i * c.N0;
    double acc = 0.0;
    for (int j = 0; j < c.N0; ++j) acc = fma(Ar[j], g[(long long)j * c.plane + cc], acc);
    c.z[t] = acc;
    return acc * c.r[t];
}

// ---- all-reduce of PB_CG_NRED sums over the ranks and the scalar recurrences ---------------------------
// kind 0: <b,b> -> bnorm2;  1: <r,z> of the first direction -> rz, epoch := epoch + 1;
//      2: <p,Ap> -> alpha;  3: (<r,r>, <r,z>) -> beta, convergence test, next epoch
PB_HD void pb_cg_allreduce_finish(const PbCgDev& c, int kind, const double* local) {
    const long long call = c.ctl[3] + 1;
    const int slot = (int)(call & 1);
    double sum[PB_CG_NRED];
    if (c.world > 1) {
        for (int q = 0; q < c.world; ++q) {
            double* v = reinterpret_cast<double*>(c.win[q] + c.off_rval) + ((long long)slot * c.world + c.rank) * PB_CG_NRED;
            for (int k = 0; k < PB_CG_NRED; ++k) v[k] = local[k];
        }
        PB_CG_FENCE();
        for (int q = 0; q < c.world; ++q) pb_cg_flag(c, q, c.off_rstamp)[(long long)slot * c.world + c.rank] = call;
        for (int k = 0; k < PB_CG_NRED; ++k) sum[k] = 0.0;
        for (int q = 0; q < c.world; ++q) {
            pb_cg_wait(pb_cg_flag(c, c.rank, c.off_rstamp) + (long long)slot * c.world + q, call);
            const volatile double* v = reinterpret_cast<volatile double*>(c.win[c.rank] + c.off_rval) + ((long long)slot * c.world + q) * PB_CG_NRED;
            for (int k = 0; k < PB_CG_NRED; ++k) sum[k] += v[k];
        }
    } else {
        for (int k = 0; k < PB_CG_NRED; ++k) sum[k] = local[k];
    }
    c.ctl[3] = call;
    if (kind == 0) {
        c.scal[5] = sum[0];
        if (sum[0] == 0.0) c.ctl[0] = 1;
    } else if (kind == 1) {
        c.scal[0] = sum[0];
        c.ctl[2] = c.ctl[2] + 1;
    } else if (kind == 2) {
        c.scal[1] = sum[0];
        c.scal[2] = c.scal[0] / sum[0];
    } else {
        c.scal[4] = sum[0];
        c.ctl[1] = c.ctl[1] + 1;
        const double tol = c.scal[6];
        if (sum[0] <= tol * tol * c.scal[5]) {
            c.ctl[0] = 1;
        } else {
            c.scal[3] = sum[1] / c.scal[0];
            c.scal[0] = sum[1];
            c.ctl[2] = c.ctl[2] + 1;
        }
    }
}

#if defined(__CUDACC__) && !defined(PB_EMULATE)
// block sum in a fixed order; result valid in thread 0
__device__ __forceinline__ double pb_cg_block_sum(double v, double* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    return sh[0];
}
__global__ void __launch_bounds__(256) pb_cg_init_kernel(const __grid_constant__ PbCgDev c, const double* b) {
    __shared__ double sh[256];
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const double v = t < c.nloc ? pb_cg_init_elem(c, b, t) : 0.0;
    const double s = pb_cg_block_sum(v, sh);
    if (threadIdx.x == 0) { c.part[blockIdx.x] = s; c.part[gridDim.x + blockIdx.x] = 0.0; }
}
__global__ void __launch_bounds__(256) pb_cg_update_kernel(const __grid_constant__ PbCgDev c) {
    __shared__ double sh[256];
    if (c.ctl[0]) return;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const double v = t < c.nloc ? pb_cg_update_elem(c, t) : 0.0;
    const double s = pb_cg_block_sum(v, sh);
    if (threadIdx.x == 0) c.part[blockIdx.x] = s;
}
// the kernel that finishes last raises the flags: a ticket counter in ctl[4]
__global__ void __launch_bounds__(256) pb_cg_direction_kernel(const __grid_constant__ PbCgDev c, int first) {
    if (c.ctl[0]) return;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t < c.nloc) pb_cg_direction_elem(c, t, first != 0);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(c.ctl + 4), 1ull) + 1;
        if (done == gridDim.x) {
            c.ctl[4] = 0;
            pb_cg_direction_signal(c);
        }
    }
}
__global__ void __launch_bounds__(256) pb_cg_matvec_kernel(const __grid_constant__ PbCgDev c, const __grid_constant__ PbMlbParams mp,
                                                           int nb_lo, int nb_hi) {
    __shared__ double sh[256];
    if (c.ctl[0]) return;
    // blocks that need halo planes come last in the launch order
    const int NB = gridDim.x;
    const int rbk = (int)((blockIdx.x + (unsigned)nb_lo) % (unsigned)NB);
    const long long r0 = (long long)rbk * 256, r1 = r0 + 256 < c.nloc ? r0 + 256 : c.nloc;
    bool lo, hi;
    pb_cg_block_needs(c, r0, r1, lo, hi);
    if (threadIdx.x == 0) {
        const long long e = c.ctl[2];
        if (lo) pb_cg_wait(pb_cg_flag(c, c.rank, c.off_hflag) + 0, e);
        if (hi) pb_cg_wait(pb_cg_flag(c, c.rank, c.off_hflag) + 1, e);
    }
    __syncthreads();
    (void)nb_hi;
    const long long r = r0 + threadIdx.x;
    const double v = r < c.nloc ? pb_cg_matvec_row(c, mp, r) : 0.0;
    const double s = pb_cg_block_sum(v, sh);
    if (threadIdx.x == 0) c.part[rbk] = s;
}
__global__ void __launch_bounds__(256) pb_cg_mode2_kernel(const __grid_constant__ PbCgDev c) {
    if (c.ctl[0]) return;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t < c.nloc) pb_cg_mode2_elem(c, t);
}
__global__ void __launch_bounds__(256) pb_cg_mode1_kernel(const __grid_constant__ PbCgDev c) {
    if (c.ctl[0]) return;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t < c.nloc) pb_cg_mode1_elem(c, t);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(c.ctl + 5), 1ull) + 1;
        if (done == gridDim.x) {
            c.ctl[5] = 0;
            pb_cg_mode1_signal(c);
        }
    }
}
__global__ void __launch_bounds__(256) pb_cg_mode0_kernel(const __grid_constant__ PbCgDev c, int second) {
    __shared__ double sh[256];
    if (c.ctl[0]) return;
    if (threadIdx.x == 0 && c.world > 1) {
        const long long e = c.ctl[2];
        for (int q = 0; q < c.world; ++q) pb_cg_wait(pb_cg_flag(c, c.rank, c.off_gflag) + q, e);
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const double v = t < c.nloc ? pb_cg_mode0_elem(c, t) : 0.0;
    const double s = pb_cg_block_sum(v, sh);
    if (threadIdx.x == 0) c.part[(second ? gridDim.x : 0) + blockIdx.x] = s;
}
// ---- tiled mode products: 8 outputs per thread, the factor rows / vector rows staged in shared memory --------
// mode 2 (last axis, contiguous): t1[a, i] = sum_j r[a, j] A2[i, j]; block: PB_CG_TI rows a, threads over i
__global__ void __launch_bounds__(128) pb_cg_mode2_tiled_kernel(const __grid_constant__ PbCgDev c) {
    extern __shared__ double pb_cg_sm[];
    if (c.ctl[0]) return;
    const int n = c.N[2];
    const long long rows = c.nloc / n;
    const long long a0 = (long long)blockIdx.x * PB_CG_TI;
    for (int t = threadIdx.x; t < PB_CG_TI * n; t += blockDim.x) {
        const long long a = a0 + t / n;
        pb_cg_sm[t] = a < rows ? c.r[a * n + t % n] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc[PB_CG_TI];
#pragma unroll
        for (int k = 0; k < PB_CG_TI; ++k) acc[k] = 0.0;
        for (int j = 0; j < n; ++j) {
            const double av = c.AinvT2[(long long)j * n + i];
#pragma unroll
            for (int k = 0; k < PB_CG_TI; ++k) acc[k] = fma(av, pb_cg_sm[k * n + j], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < PB_CG_TI; ++k)
            if (a0 + k < rows) c.t1[(a0 + k) * n + i] = acc[k];
    }
}
// modes 1 and 0: y[a, i, cc] = sum_j A[i, j] x[a, j, cc], cc contiguous; block: PB_CG_TI outputs i, 128 values of cc
//   MODE 1: x = t1 (local planes a), result -> the gather buffers of all ranks
//   MODE 0: x = the gathered vector (one "a"), rows i of my slab, result -> z and the partial sums of <r, z>
template <int MODE>
__global__ void __launch_bounds__(128) pb_cg_modek_tiled_kernel(const __grid_constant__ PbCgDev c, int second) {
    extern __shared__ double pb_cg_sm[];
    __shared__ double sh[128];
    if (c.ctl[0]) return;
    const long long e = c.ctl[2];
    const int b = (int)(e & 1);
    const int n = MODE == 1 ? c.N[1] : c.N0;
    const long long inner = MODE == 1 ? c.N[2] : c.plane;
    const int mrows = MODE == 1 ? c.N[1] : (c.rb - c.ra);
    const int i0 = blockIdx.y * PB_CG_TI;
    const long long a = MODE == 1 ? blockIdx.z : 0;
    const double* A = MODE == 1 ? c.Ainv[1] : c.Ainv[0] + (long long)c.ra * c.N0;
    if (MODE == 0 && threadIdx.x == 0 && c.world > 1)
        for (int q = 0; q < c.world; ++q) pb_cg_wait(pb_cg_flag(c, c.rank, c.off_gflag) + q, e);
    for (int t = threadIdx.x; t < PB_CG_TI * n; t += blockDim.x) {
        const int ii = i0 + t / n;
        pb_cg_sm[t] = ii < mrows ? A[(long long)ii * n + t % n] : 0.0;
    }
    __syncthreads();
    const long long cc = (long long)blockIdx.x * 128 + threadIdx.x;
    double dot = 0.0;
    if (cc < inner) {
        const double* x = MODE == 1 ? c.t1 + a * n * inner + cc : pb_cg_gath(c, c.rank, b) + cc;
        double acc[PB_CG_TI];
#pragma unroll
        for (int k = 0; k < PB_CG_TI; ++k) acc[k] = 0.0;
        for (int j = 0; j < n; ++j) {
            const double xv = x[(long long)j * inner];
#pragma unroll
            for (int k = 0; k < PB_CG_TI; ++k) acc[k] = fma(pb_cg_sm[k * n + j], xv, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < PB_CG_TI; ++k) {
            if (i0 + k >= mrows) continue;
            if (MODE == 1) {
                const long long g = (long long)c.ra * c.plane + (a * n + i0 + k) * inner + cc;
                for (int q = 0; q < c.world; ++q) pb_cg_gath(c, q, b)[g] = acc[k];
            } else {
                const long long t = (long long)(i0 + k) * inner + cc;
                c.z[t] = acc[k];
                dot = fma(acc[k], c.r[t], dot);
            }
        }
    }
    if (MODE == 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
            const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(c.ctl + 5), 1ull) + 1;
            if (done == total) {
                c.ctl[5] = 0;
                pb_cg_mode1_signal(c);
            }
        }
    } else {
        sh[threadIdx.x] = dot;
        __syncthreads();
        for (int s2 = 64; s2 > 0; s2 >>= 1) {
            if ((int)threadIdx.x < s2) sh[threadIdx.x] += sh[threadIdx.x + s2];
            __syncthreads();
        }
        if (threadIdx.x == 0) c.part2[(long long)blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
        (void)second;
    }
}

__global__ void pb_cg_bump_epoch_kernel(const __grid_constant__ PbCgDev c) { c.ctl[2] = c.ctl[2] + 1; }
__global__ void __launch_bounds__(256) pb_cg_allreduce_kernel(const __grid_constant__ PbCgDev c, int kind, int nblocks) {
    __shared__ double sh[256];
    if (c.ctl[0] && kind >= 2) return;
    double loc[PB_CG_NRED];
    for (int k = 0; k < PB_CG_NRED; ++k) {
        // <r, z> (kind 1: value 0, kind 3: value 1) is summed by the tiled mode-0 kernel into part2
        const bool rz = (kind == 1 && k == 0) || (kind == 3 && k == 1);
        const double* src = rz ? c.part2 : c.part + (long long)k * nblocks;
        const int cnt = rz ? c.nblocks2 : nblocks;
        double v = 0.0;
        for (int i = threadIdx.x; i < cnt; i += 256) v += src[i];
        loc[k] = pb_cg_block_sum(v, sh);
        __syncthreads();
    }
    if (threadIdx.x == 0) pb_cg_allreduce_finish(c, kind, loc);
}
#endif
