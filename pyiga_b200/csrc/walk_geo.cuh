// Fused stage 1 of the 3D pipeline: geometry + coefficient fields + contraction of axis 0.
//
// The unfused pipeline runs K2 (geo_fields.cuh) to write the coefficient fields F[c][g0][g1][g2]
// to HBM and two walk kernels to read them back.  Here the walk along axis 0 evaluates the
// fields itself: a thread owns the line (g1, g2); it first reduces the control net of the
// geometry over axes 1 and 2 at its (g1, g2),
//     Z[i0][c][v] = sum_{a1,a2} coeffs[i0][f1+a1][f2+a2][c] * w_v(a1, a2)
// (v = 0: values, 1: derivative on axis 1, 2: derivative on axis 2) — a 1D spline curve in g0 per
// component — and then, node by node, finishes value and Jacobian with p_g0+1 terms, applies the
// quotient rule (NURBS), runs the field program of the form (PbProgStiffness / PbProgMass) in
// registers and feeds the result to the walk as its inputs.  No field ever reaches HBM:
// 13 GB less traffic per step for 3D p=3 n=128 stiffness.  This is the sum-factorised form of
// `grid_jacobian` (pyiga/bspline.py:897-921, geometry.py:17-25,116-123) followed by
// `precompute_fields` (pyiga/assemblers.pyx:1389-1449) for one grid line.
//
// Z lives in shared memory, one column per thread; the kernel is used when the whole curve fits
// (Ng0 * NC * 3 <= PB_GEO_ZMAX doubles per thread), i.e. for geometries with a short control net
// along axis 0 — all stock geometries.  Otherwise the unfused pipeline runs.
#pragma once
#include "geo_fields.cuh"
#include "walk.cuh"

#define PB_GEO_ZMAX 60

// Thread-private columns in shared memory are laid out in PAIRS: element e of the column of a thread
// sits at ((e >> 1) * stride) * 2 + (e & 1), stride counted in pairs (128 on the device: the pairs of
// the 128 threads of a block are contiguous, so that a 16-byte access per thread is conflict-free;
// 1 in the sequential emulation).  Even / odd neighbours are read with one 128-bit load.
PB_HD long long pb_col(int e, int stride) { return (long long)(e >> 1) * 2 * stride + (e & 1); }
PB_HD double2 pb_col_pair(const double* col, int e_even, int stride) {
    return *reinterpret_cast<const double2*>(col + (long long)(e_even >> 1) * 2 * stride);
}

struct PbGeoLineParams {
    PbGeoDev geo;
    const double* gw[PB_MAXDIM];    // Gauss weights per axis
    int G1, G2;                     // nodes on axes 1 and 2 (the line index is g1 * G2 + g2)
};

template <class Plan, int Q, int NC, class Prog>
struct PbGeoLoader {
    static constexpr int NOPS = Plan::NOPS;
    static constexpr int ZI = (NC * 3 + 1) & ~1;        // doubles per control point of the reduced net (padded to pairs)
    static constexpr int NOPSP = (NOPS + 1) & ~1;       // staged fields per node (padded to pairs)
    static constexpr bool ROLLED = true;    // the walk asks for the inputs node by node (compact code)
    // members the walk fills for memory loaders; unused here
    const double* src[NOPS];
    bool has[NOPS];
    long long sc;
    int s_end;
    // geometry of the line
    const double* Z;        // element i0 * ZI + c * 3 + v of the thread's column (pb_col)
    int zs;
    int pg0;
    const double* T0;       // [g0][2][pg0+1]   indexed by absolute node (possibly a staged slice)
    const int* F0;          // [g0] first active geometry function
    const double* W0;       // [g0] Gauss weights of axis 0
    double gw12;
    double* F;              // staged fields of the current span: F[(gq * NOPS + i) * fs], private to the thread
    int fs;

    // reduce the control net over axes 1 and 2 at (g1, g2) into Zbuf (stride zstride)
    PB_HD void init(const PbGeoLineParams& gp, int x, double* Zbuf, int zstride) {
        const PbGeoDev& geo = gp.geo;
        const int g1 = x / gp.G2, g2 = x % gp.G2;
        gw12 = gp.gw[1][g1] * gp.gw[2][g2];
        Z = Zbuf; zs = zstride; pg0 = geo.pg[0];
        T0 = geo.GV[0]; F0 = geo.gfirst[0]; W0 = gp.gw[0];
        const int pg1 = geo.pg[1], pg2 = geo.pg[2];
        const int f1 = geo.gfirst[1][g1], f2 = geo.gfirst[2][g2];
        const double* T1 = geo.GV[1] + (long long)g1 * 2 * (pg1 + 1);
        const double* T2 = geo.GV[2] + (long long)g2 * 2 * (pg2 + 1);
        for (int i0 = 0; i0 < geo.Ng[0]; ++i0) {
            double s0[NC], s1[NC], s2[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) s0[c] = s1[c] = s2[c] = 0.0;
            for (int a1 = 0; a1 <= pg1; ++a1) {
                double t0[NC], t2[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) t0[c] = t2[c] = 0.0;
                const double* cf = geo.coeffs + (((long long)i0 * geo.Ng[1] + (f1 + a1)) * geo.Ng[2] + f2) * NC;
                for (int a2 = 0; a2 <= pg2; ++a2) {
                    const double w2 = T2[a2], d2 = T2[pg2 + 1 + a2];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const double cv = cf[a2 * NC + c];
                        t0[c] = fma(cv, w2, t0[c]);
                        t2[c] = fma(cv, d2, t2[c]);
                    }
                }
                const double w1 = T1[a1], d1 = T1[pg1 + 1 + a1];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    s0[c] = fma(w1, t0[c], s0[c]);
                    s1[c] = fma(d1, t0[c], s1[c]);
                    s2[c] = fma(w1, t2[c], s2[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                Zbuf[pb_col(i0 * ZI + c * 3 + 0, zstride)] = s0[c];
                Zbuf[pb_col(i0 * ZI + c * 3 + 1, zstride)] = s1[c];
                Zbuf[pb_col(i0 * ZI + c * 3 + 2, zstride)] = s2[c];
            }
        }
    }

    PB_HD void prime(int) {}
    PB_HD void next(int, double (&)[Q][NOPS]) {}        // unrolled interface of the walk: not used (ROLLED)
    template <int I> PB_HD double get(int gq) const { return F[pb_col(gq * NOPSP + I, fs)]; }
    // all inputs of node gq with 128-bit loads
    PB_HD void get_node(int gq, double (&x)[NOPS]) const {
#pragma unroll
        for (int k = 0; k < NOPSP / 2; ++k) {
            const double2 v = pb_col_pair(F, gq * NOPSP + 2 * k, fs);
            x[2 * k] = v.x;
            if (2 * k + 1 < NOPS) x[2 * k + 1] = v.y;
        }
    }
    // the same column stages finished window entries between the phase-specific code and the store loop
    PB_HD void stage_put(int e, double v) { F[pb_col(e, fs)] = v; }
    PB_HD double stage_get(int e) const { return F[pb_col(e, fs)]; }
    static constexpr int stage_doubles(int P) {
        return ((Q * NOPSP > (P + 1) * Plan::NOUT ? Q * NOPSP : (P + 1) * Plan::NOUT) + 1) & ~1;
    }

    // evaluate the fields of the Q nodes of span s into the staging column
    PB_HD void begin_span(int s) {
        constexpr int GD = 3;
        constexpr bool RAT = (NC == GD + 1);
#pragma unroll (Plan::NOUT > 1 ? 2 : 4)
        for (int gq = 0; gq < Q; ++gq) {
            const int g0 = s * Q + gq;
            const int f0 = F0[g0];
            const double* Tn = T0 + (long long)g0 * 2 * (pg0 + 1);
            double val[NC], dv[NC][3];
            for (int a = 0; a <= pg0; ++a) {
                const double w = Tn[a], d = Tn[pg0 + 1 + a];
                double z[ZI];
#pragma unroll
                for (int k = 0; k < ZI / 2; ++k) {          // the control point's ZI values, 128 bits at a time
                    const double2 v = pb_col_pair(Z, (f0 + a) * ZI + 2 * k, zs);
                    z[2 * k] = v.x;
                    z[2 * k + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const double z0 = z[c * 3 + 0], z1 = z[c * 3 + 1], z2 = z[c * 3 + 2];
                    if (a == 0) {       // first active function assigns (no zero-fill of the accumulators)
                        val[c] = w * z0;
                        dv[c][0] = d * z0;          // derivative along tensor axis 0
                        dv[c][1] = w * z1;
                        dv[c][2] = w * z2;
                    } else {
                        val[c] = fma(w, z0, val[c]);
                        dv[c][0] = fma(d, z0, dv[c][0]);
                        dv[c][1] = fma(w, z1, dv[c][1]);
                        dv[c][2] = fma(w, z2, dv[c][2]);
                    }
                }
            }
            PbPoint pt;
            pt.idx = 0;
            pt.gw = W0[g0] * gw12;
            if constexpr (RAT) {    // quotient rule without the division; the program folds 1 / W^2 in
                const double W = val[GD];
                pt.jden = W * W;
#pragma unroll
                for (int i = 0; i < GD; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k) pt.J[i][2 - k] = dv[i][k] * W - val[i] * dv[GD][k];
            } else {
                pt.jden = 1.0;
#pragma unroll
                for (int i = 0; i < GD; ++i)
#pragma unroll
                    for (int k = 0; k < 3; ++k) pt.J[i][2 - k] = dv[i][k];
            }
            double f[Prog::NF];
            Prog::template point<RAT, true>(pt, f);
            pb_static_for<0, NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                F[pb_col(gq * NOPSP + i, fs)] = f[Plan::field(i)];
            });
        }
    }
};

// sequential emulation of one line
template <class Plan, int P, int Q, int NC, class Prog>
PB_HD void pb_walk_geo_line(const PbWalkParams& prm, const PbGeoLineParams& gp, long long tid, int piece) {
    alignas(16) double Zloc[PB_GEO_ZMAX + 8];
    alignas(16) double Floc[PbGeoLoader<Plan, Q, NC, Prog>::stage_doubles(P)];
    PbGeoLoader<Plan, Q, NC, Prog> ld;
    ld.init(gp, (int)(tid % prm.X), Zloc, 1);
    ld.F = Floc; ld.fs = 1;
    PbWalkTables tb;
    tb.V = prm.V2; tb.first = prm.first; tb.ret_mu = prm.ret_mu;
    const PbWalkRange rg = pb_walk_range(prm, piece);
    pb_walk_line_impl<Plan, P, Q, false, true>(prm, rg, tid, tb, ld);
}

#if defined(__CUDACC__)
// shared memory: [ V slice | first + retire tables | geometry axis-0 slice (T0, weights, first) | Z columns ]
template <int P, int Q>
PB_HD void pb_walk_geo_smem(const PbWalkRange& rg, int pg0, int nz, size_t& vbytes, size_t& ibytes, size_t& gbytes, size_t& zbytes) {
    // nz: doubles per thread of the Z column plus the field staging column
    const size_t nodes = (size_t)(rg.s_end - rg.s_begin) * Q;
    const int nsp = rg.s_end - rg.s_begin;
    vbytes = (nodes * 2 * (P + 1) * sizeof(double) + 127) & ~size_t(127);
    ibytes = (((size_t)((nsp + 3) & ~3) + (size_t)(rg.f_hi - rg.f_lo) * (2 * P + 1)) * sizeof(int) + 127) & ~size_t(127);
    gbytes = (nodes * (2 * (pg0 + 1) + 1) * sizeof(double) + nodes * sizeof(int) + 127) & ~size_t(127);
    zbytes = (size_t)nz * 128 * sizeof(double);
}

template <class Plan, int P, int Q, int NC, class Prog, int MINB>
__global__ void __launch_bounds__(128, MINB) pb_walk_geo_kernel(const __grid_constant__ PbWalkParams prm,
                                                                const __grid_constant__ PbGeoLineParams gp) {
    extern __shared__ __align__(128) unsigned char pb_smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const PbWalkRange rg = pb_walk_range(prm, blockIdx.y);
    const int pg0 = gp.geo.pg[0];
    const int nz = gp.geo.Ng[0] * PbGeoLoader<Plan, Q, NC, Prog>::ZI + PbGeoLoader<Plan, Q, NC, Prog>::stage_doubles(P);
    size_t vbytes, ibytes, gbytes, zbytes;
    pb_walk_geo_smem<P, Q>(rg, pg0, nz, vbytes, ibytes, gbytes, zbytes);
    const long long first_node = (long long)rg.s_begin * Q;
    const int nnodes = (rg.s_end - rg.s_begin) * Q;
    // ---- basis table slice of the walk axis: TMA bulk copy --------------------------------------
    double* sV = reinterpret_cast<double*>(pb_smem_raw);
    {
        const uint32_t bytes = (uint32_t)((long long)nnodes * 2 * (P + 1) * sizeof(double));
        if (threadIdx.x == 0) pb_mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const char* g = reinterpret_cast<const char*>(prm.V2 + first_node * 2 * (P + 1));
            char* d = reinterpret_cast<char*>(sV);
            uint32_t done = 0;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_u32(&bar)), "r"(bytes) : "memory");
            while (done < bytes) {
                const uint32_t piece = (bytes - done) < 32768u ? (bytes - done) : 32768u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 pb_smem_u32(d + done)),
                             "l"(g + done), "r"(piece), "r"(pb_smem_u32(&bar))
                             : "memory");
                done += piece;
            }
        }
    }
    // ---- integer tables with the slab filter folded in (as in pb_walk_kernel) ---------------------
    const int nsp = rg.s_end - rg.s_begin;
    const int f_lo = rg.f_lo, f_hi = rg.f_hi;
    int* s_first = reinterpret_cast<int*>(pb_smem_raw + vbytes);
    int* s_ret = s_first + ((nsp + 3) & ~3);
    for (int t = threadIdx.x; t < nsp; t += blockDim.x) s_first[t] = prm.first[rg.s_begin + t];
    for (int t = threadIdx.x; t < (f_hi - f_lo) * (2 * P + 1); t += blockDim.x) {
        int mu = prm.ret_mu[(long long)f_lo * (2 * P + 1) + t];
        if (mu >= 0) {
            const int f = f_lo + t / (2 * P + 1), k = t % (2 * P + 1);
            const int i = (k <= P) ? f : f + (k - P), j = (k <= P) ? f + k : f;
            int mask = 0;
            for (int o = 0; o < Plan::NOUT; ++o) mask |= pb_walk_keep(prm, rg, o, i, j) ? (1 << o) : 0;
            mu = mask ? (mu | (mask << 24)) : -1;
        }
        s_ret[t] = mu;
    }
    // ---- geometry tables of axis 0 on the walked nodes --------------------------------------------
    double* sT0 = reinterpret_cast<double*>(pb_smem_raw + vbytes + ibytes);
    double* sW0 = sT0 + (size_t)nnodes * 2 * (pg0 + 1);
    int* sF0 = reinterpret_cast<int*>(sW0 + nnodes);
    for (int t = threadIdx.x; t < nnodes * 2 * (pg0 + 1); t += blockDim.x) sT0[t] = gp.geo.GV[0][first_node * 2 * (pg0 + 1) + t];
    for (int t = threadIdx.x; t < nnodes; t += blockDim.x) {
        sW0[t] = gp.gw[0][first_node + t];
        sF0[t] = gp.geo.gfirst[0][first_node + t];
    }
    double* sZ = reinterpret_cast<double*>(pb_smem_raw + vbytes + ibytes + gbytes);
    pb_mbar_wait(&bar, 0);
    __syncthreads();

    PbWalkTables tb;
    tb.V = sV - first_node * 2 * (P + 1);
    tb.first = s_first - rg.s_begin;
    tb.ret_mu = s_ret - (long long)f_lo * (2 * P + 1);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < prm.nthreads) {
        PbGeoLoader<Plan, Q, NC, Prog> ld;
        ld.init(gp, (int)(tid % prm.X), sZ + 2 * threadIdx.x, 128);
        ld.F = sZ + (size_t)gp.geo.Ng[0] * PbGeoLoader<Plan, Q, NC, Prog>::ZI * 128 + 2 * threadIdx.x;
        ld.fs = 128;
        ld.T0 = sT0 - first_node * 2 * (pg0 + 1);
        ld.W0 = sW0 - first_node;
        ld.F0 = sF0 - first_node;
        pb_walk_line_impl<Plan, P, Q, true, true>(prm, rg, tid, tb, ld);
    }
}
#endif
