// Fused stages 2 + 3 of the 3D pipeline ("S32"): X1[t][mu0][g1][g2] -> data[mu0][mu1][mu2].
//
// The unfused pipeline contracts axis 1 into X2[t][mu0][mu1][g2] (HBM, 7 GB written and 9 GB read
// back for 3D p=3 n=128 stiffness) and then axis 2.  Here one thread block owns one band entry mu0
// of the first axis and a batch of 32 spans of the LAST axis, and contracts in the other order:
//
//   phase A (producer warps, one grid row g1 each): the lane-per-span contraction of axis 2 of
//       walk.cuh (`pb_lane_span_kernel_v2`): lane = span, inputs copied global -> shared with
//       cp.async several rows ahead, local (p+1)^2 blocks built in registers, neighbouring blocks
//       summed with warp shuffles.  It yields, for this g1, the terms
//           T[t'][mu2]   (t' = remaining derivative slots on axis 1)
//       in shared memory — never in HBM.
//   phase B (consumer warps, one thread per band entry mu2 of the batch): the rotating-window walk
//       along axis 1 over the T rows, (p+1)^2 accumulators per thread, finished entries
//       data[mu0][mu1][mu2] stored coalesced along mu2.
//
// Producers work one g1 span ahead of the consumers (double-buffered T, one __syncthreads per
// span).  Symmetric forms compute only the "upper" mu0 (i0 <= j0, or partner row outside the
// slab) and write every result also at the transposed position (mu0^T, mu1^T, mu2^T).
//
// Same sums as `combine` / `entry_impl` of the reference (pyiga/assemblers.pyx:1455-1540), evaluated
// by sum factorisation; terms and flags of the stiffness form as in plans.cuh.
#pragma once
#include "plans.cuh"

struct PbPlanPairU {    // [0,0] + [0,1] -> one output
    static constexpr int NOPS = 2, NOUT = 1, MINB = 4, NPF = 3;
    static constexpr int MINB4 = 4;
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[2] = {{0, 0, 0, 0, 0}, {1, 0, 0, 1, 0}};
        return t[i];
    }
};

#define PB_S32_MAXPIECE 8
struct PbS32Params {
    // ---- input: X1 terms, plane (mu0 - x1_mu_base) of term t at X1 + t * x1_stride ----------------
    const double* X1;
    long long x1_stride;
    int x1_mu_base;
    // ---- axis 0: band entries visited (the rows of the slab) and the slab itself ------------------
    int mu0_begin, mu0_count;
    int u_lo, u_hi;
    const int* pair_i0;
    const int* pair_j0;
    const int* tr0;
    int symmetric;              // 1: compute upper mu0 only and mirror; 0: every mu0 on its own
    // ---- axis 1 (walked by the consumers) and axis 2 (lane-per-span) ------------------------------
    int G1, G2, n1, n2, N1, N2, M1, M2;
    const int* first1; const double* V1; const int* ret_mu1; const int* tr1; const int* pair_i1;
    const int* first2; const double* V2; const int* ret_mu2; const int* tr2;
    // ---- output: data[(mu0 - out_mu_base)][mu1][mu2] -----------------------------------------------
    double* out;
    int out_mu_base;
    int nbatch;
    // ---- optional split of axis 1 into pieces (more, shorter blocks: fills the tail wave of small slabs) --
    // piece y retires the pairs (i1, j1) whose row i1 lies in [pw_lo[y], pw_hi[y]) and walks only the spans
    // those rows see; the pieces overlap by p spans (same scheme as the walk-axis pieces of walk.cuh)
    int npiece;
    int ps_begin[PB_S32_MAXPIECE], ps_end[PB_S32_MAXPIECE], pw_lo[PB_S32_MAXPIECE], pw_hi[PB_S32_MAXPIECE];
    // ---- generic forms (PbS32Generic): X1 term read by input stream i, or -1 when the form has no such term ----
    int in_slot[9];
    // ---- tasks of the CUDA kernel: the band entries mu0 this launch computes (pb_s32_keep), in order ----------
    // Full batches are one block each.  The LAST batch of an axis rarely fills the warp (n2 = 128, p = 3: 15 of
    // 32 lanes); when tail_k >= 2 such tails of tail_k consecutive kept entries share one block, tail_w lanes each.
    const int* keep;
    int nkeep, tail_k, tail_w;
    // tasks [0, n_whole) run unsplit even when npiece > 1: full waves of whole tasks, only the remainder in pieces
    int n_whole;
    // nodes of axis 1 whose basis table a block stages in shared memory: G1, or the longest piece when all tasks are
    // cut (long axes: the table of the whole axis would not fit)
    int v1_rows;
};
// blocks of a launch (per axis-1 piece)
PB_HD long long pb_s32_tasks(const PbS32Params& prm) {
    if (prm.tail_k > 1) return (long long)prm.nkeep * (prm.nbatch - 1) + (prm.nkeep + prm.tail_k - 1) / prm.tail_k;
    return (long long)prm.nkeep * prm.nbatch;
}
struct PbS32Piece { int s_begin, s_end, w_lo, w_hi; };
PB_HD PbS32Piece pb_s32_piece(const PbS32Params& prm, int y) {
    PbS32Piece p;
    if (prm.npiece > 1 && y >= 0) { p.s_begin = prm.ps_begin[y]; p.s_end = prm.ps_end[y]; p.w_lo = prm.pw_lo[y]; p.w_hi = prm.pw_hi[y]; }
    else { p.s_begin = 0; p.s_end = prm.n1; p.w_lo = 0; p.w_hi = 0x7fffffff; }
    return p;
}

// ---- forms ------------------------------------------------------------------------------------------
// Stiffness.  X1 terms (plans.cuh): 0 (v,v)  1 (v,d1)  2 (v,d2)  3 (d1,d1)  4 (d1,d2)  5 (d2,d2);
// (d1,v)[mu0] = (v,d1)[mu0^T], (d2,v)[mu0] = (v,d2)[mu0^T], (d2,d1) = (d1,d2).
// Phase A contracts axis 2 (flags = derivative on axis 2 of the test / trial slot):
//   T0 = (v,v)'   = [0,0](v,v) + [0,1](v,d2) + [1,0](d2,v) + [1,1](d2,d2)
//   T1 = (v,d1)'  = [0,0](v,d1) + [1,0](d2,d1)
//   T2 = (d1,v)'  = [0,0](d1,v) + [0,1](d1,d2)
//   T3 = (d1,d1)' = [0,0](d1,d1)
// Phase B contracts axis 1:  K = [0,0]T0 + [0,1]T1 + [1,0]T2 + [1,1]T3.
template <int T> struct PbS32StiffTerm;
template <> struct PbS32StiffTerm<0> { using Plan = PbPlanGen4; static constexpr int stream(int i) { return i; } };
template <> struct PbS32StiffTerm<1> { using Plan = PbPlanPairT; static constexpr int stream(int i) { return i == 0 ? 4 : 5; } };
template <> struct PbS32StiffTerm<2> { using Plan = PbPlanPairU; static constexpr int stream(int i) { return i == 0 ? 6 : 5; } };
template <> struct PbS32StiffTerm<3> { using Plan = PbPlanCopy; static constexpr int stream(int) { return 7; } };
struct PbS32Stiffness {
    static constexpr int NIN = 8, NT = 4;
    static constexpr bool RUNTIME = false;
    static constexpr int in_term(int i) {       // X1 term read by input stream i
        constexpr int t[8] = {0, 2, 2, 5, 1, 4, 1, 3};
        return t[i];
    }
    static constexpr bool in_tr(int i) { return i == 2 || i == 6; }     // read at the transposed mu0
    template <int T> using Term = PbS32StiffTerm<T>;
    using PlanB = PbPlanGen4;
};
template <int T> struct PbS32MassTerm { using Plan = PbPlanCopy; static constexpr int stream(int) { return 0; } };
struct PbS32Mass {
    static constexpr int NIN = 1, NT = 1;
    static constexpr bool RUNTIME = false;
    static constexpr int in_term(int) { return 0; }
    static constexpr bool in_tr(int) { return false; }
    template <int T> using Term = PbS32MassTerm<T>;
    using PlanB = PbPlanCopy;
};

// General scalar forms with at most first derivatives (no symmetry assumed: every band entry mu0 is
// computed, nothing is mirrored).  After axis 0 an X1 term is named by its remaining slots
// (test, trial) in {v, d1, d2}; input stream 3 * test + trial reads it (prm.in_slot: its position in the
// X1 buffer, -1 if the form has no such term).  Terms and flags as for stiffness, with nine distinct inputs:
//   T0 = [0,0](v,v) + [0,1](v,d2) + [1,0](d2,v) + [1,1](d2,d2)      T1 = [0,0](v,d1) + [1,0](d2,d1)
//   T2 = [0,0](d1,v) + [0,1](d1,d2)                                 T3 = [0,0](d1,d1)
template <int T> struct PbS32GenTerm;
template <> struct PbS32GenTerm<0> { using Plan = PbPlanGen4; static constexpr int stream(int i) { constexpr int s[4] = {0, 2, 6, 8}; return s[i]; } };
template <> struct PbS32GenTerm<1> { using Plan = PbPlanPairT; static constexpr int stream(int i) { return i == 0 ? 1 : 7; } };
template <> struct PbS32GenTerm<2> { using Plan = PbPlanPairU; static constexpr int stream(int i) { return i == 0 ? 3 : 5; } };
template <> struct PbS32GenTerm<3> { using Plan = PbPlanCopy; static constexpr int stream(int) { return 4; } };
struct PbS32Generic {
    static constexpr int NIN = 9, NT = 4;
    static constexpr bool RUNTIME = true;
    static constexpr int in_term(int i) { return i; }
    static constexpr bool in_tr(int) { return false; }
    template <int T> using Term = PbS32GenTerm<T>;
    using PlanB = PbPlanGen4;
};
// X1 term of input stream i (-1: absent, reads as zero)
template <class Form> PB_HD int pb_s32_slot(const PbS32Params& prm, int i) { return Form::RUNTIME ? prm.in_slot[i] : Form::in_term(i); }

template <int P> struct PbS32Cfg {
    static constexpr int TPAD = ((32 + P) * (2 * P + 1) + 31) / 32 * 32;    // band positions of a batch, padded
    static constexpr int NCW = TPAD / 32;                                    // consumer warps
};

// is the band entry mu0 computed by this launch, and is its result mirrored?
PB_HD bool pb_s32_keep(const PbS32Params& prm, int mu0, bool& mirror) {
    mirror = false;
    if (!prm.symmetric) return true;
    const int i0 = prm.pair_i0[mu0], j0 = prm.pair_j0[mu0];
    const bool j_in = j0 >= prm.u_lo && j0 < prm.u_hi;
    if (j_in && j0 < i0) return false;      // lower pair whose partner row is owned: written by the upper one
    mirror = j_in && j0 > i0;
    return true;
}

// ---- sequential emulation of one (mu0, batch) block (host build; same plans and tables) -------------
template <class Form, int P, int Q>
PB_HD void pb_s32_seq(const PbS32Params& prm, int mu0, int batch, int piece, double* Trow /* [G1][NT][TPAD] scratch */) {
    const PbS32Piece pc = pb_s32_piece(prm, piece);
    constexpr int P1 = P + 1, NT = Form::NT, TPAD = PbS32Cfg<P>::TPAD;
    bool mirror;
    if (!pb_s32_keep(prm, mu0, mirror)) return;
    const int mu0t = prm.symmetric ? prm.tr0[mu0] : mu0;
    const long long plane = (long long)prm.G1 * prm.G2;
    const int sb = batch * (32 - P);
    const int f0 = prm.first2[0];
    // band positions written by this batch
    int mu_lo = 0x7fffffff, mu_hi = -1;
    for (int lane = 0; lane < 32; ++lane) {
        const int m = f0 + sb + lane;
        if (!((batch == 0 || lane >= P) && m < prm.N2)) continue;
        for (int k = 0; k <= 2 * P; ++k) {
            const int mu = prm.ret_mu2[(long long)m * (2 * P + 1) + k];
            if (mu >= 0) { mu_lo = pb_min(mu_lo, mu); mu_hi = pb_max(mu_hi, mu); }
        }
    }
    if (mu_hi < 0) return;
    bool act[TPAD];
    for (int e = 0; e < TPAD; ++e) act[e] = false;
    // phase A for every row
    for (int r = pc.s_begin * Q; r < pc.s_end * Q; ++r) {
        double Ls[NT][32][P1][P1];
        for (int lane = 0; lane < 32; ++lane) {
            const int s = sb + lane;
            double D[Q][2][P1];
            for (int gq = 0; gq < Q; ++gq)
                for (int a = 0; a < P1; ++a) {
                    const double* Vn = prm.V2 + (long long)((s < prm.n2 ? s : 0) * Q + gq) * 2 * P1;
                    D[gq][0][a] = s < prm.n2 ? Vn[a] : 0.0;
                    D[gq][1][a] = s < prm.n2 ? Vn[P1 + a] : 0.0;
                }
            pb_static_for<0, NT>([&](auto TT) {
                constexpr int t = decltype(TT)::value;
                using TP = typename Form::template Term<t>::Plan;
                double xt[Q][TP::NOPS];
                for (int gq = 0; gq < Q; ++gq)
                    for (int i = 0; i < TP::NOPS; ++i) {
                        const int st = Form::template Term<t>::stream(i);
                        const int slot = pb_s32_slot<Form>(prm, st);
                        const double* src = prm.X1 + (long long)slot * prm.x1_stride
                                            + (long long)((Form::in_tr(st) ? mu0t : mu0) - prm.x1_mu_base) * plane + (long long)r * prm.G2;
                        xt[gq][i] = (s < prm.n2 && slot >= 0) ? src[(long long)s * Q + gq] : 0.0;
                    }
                pb_span_block<TP, P, Q>(xt, D, Ls[t][lane]);
            });
        }
        for (int lane = 0; lane < 32; ++lane) {
            const int m = f0 + sb + lane;
            if (!((batch == 0 || lane >= P) && m < prm.N2)) continue;
            const int* rm = prm.ret_mu2 + (long long)m * (2 * P + 1);
            for (int k = 0; k <= 2 * P; ++k) {
                if (rm[k] < 0) continue;
                const int d = (k <= P) ? k : k - P;
                for (int t = 0; t < NT; ++t) {
                    double sum = 0.0;
                    for (int q = 0; q <= P - d; ++q)
                        if (lane - q >= 0) sum += (k <= P) ? Ls[t][lane - q][q][q + d] : Ls[t][lane - q][q + d][q];
                    Trow[((long long)r * NT + t) * TPAD + (rm[k] - mu_lo)] = sum;
                }
                act[rm[k] - mu_lo] = true;
            }
        }
    }
    // phase B: walk along axis 1 for every owned position
    using PB = typename Form::PlanB;
    for (int e = 0; e < TPAD; ++e) {
        if (!act[e]) continue;
        PbWalkParams w;
        memset(&w, 0, sizeof w);
        w.nthreads = 1; w.X = 1; w.V = 1;
        for (int t = 0; t < NT; ++t) w.in[t] = Trow + (long long)t * TPAD + e;
        w.in_sc = (long long)NT * TPAD;
        w.out[0] = prm.out + (long long)(mu0 - prm.out_mu_base) * prm.M1 * prm.M2 + (mu_lo + e);
        w.out_smu = prm.M2; w.mu_base = 0;
        w.s_begin = pc.s_begin; w.s_end = pc.s_end; w.N = prm.N1;
        if (prm.npiece > 1) {           // the piece filter of the walk: rows of axis 1 in [w_lo, w_hi)
            w.nsplit = 2;               // (pb_walk_range reads entry `piece` of the sp_ arrays; only slot 0 is filled)
            w.sp_s_begin[0] = pc.s_begin; w.sp_s_end[0] = pc.s_end;
            w.sp_w_lo[0] = pc.w_lo; w.sp_w_hi[0] = pc.w_hi;
            w.sp_f_lo[0] = prm.first1[pc.s_begin]; w.sp_f_hi[0] = prm.N1;
        }
        w.first = prm.first1; w.V2 = prm.V1; w.ret_mu = prm.ret_mu1;
        w.f_lo = prm.first1[0]; w.f_hi = prm.N1;
        w.regular = 1;
        pb_walk_line<PB, P, Q>(w, 0, prm.V1);
        if (mirror) {
            const double* src = w.out[0];
            double* dst = prm.out + (long long)(mu0t - prm.out_mu_base) * prm.M1 * prm.M2 + prm.tr2[mu_lo + e];
            for (int mu1 = 0; mu1 < prm.M1; ++mu1) {
                const int i1 = prm.pair_i1[mu1];
                if (i1 >= pc.w_lo && i1 < pc.w_hi) dst[(long long)prm.tr1[mu1] * prm.M2] = src[(long long)mu1 * prm.M2];
            }
        }
    }
}

#if defined(__CUDACC__)
// Producer warps: NH per grid row.  Forms with several terms split a row's work between two warps
// (terms / input streams of "half" 0 and 1) so that a producer warp carries about as many FP64
// instructions per span as a consumer warp — the block synchronises once per span and the slowest
// warp sets the pace (ncu: with one warp per row the consumers waited 12 % of the time).
// Measured on B200 (3D p=3 n=128 stiffness): one warp per row with the lane's basis values in
// registers 4.9 ms; two warps per row with the values in shared memory (512 threads need <= 128
// registers) 5.6 ms — the extra shared-memory reads (32 per term and row) cost more than the better
// balance gains; with the values in registers and the register file re-partitioned by setmaxnreg
// (PB_S32_SPLIT == 2, below) 4.7 ms against 4.5 ms for one warp per row.  PB_S32_SPLIT selects the variant.
#ifndef PB_S32_SPLIT
#define PB_S32_SPLIT 0
#endif
// PB_S32_SPLIT == 2: two warps per row AND the values in registers: the block is launched with 512 threads
// (128 registers each) and re-partitions its register file with `setmaxnreg` — the two producer warpgroups
// grow to 168 registers, the two consumer warpgroups shrink to 88 (8 x 32 x (168 + 88) = 64 K registers).
template <class Form, int P, int Q> struct PbS32Split {
    static constexpr bool ALIGNED = (2 * Q) % 4 == 0 && PbS32Cfg<P>::NCW % 4 == 0;     // setmaxnreg works on warpgroups of 4 warps
    static constexpr bool REGSPLIT = (PB_S32_SPLIT == 2) && Form::NT > 1 && ALIGNED;
    static constexpr int NH = ((PB_S32_SPLIT == 1 && Form::NT > 1) || REGSPLIT) ? 2 : 1;
    static constexpr int NSTR = Form::NIN / NH;                 // input streams per half
    static constexpr int half_of_term(int t) { return (NH == 1 || t == 0) ? 0 : 1; }
};

template <class Form, int P, int Q>
struct PbS32Smem {     // dynamic shared memory layout (bytes)
    static constexpr int P1 = P + 1, NIN = Form::NIN, NT = Form::NT, NST = 3;
    static constexpr int NH = PbS32Split<Form, P, Q>::NH, NSTR = PbS32Split<Form, P, Q>::NSTR;
    static constexpr int SEG = 32 * Q, TPAD = PbS32Cfg<P>::TPAD;
    size_t v1, ret1, dlane, ring, tbuf, actv, total;
    PB_HD PbS32Smem(int G1, int N1) {
        size_t o = 0;
        v1 = o;    o += ((size_t)G1 * 2 * P1 * sizeof(double) + 127) & ~size_t(127);
        ret1 = o;  o += ((size_t)N1 * (2 * P + 1) * 2 * sizeof(int) + 127) & ~size_t(127);
        dlane = o; o += (size_t)Q * 2 * P1 * 32 * sizeof(double);
        ring = o;  o += (size_t)Q * NH * NST * NSTR * SEG * sizeof(double);
        tbuf = o;  o += (size_t)2 * Q * NT * TPAD * sizeof(double);
        actv = o;  o += (size_t)TPAD * sizeof(int) + 128;
        total = o;
    }
};

// local block of one span with the lane's basis values read from shared memory (sD[(gq*2+fl)*P1 + a][lane])
template <class Plan, int P, int Q>
PB_D void pb_span_block_sd(const double (&x)[Q][Plan::NOPS], const double* sDl, double (&L)[P + 1][P + 1]) {
    constexpr int P1 = P + 1;
    double D[Q][2][P1];
#pragma unroll
    for (int gq = 0; gq < Q; ++gq)
#pragma unroll
        for (int fl = 0; fl < 2; ++fl)
#pragma unroll
            for (int a = 0; a < P1; ++a) D[gq][fl][a] = sDl[((gq * 2 + fl) * P1 + a) * 32];
    pb_span_block<Plan, P, Q>(x, D, L);
}

template <class Form, int P, int Q>
// (two blocks per SM for single-term forms were tried: 80 registers per thread spill in the consumers' walk, 2.0 -> 2.7 ms)
__global__ void __launch_bounds__((Q * PbS32Split<Form, P, Q>::NH + PbS32Cfg<P>::NCW) * 32, 1) pb_s32_kernel(const __grid_constant__ PbS32Params prm) {
    constexpr int P1 = P + 1, NIN = Form::NIN, NT = Form::NT;
    using SP = PbS32Split<Form, P, Q>;
    constexpr int NH = SP::NH, NSTR = SP::NSTR;
    constexpr int TPAD = PbS32Cfg<P>::TPAD, NPROD = Q * NH;
    using SM = PbS32Smem<Form, P, Q>;
    constexpr int NST = SM::NST, SEG = SM::SEG;
    constexpr bool VEC = (Q % 2 == 0);
    constexpr int STAGE = NSTR * SEG;
    extern __shared__ __align__(128) unsigned char pb_s32_raw[];
    const SM lay(prm.v1_rows, prm.N1);
    double* sV1 = reinterpret_cast<double*>(pb_s32_raw + lay.v1);
    int* sOff = reinterpret_cast<int*>(pb_s32_raw + lay.ret1);      // per (function, k) of axis 1: mu1 * M2 and tr1[mu1] * M2 (or -1)
    double* sD = reinterpret_cast<double*>(pb_s32_raw + lay.dlane);
    double* sRing = reinterpret_cast<double*>(pb_s32_raw + lay.ring);
    double* sT = reinterpret_cast<double*>(pb_s32_raw + lay.tbuf);
    int* sAct = reinterpret_cast<int*>(pb_s32_raw + lay.actv);

    const int npc = prm.npiece > 1 ? prm.npiece : 1;
    // block -> (task, piece of axis 1); the first n_whole tasks are not cut
    const bool whole = (int)blockIdx.x < prm.n_whole;
    const int task = whole ? (int)blockIdx.x : prm.n_whole + ((int)blockIdx.x - prm.n_whole) / npc;
    const PbS32Piece pc = pb_s32_piece(prm, whole ? -1 : ((int)blockIdx.x - prm.n_whole) % npc);
    // task -> (kept entries, batch): one entry and 32 lanes, or up to tail_k entries with tail_w lanes each
    const int nb_full = prm.tail_k > 1 ? prm.nbatch - 1 : prm.nbatch;
    const bool packed = task >= prm.nkeep * nb_full;
    const int kidx = packed ? (task - prm.nkeep * nb_full) * prm.tail_k : task / nb_full;
    const int batch = packed ? prm.nbatch - 1 : task % nb_full;
    const int ngrp = packed ? pb_min(prm.tail_k, prm.nkeep - kidx) : 1;
    const int GL = packed ? prm.tail_w : 32;                    // lanes per entry
    const int TG = packed ? prm.tail_w * (2 * P + 1) : TPAD;    // T positions per entry
    const int mu0 = prm.keep[kidx];                             // entry of group 0
    const int mu0t = prm.symmetric ? prm.tr0[mu0] : mu0;
    // (broadcast from lane 0: tells the compiler that the role is uniform across the warp — otherwise every
    // shuffle of the producers is wrapped in a convergence barrier, ~35 % more instructions per row)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool producer = warp < NPROD;
    const int gl = lane / GL, lg = lane - gl * GL;              // entry and lane inside it (gl = 0, lg = lane unless packed)

    // ---- block setup: axis-1 tables, zeroed rings and T buffers, owned positions ---------------------
    const int v1_row0 = pc.s_begin * Q;                 // the axis-1 table of the spans this block walks
    {
        const int cnt = (pc.s_end - pc.s_begin) * Q * 2 * P1;
        const double* src = prm.V1 + (long long)v1_row0 * 2 * P1;
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) sV1[t] = src[t];
    }
    for (int t = threadIdx.x; t < prm.N1 * (2 * P + 1); t += blockDim.x) {
        int mu1 = prm.ret_mu1[t];
        if (mu1 >= 0) {         // piece filter on the row of the pair: entry k of function f is (f, f+k) or (f+k-P, f)
            const int f1 = t / (2 * P + 1), k = t % (2 * P + 1);
            const int i1 = (k <= P) ? f1 : f1 + (k - P);
            if (i1 < pc.w_lo || i1 >= pc.w_hi) mu1 = -1;
        }
        sOff[2 * t] = mu1 >= 0 ? mu1 * prm.M2 : -1;
        sOff[2 * t + 1] = (mu1 >= 0 && prm.symmetric) ? prm.tr1[mu1] * prm.M2 : -1;
    }
    for (int t = threadIdx.x; t < NPROD * NST * STAGE; t += blockDim.x) sRing[t] = 0.0;
    for (int t = threadIdx.x; t < 2 * Q * NT * TPAD; t += blockDim.x) sT[t] = 0.0;
    for (int t = threadIdx.x; t <= TPAD; t += blockDim.x) sAct[t] = 0;
    const int sb = batch * (32 - P);
    {   // basis values of axis 2 on the 32 spans of the batch, lane-contiguous
        for (int t = threadIdx.x; t < Q * 2 * P1 * 32; t += blockDim.x) {
            const int l = t & 31, c = t >> 5;                       // c = (gq*2 + fl)*P1 + a
            const int gq = c / (2 * P1), r = c % (2 * P1);
            const int sp = sb + l % GL;
            sD[t] = (sp < prm.n2 && l / GL < ngrp) ? prm.V2[(long long)(sp * Q + gq) * 2 * P1 + r] : 0.0;
        }
    }
    __syncthreads();

    const int s = sb + lg;
    const int m = prm.first2[0] + sb + lg;
    const bool writer = (batch == 0 || lg >= P) && m < prm.N2 && gl < ngrp;
    // band positions of this lane's entries and the smallest one of the batch (every warp computes it)
    int mu[2 * P + 1];
    int mu_lo = 0x7fffffff;
#pragma unroll
    for (int k = 0; k <= 2 * P; ++k) {
        mu[k] = writer ? __ldg(prm.ret_mu2 + (long long)m * (2 * P + 1) + k) : -1;
        if (mu[k] >= 0) mu_lo = pb_min(mu_lo, mu[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mu_lo = pb_min(mu_lo, __shfl_xor_sync(0xffffffffu, mu_lo, o));
    if (mu_lo == 0x7fffffff) return;                    // nothing to write in this batch (block-uniform)
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k <= 2 * P; ++k)
            if (mu[k] >= 0) sAct[gl * TG + mu[k] - mu_lo] = 1;
    }
    __syncthreads();
    (void)s;

    const long long plane = (long long)prm.G1 * prm.G2;
    if constexpr (SP::REGSPLIT) {
        if (producer) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    }
    if (producer) {
        // ========================= phase A: one grid row g1 per NH warps and span ====================
        const int prow = warp / NH, half = warp % NH;
        double* ring = sRing + (size_t)warp * NST * STAGE;
        const double* sDl = sD + lane;
#if PB_S32_SPLIT != 1
        double Dreg[Q][2][P1];          // (also the split variant with re-partitioned registers)          // basis values of this lane's span
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
#pragma unroll
            for (int fl = 0; fl < 2; ++fl)
#pragma unroll
                for (int a = 0; a < P1; ++a) Dreg[gq][fl][a] = sDl[((gq * 2 + fl) * P1 + a) * 32];
#endif
        // slot offsets of this lane's entries in a T row (-1: not a writer)
        int tpos[2 * P + 1];
#pragma unroll
        for (int k = 0; k <= 2 * P; ++k) tpos[k] = mu[k] >= 0 ? gl * TG + mu[k] - mu_lo : -1;
        const long long seg_node0 = (long long)sb * Q;
        const int seg_nodes = (pb_min(prm.n2, sb + GL) - sb) * Q;   // nodes of one entry's segment
        const int ent_nodes = GL * Q;                                  // ring doubles per entry
        constexpr int NPIECE = VEC ? (SEG / 2 + 31) / 32 : (SEG + 31) / 32;
        // ring element -> (entry, node of its segment); entries other than the first read another plane of X1:
        // dpl[h][0] for the streams read at mu0, dpl[h][1] for those read at the transposed entry
        int goff[NPIECE], soff[NPIECE];
        long long dpl[NPIECE][2];
#pragma unroll
        for (int h = 0; h < NPIECE; ++h) {
            const int c = lane + 32 * h;
            const int r = VEC ? 2 * c : c;                              // first ring double of this copy
            const int ge = r / ent_nodes, within = r - ge * ent_nodes;
            goff[h] = (r < SEG && ge < ngrp && within < seg_nodes) ? within : -1;
            if constexpr (VEC) soff[h] = 2 * ((Q == 4) ? (c ^ ((c >> 3) & 1)) : c);
            else soff[h] = c;
            dpl[h][0] = dpl[h][1] = 0;
            if (goff[h] >= 0 && ge > 0) {
                const int mg = prm.keep[kidx + ge];
                dpl[h][0] = (long long)(mg - mu0) * ((long long)prm.G1 * prm.G2);
                dpl[h][1] = (long long)((prm.symmetric ? prm.tr0[mg] : mg) - mu0t) * ((long long)prm.G1 * prm.G2);
            }
        }
        // Source addresses.  Per copy piece and per kind of stream (read at mu0 / at the transposed entry) the lane
        // keeps ONE running byte pointer for the next span to request; a stream adds the (warp-uniform) offset of
        // its X1 term.  (Composing every address from scratch cost six integer instructions per copy, a tenth of
        // the producers' instruction stream.)
        const long long span_bytes = (long long)Q * prm.G2 * (long long)sizeof(double);     // consecutive spans of this warp's row
        const char* cpn[NPIECE][2];
#pragma unroll
        for (int h = 0; h < NPIECE; ++h)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                cpn[h][k] = reinterpret_cast<const char*>(prm.X1 + ((long long)((k ? mu0t : mu0) - prm.x1_mu_base) * plane + seg_node0
                                                                    + (long long)prow * prm.G2 + (goff[h] >= 0 ? goff[h] : 0) + dpl[h][k]
                                                                    + (long long)pc.s_begin * Q * prm.G2));
        // requests must come in span order (each call is for the span after the previous one)
        auto issue = [&](int st) {
            double* dst = ring + (size_t)st * STAGE;
            pb_static_for<0, NSTR>([&](auto J) {
                constexpr int j = decltype(J)::value;
                // the input stream of this half: global stream index = half * NSTR + j (both halves are instantiated)
                const int t0 = pb_s32_slot<Form>(prm, j), t1 = pb_s32_slot<Form>(prm, (NH - 1) * NSTR + j);
                const bool r0 = Form::in_tr(j), r1 = Form::in_tr((NH - 1) * NSTR + j);
                const int term = half == 0 ? t0 : t1;
                const bool trn = half == 0 ? r0 : r1;
                if (Form::RUNTIME && term < 0) return;      // the form has no such term: the (zero-filled) ring slot is never written
                const long long toff = (long long)term * prm.x1_stride * (long long)sizeof(double);
#pragma unroll
                for (int h = 0; h < NPIECE; ++h) {
                    if (goff[h] >= 0) {
                        const char* sp = (trn ? cpn[h][1] : cpn[h][0]) + toff;
                        if constexpr (VEC) pb_cp_async16(dst + j * SEG + soff[h], sp);
                        else pb_cp_async8(dst + j * SEG + soff[h], sp);
                    }
                }
            });
#pragma unroll
            for (int h = 0; h < NPIECE; ++h) { cpn[h][0] += span_bytes; cpn[h][1] += span_bytes; }
        };
#pragma unroll
        const int nsp = pc.s_end - pc.s_begin;
        for (int j = 0; j < NST - 1; ++j) {
            if (j < nsp) issue(j);
            pb_cp_async_commit();
        }
        double nmask[P > 0 ? P : 1];             // 1.0 where the lane has a neighbour q spans below inside its entry
#pragma unroll
        for (int q = 1; q <= P; ++q) nmask[q - 1] = lg >= q ? 1.0 : 0.0;
        // one term: local block, neighbour sums, slots of the T row
        auto do_term = [&](auto TT, const double* src, double* Tw) {
            constexpr int t = decltype(TT)::value;
            using TP = typename Form::template Term<t>::Plan;
            double xt[Q][TP::NOPS];
            pb_static_for<0, TP::NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                constexpr int sidx = Form::template Term<t>::stream(i) % NSTR;     // slot of the stream inside its half
                if constexpr (VEC) {
#pragma unroll
                    for (int h = 0; h < Q / 2; ++h) {
                        const int c = lane * (Q / 2) + h;
                        const int pc = (Q == 4) ? (c ^ ((c >> 3) & 1)) : c;
                        const double2 v = *reinterpret_cast<const double2*>(src + sidx * SEG + 2 * pc);
                        xt[2 * h][i] = v.x;
                        xt[2 * h + 1][i] = v.y;
                    }
                } else {
#pragma unroll
                    for (int gq = 0; gq < Q; ++gq) xt[gq][i] = src[sidx * SEG + lane * Q + gq];
                }
            });
            double L[P1][P1];
#if PB_S32_SPLIT == 1
            pb_span_block_sd<TP, P, Q>(xt, sDl, L);
#else
            pb_span_block<TP, P, Q>(xt, Dreg, L);
#endif
            // a single input with the same derivative flag on both sides gives a symmetric block: the entry (f+d, f)
            // equals (f, f+d) of the same lane — no second neighbour sum, and the lower half of L is never formed
            constexpr bool tsym = TP::NOPS == 1 && TP::op(0).ft == TP::op(0).fu;
            double upper[P1];
#pragma unroll
            for (int kk = 0; kk <= 2 * P; ++kk) {
                const int d = (kk <= P) ? kk : kk - P;
                double sum;
                if (tsym && kk > P) {
                    sum = upper[d];
                } else {
                    sum = (kk <= P) ? L[0][d] : L[d][0];
#pragma unroll
                    for (int q = 1; q <= P; ++q) {
                        if (q <= P - d) {
                            const double vsh = __shfl_up_sync(0xffffffffu, (kk <= P) ? L[q][q + d] : L[q + d][q], q);
                            sum = fma(nmask[q - 1], vsh, sum);      // lanes below q take nothing (0/1 mask: one FMA, no select)
                        }
                    }
                    if (kk <= P) upper[d] = sum;
                }
                if (tpos[kk] >= 0) Tw[t * TPAD + tpos[kk]] = sum;
            }
        };
        int st = 0;
        for (int s1 = 0; s1 <= nsp; ++s1) {
            if (s1 < nsp) {
                __syncwarp();
                if (s1 + NST - 1 < nsp) issue((st + NST - 1) % NST);
                pb_cp_async_commit();
                pb_cp_async_wait<NST - 1>();
                __syncwarp();
                const double* src = ring + (size_t)st * STAGE;
                double* Tw = sT + ((size_t)(s1 & 1) * Q + prow) * NT * TPAD;
                pb_static_for<0, NT>([&](auto TT) {
                    constexpr int t = decltype(TT)::value;
                    if (SP::half_of_term(t) == half) do_term(TT, src, Tw);      // warp-uniform
                });
                st = (st + 1) % NST;
            }
            __syncthreads();
        }
        return;
    }

    // ==================================== phase B: walk along axis 1 ==================================
    using PB = typename Form::PlanB;
    const int e = threadIdx.x - NPROD * 32;             // T position owned by this thread: entry ge, band position mu_lo + el
    const bool act = sAct[e] != 0;
    const int ge = e / TG, el = e - ge * TG;
    const int mu0e = act ? prm.keep[kidx + ge] : mu0;
    bool mirror;
    pb_s32_keep(prm, mu0e, mirror);
    double* outp = prm.out + (long long)(mu0e - prm.out_mu_base) * prm.M1 * prm.M2 + (mu_lo + el);
    double* outp_t = outp;
    if (mirror && act)
        outp_t = prm.out + (long long)(prm.tr0[mu0e] - prm.out_mu_base) * prm.M1 * prm.M2 + __ldg(prm.tr2 + mu_lo + el);
    double acc[P1][P1];
#pragma unroll
    for (int a = 0; a < P1; ++a)
#pragma unroll
        for (int b = 0; b < P1; ++b) acc[a][b] = 0.0;
    int f = prm.first1[pc.s_begin];

    auto retire = [&](auto RC, int fr) {                // RC: phase in which function fr is row / column 0
        constexpr int R = decltype(RC)::value;
        const int2* ro = reinterpret_cast<const int2*>(sOff) + fr * (2 * P + 1);
#pragma unroll
        for (int k = 0; k <= 2 * P; ++k) {
            const int2 o = ro[k];
            const int a = (k <= P) ? 0 : (k - P);
            const int b = (k <= P) ? k : 0;
            if (o.x >= 0 && act) {
                const double v = acc[(a + R) % P1][(b + R) % P1];
                outp[o.x] = v;
                if (mirror) outp_t[o.y] = v;
            }
        }
    };
    auto node = [&](auto RC, int row, const double* Tg) {
        constexpr int R = decltype(RC)::value;
        const double* Vn = sV1 + (long long)(row - v1_row0) * (2 * P1);
        double D[2][P1];
#pragma unroll
        for (int a = 0; a < P1; ++a) { D[0][a] = Vn[a]; D[1][a] = Vn[P1 + a]; }
        constexpr bool by_fu = pb_group_by_fu<PB>(0);
        pb_static_for<0, 2>([&](auto FL) {
            constexpr int fl = decltype(FL)::value;
            constexpr int cnt = by_fu ? pb_count_fu<PB>(0, fl) : pb_count_ft<PB>(0, fl);
            if constexpr (cnt > 0) {
                double y[P1];
                constexpr int lead = pb_first_in_group<PB>(0, fl, by_fu);
                pb_static_for<0, PB::NOPS>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr PbOp op = PB::op(i);
                    if constexpr ((by_fu ? op.fu : op.ft) == fl) {
                        constexpr int other = by_fu ? op.ft : op.fu;
                        const double xv = Tg[i * TPAD];
                        if constexpr (i == lead) {
#pragma unroll
                            for (int c = 0; c < P1; ++c) y[c] = D[other][c] * xv;
                        } else {
#pragma unroll
                            for (int c = 0; c < P1; ++c) y[c] = fma(D[other][c], xv, y[c]);
                        }
                    }
                });
#pragma unroll
                for (int a = 0; a < P1; ++a)
#pragma unroll
                    for (int b = 0; b < P1; ++b) {
                        const double l = by_fu ? y[a] : D[fl][a], r = by_fu ? D[fl][b] : y[b];
                        acc[(a + R) % P1][(b + R) % P1] = fma(l, r, acc[(a + R) % P1][(b + R) % P1]);
                    }
            }
        });
    };
    int phase = 0;
    const int nsp = pc.s_end - pc.s_begin;
    for (int s1 = 0; s1 <= nsp; ++s1) {
        if (s1 > 0) {
            const int sp = s1 - 1;                      // the span the producers finished before the last barrier
            const double* Tg = sT + (size_t)(sp & 1) * Q * NT * TPAD + e;
            pb_static_for<0, P1>([&](auto RC) {
                constexpr int R = decltype(RC)::value;
                if (phase == R) {
                    if (sp > 0) {
                        retire(PbIC<(R + P) % P1>{}, f);
                        ++f;
#pragma unroll
                        for (int a = 0; a < P1; ++a) {  // row P and column P enter the window empty
                            acc[(a + R) % P1][(P + R) % P1] = 0.0;
                            acc[(P + R) % P1][(a + R) % P1] = 0.0;
                        }
                    }
#pragma unroll
                    for (int gq = 0; gq < Q; ++gq) node(RC, (pc.s_begin + sp) * Q + gq, Tg + (size_t)gq * NT * TPAD);
                }
            });
            phase = (phase + 1 == P1) ? 0 : phase + 1;
        }
        __syncthreads();
    }
    // flush: the functions still in the window, starting in the phase of the last span
    const int last_ph = (nsp - 1) % P1;
    for (int t = 0; t < P1; ++t) {
        if (f < prm.N1) {
            pb_static_for<0, P1>([&](auto PH) {
                constexpr int ph = decltype(PH)::value;
                if ((last_ph + t) % P1 == ph) retire(PH, f);
            });
        }
        ++f;
    }
}
#endif

typedef int (*PbS32Launch)(const PbS32Params* prm, void* stream);
extern "C" __attribute__((visibility("default"))) void pb200_register_s32(int form, int P, int Q, PbS32Launch fn);
PbS32Launch pb_find_s32(int form, int P, int Q);
