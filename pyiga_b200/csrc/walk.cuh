// Sum-factorised banded contraction along one tensor axis ("walk" stage).
//
// This is the core of the B200 assembly path.  The reference computes every matrix entry by a
// full d-dimensional quadrature loop (pyiga/assemblers.pyx:1455-1540 `combine` + `entry_impl`,
// driven by pyiga/genericasm.pxi:691-758).  Here the same sums are evaluated by *pairwise sum
// factorisation*: the d-dimensional sum is split into d one-dimensional contractions, and each
// contraction turns a Gauss-node axis (length n*q) into a band axis (the list of function pairs
// (i,j) with joint support on that axis, in the order of `compute_sparsity_ij`,
// pyiga/mlmatrix.py:420-440).
//
// One thread owns one "line" (all other indices fixed) and walks along the node axis span by span.
// On span s exactly the p+1 consecutive functions first[s] .. first[s]+p are active, so the thread
// keeps a (p+1)x(p+1) register window acc[a][b] of partial sums for the pairs
// (f+a, f+b), f = window base.  Per node it performs a rank-1 style update
//     acc[a][b] += D_ft[a] * D_fu[b] * x          (D_0 = values, D_1 = first derivatives)
// factored as y[b] = D_fu[b]*x ; acc[a][b] += D_ft[a]*y[b]  (or the mirrored grouping), which needs
// only 2(p+1) table values per node; those are warp-uniform and come from shared memory (staged by
// a TMA bulk copy).  When the window base moves on, the pairs that lose their last span are
// complete and are written out ("retired").
//
// Lanes of a warp own neighbouring lines, so the loads of x and the stores of retired entries are
// coalesced whenever the walk axis is not the fastest one in memory.
#pragma once
#include "common.cuh"

struct PbOp {       // acc[out] += D_ft (x) D_fu * in
    int in;         // input term
    int tr;         // read the input at the transposed (u, v) line
    int ft;         // derivative order of the test function on this axis (0/1)
    int fu;         // derivative order of the trial function on this axis (0/1)
    int out;        // output term
};

#define PB_WALK_MAXOPS 8
#define PB_WALK_MAXSPLIT 4
#define PB_WALK_MAXOUT 6

struct PbWalkParams {
    // ---- thread grid: tid -> (u, v, x),  x fastest -------------------------------------------
    long long nthreads;
    int X, V;
    int u_begin;                    // absolute index of the first u line
    int u_base_in, u_base_out;      // absolute u that sits in slot 0 of the in / out buffers
    const int* tr_u;                // transposed-pair tables for the u / v coordinates (or null)
    const int* tr_v;
    const int* u_pair_i;            // (i,j) of band entry u, for the slab filter (or null)
    const int* u_pair_j;
    // line filter on the band entry u = (i,j) of axis 0, per output term:
    //   0: all u;  1: i in [lo,hi);  2: i or j in [lo,hi);
    //   3: i in [lo,hi) and (i <= j or j outside [lo,hi))   ("upper half": the rest is mirrored)
    // a line runs if any output wants it; inputs that only feed unwanted outputs are not loaded
    int u_mode[PB_WALK_MAXOUT], u_lo, u_hi;
    const int* v_pair_i;            // (i,j) of band entry v (final stage with mirroring)
    const int* v_pair_j;
    int mirror;                     // final stage: also write the transposed line (symmetric forms)
    // ---- input / output terms ---------------------------------------------------------------
    const double* in[PB_WALK_MAXOPS];   // per op: base pointer of its input term
    long long in_su, in_sv, in_sx, in_sc;   // strides (doubles) of u, v, x and of the node index
    double* out[PB_WALK_MAXOUT];
    long long out_su, out_sv, out_sx, out_smu;
    int mu_base;                    // band index stored in slot 0 of the output band axis
    // ---- walk axis ---------------------------------------------------------------------------
    int s_begin, s_end;             // spans walked
    int N;                          // number of functions on the axis
    const int* first;               // [n]
    const double* V2;               // [G][2][P+1]
    const int* ret_mu;              // [N][2P+1]
    int w_mode[PB_WALK_MAXOUT], w_lo, w_hi;     // retire filter on the walk-axis pair (i,j), as u_mode
    int f_lo, f_hi;                 // functions retired by this walk: [first[s_begin], min(N, first[s_end-1]+P+1))
    // ---- optional split of the walk axis (grid.y pieces) --------------------------------------
    // A stage whose lines fill the GPU only 1.x times is cut into `nsplit` pieces along the walk
    // axis: piece y retires the pairs whose row lies in [sp_w_lo[y], sp_w_hi[y]) and walks only the
    // spans those rows see, so that the tail wave is shorter (the pieces overlap by p spans).
    // The piece filter is applied on top of the stage's own walk-axis filter; `w_ext_lo/hi` is the
    // row range the pieces partition (all rows for unfiltered stages).
    int nsplit, w_ext_lo, w_ext_hi;
    int regular;                    // every span advances `first` by one: the rotating-window walk may be used
    int sp_s_begin[PB_WALK_MAXSPLIT], sp_s_end[PB_WALK_MAXSPLIT];
    int sp_w_lo[PB_WALK_MAXSPLIT], sp_w_hi[PB_WALK_MAXSPLIT];
    int sp_f_lo[PB_WALK_MAXSPLIT], sp_f_hi[PB_WALK_MAXSPLIT];
    // ---- fused stage 1 (walk_geo.cuh): host pointer to the PbGeoLineParams of the launch ---------
    const void* geo_line;
};

PB_HD bool pb_keep(int mode, int i, int j, int lo, int hi) {
    const bool ki = (i >= lo && i < hi), kj = (j >= lo && j < hi);
    return mode == 0 || (mode == 1 && ki) || (mode == 2 && (ki || kj)) || (mode == 3 && ki && (i <= j || !kj));
}

// the part of the walk axis one thread block (or emulated piece) works on
struct PbWalkRange { int s_begin, s_end, p_lo, p_hi, f_lo, f_hi; bool split; };
PB_HD PbWalkRange pb_walk_range(const PbWalkParams& prm, int y) {
    PbWalkRange r;
    r.split = prm.nsplit > 1;
    if (r.split) {
        r.s_begin = prm.sp_s_begin[y]; r.s_end = prm.sp_s_end[y];
        r.p_lo = prm.sp_w_lo[y]; r.p_hi = prm.sp_w_hi[y];
        r.f_lo = prm.sp_f_lo[y]; r.f_hi = prm.sp_f_hi[y];
    } else {
        r.s_begin = prm.s_begin; r.s_end = prm.s_end;
        r.p_lo = 0; r.p_hi = 0x7fffffff;
        r.f_lo = prm.f_lo; r.f_hi = prm.f_hi;
    }
    return r;
}
// retire filter of output o for the pair (i, j) of the walk axis: the stage's own filter and the piece
PB_HD bool pb_walk_keep(const PbWalkParams& prm, const PbWalkRange& rg, int o, int i, int j) {
    return pb_keep(prm.w_mode[o], i, j, prm.w_lo, prm.w_hi) && i >= rg.p_lo && i < rg.p_hi;
}

template <int I> struct PbIC { static constexpr int value = I; };

template <int B, int E, class F>
PB_HD void pb_static_for(F&& f) {
    if constexpr (B < E) {
        f(PbIC<B>{});
        pb_static_for<B + 1, E>(f);
    }
}

// compile-time queries on a plan
template <class Plan> constexpr int pb_count_ft(int out, int ft) {
    int c = 0;
    for (int i = 0; i < Plan::NOPS; ++i) c += (Plan::op(i).out == out && Plan::op(i).ft == ft) ? 1 : 0;
    return c;
}
template <class Plan> constexpr int pb_count_fu(int out, int fu) {
    int c = 0;
    for (int i = 0; i < Plan::NOPS; ++i) c += (Plan::op(i).out == out && Plan::op(i).fu == fu) ? 1 : 0;
    return c;
}
// first op of the group (out, flag) under the chosen grouping
template <class Plan> constexpr int pb_first_in_group(int out, int fl, bool by_fu) {
    for (int i = 0; i < Plan::NOPS; ++i)
        if (Plan::op(i).out == out && (by_fu ? Plan::op(i).fu : Plan::op(i).ft) == fl) return i;
    return -1;
}
// group the ops of an output by the trial flag when that gives fewer rank-1 updates
template <class Plan> constexpr bool pb_group_by_fu(int out) {
    int nft = (pb_count_ft<Plan>(out, 0) > 0) + (pb_count_ft<Plan>(out, 1) > 0);
    int nfu = (pb_count_fu<Plan>(out, 0) > 0) + (pb_count_fu<Plan>(out, 1) > 0);
    return nfu < nft;
}

// The body executed by one thread.  `Vt` points at the [G][2][P+1] table (shared or global).
// Input staging policies of the walk: how the Q*NOPS inputs of a span reach the registers.
//  * PbRegLoader: plain loads one span ahead into registers (host emulation, fall-back).
//  * PbAsyncLoader (device): cp.async into a per-thread ring in shared memory, NST-1 spans ahead;
//    data in flight holds no registers and needs no register moves, so the lead is really NST-1.
template <class Plan, int Q>
struct PbRegLoader {
    static constexpr int NOPS = Plan::NOPS;
    static constexpr bool ROLLED = false;
    const double* src[NOPS];
    bool has[NOPS];
    long long sc;
    int s_end;
    double xq[Q][NOPS];
    PB_HD void load(int s) {
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
            pb_static_for<0, NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                xq[gq][i] = has[i] ? src[i][(long long)(s * Q + gq) * sc] : 0.0;
            });
    }
    PB_HD void prime(int s_begin) { load(s_begin); }
    // hand out the inputs of span s and start fetching a later span
    PB_HD void next(int s, double (&xc)[Q][NOPS]) {
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
            pb_static_for<0, NOPS>([&](auto I) { constexpr int i = decltype(I)::value; xc[gq][i] = xq[gq][i]; });
        if (s + 1 < s_end) load(s + 1);
    }
};

struct PbWalkTables {       // where the thread finds the walk-axis tables (shared or global memory)
    const double* V;        // indexed by absolute node
    const int* first;       // indexed by absolute span
    const int* ret_mu;      // indexed by absolute function * (2P+1); the staged copy (template flag ENC
                            // of the walk) carries the per-output slab filter in bits 24..27
};
template <class Plan, int P, int Q, bool ENC, bool ROT, class Loader>
PB_HD void pb_walk_line_impl(const PbWalkParams& prm, const PbWalkRange& rg, long long tid, const PbWalkTables& tb, Loader& ld);

template <class Plan, int P, int Q, int NPF = 1>
PB_HD void pb_walk_line(const PbWalkParams& prm, long long tid, const double* __restrict__ Vt, int piece = 0) {
    PbRegLoader<Plan, Q> ld;
    PbWalkTables tb;
    tb.V = Vt; tb.first = prm.first; tb.ret_mu = prm.ret_mu;
    const PbWalkRange rg = pb_walk_range(prm, piece);
    if (prm.regular) pb_walk_line_impl<Plan, P, Q, false, true>(prm, rg, tid, tb, ld);
    else pb_walk_line_impl<Plan, P, Q, false, false>(prm, rg, tid, tb, ld);
}

template <class Plan, int P, int Q, bool ENC, bool ROT, class Loader>
PB_HD void pb_walk_line_impl(const PbWalkParams& prm, const PbWalkRange& rg, long long tid, const PbWalkTables& tb, Loader& ld) {
    const double* __restrict__ Vt = tb.V;
    constexpr int P1 = P + 1;
    constexpr int NOPS = Plan::NOPS, NOUT = Plan::NOUT;

    // ---- decode the line ------------------------------------------------------------------
    const int x = (int)(tid % prm.X);
    const long long t2 = tid / prm.X;
    const int v = (int)(t2 % prm.V);
    const int u = (int)(t2 / prm.V) + prm.u_begin;
    bool want[NOUT];
    {
        bool any = false;
        const bool filt = prm.u_pair_i != nullptr;
        const int ui = filt ? prm.u_pair_i[u] : 0, uj = filt ? prm.u_pair_j[u] : 0;
        pb_static_for<0, NOUT>([&](auto O) {
            constexpr int o = decltype(O)::value;
            want[o] = !filt || pb_keep(prm.u_mode[o], ui, uj, prm.u_lo, prm.u_hi);
            any = any || want[o];
        });
        if (!any) return;
    }
    const long long off_in = (long long)(u - prm.u_base_in) * prm.in_su + (long long)v * prm.in_sv
                             + (long long)x * prm.in_sx;
    long long off_in_tr = off_in;
    if (Plan::HAS_TR) {
        const int ut = prm.tr_u ? prm.tr_u[u] : u;
        const int vt = prm.tr_v ? prm.tr_v[v] : v;
        off_in_tr = (long long)(ut - prm.u_base_in) * prm.in_su + (long long)vt * prm.in_sv
                    + (long long)x * prm.in_sx;
    }
    const long long off_out = (long long)(u - prm.u_base_out) * prm.out_su + (long long)v * prm.out_sv
                              + (long long)x * prm.out_sx;

    // a null input pointer means "term absent" (generic forms): it reads as zero
    pb_static_for<0, NOPS>([&](auto I) {
        constexpr int i = decltype(I)::value;
        ld.has[i] = prm.in[i] != nullptr && want[Plan::op(i).out];
        ld.src[i] = prm.in[i] + (Plan::op(i).tr ? off_in_tr : off_in);
    });
    ld.sc = prm.in_sc;
    ld.s_end = rg.s_end;

    double acc[NOUT][P1][P1];
    pb_static_for<0, NOUT>([&](auto O) {
        constexpr int o = decltype(O)::value;
#pragma unroll
        for (int a = 0; a < P1; ++a)
#pragma unroll
            for (int b = 0; b < P1; ++b) acc[o][a][b] = 0.0;
    });

    // per-thread output bases and the set of outputs this line feeds; the band stride fits 32 bits
    // (checked on the host), so the offset of a band entry is one widening multiply-add
    double* outp[NOUT];
    int wantbits = 0;
    pb_static_for<0, NOUT>([&](auto O) {
        constexpr int o = decltype(O)::value;
        outp[o] = prm.out[o] + off_out;
        wantbits |= want[o] ? (1 << o) : 0;
    });
    const int smu = (int)prm.out_smu, mu_base = prm.mu_base;

    // retire the pairs that involve function f (first row / first column of the window), then
    // slide the window down by one function
    auto retire_shift = [&](int f) {
        const int* rm = tb.ret_mu + (long long)f * (2 * P + 1);
#pragma unroll
        for (int k = 0; k <= 2 * P; ++k) {
            const int mu = rm[k];
            const int a = (k <= P) ? 0 : (k - P);
            const int b = (k <= P) ? k : 0;
            if (mu >= 0) {
                // the slab filter of the walk-axis pair is the same for every line: the staged table
                // carries it as a bit mask; the plain table needs the test here
                int band = mu, mask = 0;
                if constexpr (ENC) {
                    band = mu & 0xFFFFFF;
                    mask = mu >> 24;
                } else {
                    pb_static_for<0, NOUT>([&](auto O) {
                        constexpr int o = decltype(O)::value;
                        mask |= pb_walk_keep(prm, rg, o, f + a, f + b) ? (1 << o) : 0;
                    });
                }
                mask &= wantbits;
                const long long boff = (long long)(band - mu_base) * (long long)smu;
                pb_static_for<0, NOUT>([&](auto O) {
                    constexpr int o = decltype(O)::value;
                    if (mask & (1 << o)) outp[o][boff] = acc[o][a][b];
                });
            }
        }
        pb_static_for<0, NOUT>([&](auto O) {
            constexpr int o = decltype(O)::value;
#pragma unroll
            for (int a = 0; a < P; ++a)
#pragma unroll
                for (int b = 0; b < P; ++b) acc[o][a][b] = acc[o][a + 1][b + 1];
#pragma unroll
            for (int a = 0; a < P1; ++a) {
                acc[o][a][P] = 0.0;
                acc[o][P][a] = 0.0;
            }
        });
    };

    int f = tb.first[rg.s_begin];
    ld.prime(rg.s_begin);

    if constexpr (ROT) {
        // Regular axis (every span advances the first active function by one): the window is not
        // shifted but ROTATED.  On the k-th span of the walk the pair (f+a, f+b) lives in
        // acc[(a+k) % (P+1)][(b+k) % (P+1)]; the span loop is unrolled P+1 times so that all indices
        // are compile-time constants and no register moves are needed.  The row / column that
        // enters the window is not cleared either: its first contribution of the span assigns.
        auto retire_rot = [&](auto RC, int fr) {            // RC: phase in which function fr is row/column 0
            constexpr int R = decltype(RC)::value;
            const int* rm = tb.ret_mu + (long long)fr * (2 * P + 1);
#pragma unroll
            for (int k = 0; k <= 2 * P; ++k) {
                const int mu = rm[k];
                const int a = (k <= P) ? 0 : (k - P);
                const int b = (k <= P) ? k : 0;
                if (mu >= 0) {
                    int band = mu, mask = 0;
                    if constexpr (ENC) {
                        band = mu & 0xFFFFFF;
                        mask = mu >> 24;
                    } else {
                        pb_static_for<0, NOUT>([&](auto O) {
                            constexpr int o = decltype(O)::value;
                            mask |= pb_walk_keep(prm, rg, o, fr + a, fr + b) ? (1 << o) : 0;
                        });
                    }
                    mask &= wantbits;
                    const long long boff = (long long)(band - mu_base) * (long long)smu;
                    pb_static_for<0, NOUT>([&](auto O) {
                        constexpr int o = decltype(O)::value;
                        // outputs with a symmetric window keep only the labels lo <= hi (see node())
                        const int la = (a + R) % P1, lb = (b + R) % P1;
                        if (mask & (1 << o)) outp[o][boff] = Plan::sym(o) ? acc[o][la < lb ? la : lb][la < lb ? lb : la] : acc[o][la][lb];
                    });
                }
            }
        };
        auto node = [&](auto RC, auto GQ, const double (&xc)[Q][NOPS], int s) {
            constexpr int R = decltype(RC)::value, gq = decltype(GQ)::value;
            const double* Vn = Vt + (long long)(s * Q + gq) * (2 * P1);
            double D[2][P1];
#pragma unroll
            for (int a = 0; a < P1; ++a) { D[0][a] = Vn[a]; D[1][a] = Vn[P1 + a]; }
            pb_static_for<0, NOUT>([&](auto O) {
                constexpr int o = decltype(O)::value;
                constexpr bool by_fu = pb_group_by_fu<Plan>(o);
                constexpr int fl_first = ((by_fu ? pb_count_fu<Plan>(o, 0) : pb_count_ft<Plan>(o, 0)) > 0) ? 0 : 1;
                pb_static_for<0, 2>([&](auto FL) {
                    constexpr int fl = decltype(FL)::value;
                    constexpr int cnt = by_fu ? pb_count_fu<Plan>(o, fl) : pb_count_ft<Plan>(o, fl);
                    if constexpr (cnt > 0) {
                        double y[P1];
                        constexpr int lead = pb_first_in_group<Plan>(o, fl, by_fu);
                        pb_static_for<0, NOPS>([&](auto I) {
                            constexpr int i = decltype(I)::value;
                            constexpr PbOp op = Plan::op(i);
                            if constexpr (op.out == o && (by_fu ? op.fu : op.ft) == fl) {
                                constexpr int other = by_fu ? op.ft : op.fu;
                                const double xv = xc[gq][i];
                                if constexpr (i == lead) {
#pragma unroll
                                    for (int c = 0; c < P1; ++c) y[c] = D[other][c] * xv;
                                } else {
#pragma unroll
                                    for (int c = 0; c < P1; ++c) y[c] = fma(D[other][c], xv, y[c]);
                                }
                            }
                        });
                        constexpr bool assign_new = (gq == 0 && fl == fl_first);    // first contribution of the span
#pragma unroll
                        for (int a = 0; a < P1; ++a)
#pragma unroll
                            for (int b = 0; b < P1; ++b) {
                                // Outputs whose ops all have ft == fu have a symmetric window
                                // (acc[a][b] == acc[b][a]): only a <= b is accumulated, under the
                                // ordered pair of rotated labels, and the retire reads it from there.
                                if (Plan::sym(o) && a > b) continue;
                                const double l = by_fu ? y[a] : D[fl][a], r = by_fu ? D[fl][b] : y[b];
                                const int la = (a + R) % P1, lb = (b + R) % P1;
                                double& dst = Plan::sym(o) ? acc[o][la < lb ? la : lb][la < lb ? lb : la] : acc[o][la][lb];
                                if (assign_new && (a == P || b == P)) dst = l * r;
                                else dst = fma(l, r, dst);
                            }
                    }
                });
            });
        };
        int s = rg.s_begin;
        if constexpr (Loader::ROLLED) {
            // Compact variant for loaders that stage the inputs of a span themselves (fused stage 1,
            // walk_geo.cuh): the node loop is NOT unrolled — the loader hands out the input of
            // (node, op) on request — and the row / column entering the window is cleared instead of
            // being assigned by its first contribution, so that all nodes run the same code.  The
            // unrolled body exceeds the instruction caches there (ncu: 21 % no_instructions stalls).
            auto node_rolled = [&](auto RC, int gq, int sp) {
                constexpr int R = decltype(RC)::value;
                const double* Vn = Vt + (long long)(sp * Q + gq) * (2 * P1);
                double D[2][P1];
#pragma unroll
                for (int a = 0; a < P1; ++a) { D[0][a] = Vn[a]; D[1][a] = Vn[P1 + a]; }
                double xn[NOPS];
                ld.get_node(gq, xn);
                pb_static_for<0, NOUT>([&](auto O) {
                    constexpr int o = decltype(O)::value;
                    constexpr bool by_fu = pb_group_by_fu<Plan>(o);
                    pb_static_for<0, 2>([&](auto FL) {
                        constexpr int fl = decltype(FL)::value;
                        constexpr int cnt = by_fu ? pb_count_fu<Plan>(o, fl) : pb_count_ft<Plan>(o, fl);
                        if constexpr (cnt > 0) {
                            double y[P1];
                            constexpr int lead = pb_first_in_group<Plan>(o, fl, by_fu);
                            pb_static_for<0, NOPS>([&](auto I) {
                                constexpr int i = decltype(I)::value;
                                constexpr PbOp op = Plan::op(i);
                                if constexpr (op.out == o && (by_fu ? op.fu : op.ft) == fl) {
                                    constexpr int other = by_fu ? op.ft : op.fu;
                                    const double xv = xn[i];
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
                                    if (Plan::sym(o) && a > b) continue;
                                    const double l = by_fu ? y[a] : D[fl][a], r = by_fu ? D[fl][b] : y[b];
                                    const int la = (a + R) % P1, lb = (b + R) % P1;
                                    double& dst = Plan::sym(o) ? acc[o][la < lb ? la : lb][la < lb ? lb : la] : acc[o][la][lb];
                                    dst = fma(l, r, dst);
                                }
                        }
                    });
                });
            };
            auto clear_entering = [&](auto RC) {        // row P and column P of the window in phase R
                constexpr int R = decltype(RC)::value;
                pb_static_for<0, NOUT>([&](auto O) {
                    constexpr int o = decltype(O)::value;
#pragma unroll
                    for (int a = 0; a < P1; ++a) {
                        const int la = (a + R) % P1, lp = (P + R) % P1;
                        acc[o][la < lp ? la : lp][la < lp ? lp : la] = 0.0;
                        if (!Plan::sym(o)) acc[o][la < lp ? lp : la][la < lp ? la : lp] = 0.0;
                    }
                });
            };
            // finished pairs go through the loader's staging column (thread-private shared memory):
            // the phase-specific code only copies registers there, one rolled loop stores them
            // (in two halves — row 0, then column 0 — to keep the staging column short)
            auto stage_rot = [&](auto RC, auto HALF) {  // RC: phase in which the leaving function is row/column 0
                constexpr int R = decltype(RC)::value;
                constexpr int k0 = decltype(HALF)::value == 0 ? 0 : P + 1, k1 = decltype(HALF)::value == 0 ? P + 1 : 2 * P + 1;
#pragma unroll
                for (int k = k0; k < k1; ++k) {
                    const int a = (k <= P) ? 0 : (k - P);
                    const int b = (k <= P) ? k : 0;
                    const int la = (a + R) % P1, lb = (b + R) % P1;
                    pb_static_for<0, NOUT>([&](auto O) {
                        constexpr int o = decltype(O)::value;
                        ld.stage_put((k - k0) * NOUT + o, Plan::sym(o) ? acc[o][la < lb ? la : lb][la < lb ? lb : la] : acc[o][la][lb]);
                    });
                }
            };
            const long long out_stride = NOUT > 1 ? (long long)(prm.out[NOUT > 1 ? 1 : 0] - prm.out[0]) : 0;
            auto store_staged = [&](int fr, int k0, int k1) {
                const int* rm = tb.ret_mu + (long long)fr * (2 * P + 1);
                // unrolled (P+1 entries at most): the table look-ups and staged values of all entries
                // are in flight together instead of one dependent chain per entry
#pragma unroll
                for (int k = k0; k < k0 + P + 1; ++k) {
                    if (k >= k1) break;
                    const int mu = rm[k];
                    if (mu < 0) continue;
                    int band = mu, mask = 0;
                    if constexpr (ENC) {
                        band = mu & 0xFFFFFF;
                        mask = mu >> 24;
                    } else {
                        const int a = (k <= P) ? 0 : (k - P), b = (k <= P) ? k : 0;
                        pb_static_for<0, NOUT>([&](auto O) {
                            constexpr int o = decltype(O)::value;
                            mask |= pb_walk_keep(prm, rg, o, fr + a, fr + b) ? (1 << o) : 0;
                        });
                    }
                    mask &= wantbits;
                    // (the outputs of a rolled walk are equally spaced in memory — the X1 terms of the fused stage 1 —
                    // so one address per entry plus a warp-uniform stride per output replaces NOUT pointer look-ups)
                    double* const p0 = outp[0] + (long long)(band - mu_base) * (long long)smu;
                    pb_static_for<0, NOUT>([&](auto O) {
                        constexpr int o = decltype(O)::value;
                        if (mask & (1 << o)) p0[o * out_stride] = ld.stage_get((k - k0) * NOUT + o);
                    });
                }
            };
            int phase = 0;
            for (; s < rg.s_end; ++s) {
                if (s > rg.s_begin) {           // the function that left the span range
                    pb_static_for<0, P1>([&](auto RC) {
                        constexpr int R = decltype(RC)::value;
                        if (phase == R) stage_rot(PbIC<(R + P) % P1>{}, PbIC<0>{});
                    });
                    store_staged(f, 0, P + 1);
                    pb_static_for<0, P1>([&](auto RC) {
                        constexpr int R = decltype(RC)::value;
                        if (phase == R) { stage_rot(PbIC<(R + P) % P1>{}, PbIC<1>{}); clear_entering(RC); }
                    });
                    store_staged(f, P + 1, 2 * P + 1);
                    ++f;
                }
                ld.begin_span(s);
                pb_static_for<0, P1>([&](auto RC) {
                    constexpr int R = decltype(RC)::value;
                    if (phase == R) {
#pragma unroll 1
                        for (int gq = 0; gq < Q; ++gq) node_rolled(RC, gq, s);
                    }
                });
                phase = (phase + 1 == P1) ? 0 : phase + 1;
            }
            // flush: the functions still in the window, starting in the phase of the last span
            const int last_ph = (rg.s_end - rg.s_begin - 1) % P1;
            for (int t = 0; t < P1; ++t) {
                if (f < prm.N) {
                    pb_static_for<0, P1>([&](auto PH) {
                        constexpr int ph = decltype(PH)::value;
                        if ((last_ph + t) % P1 == ph) stage_rot(PH, PbIC<0>{});
                    });
                    store_staged(f, 0, P + 1);
                    pb_static_for<0, P1>([&](auto PH) {
                        constexpr int ph = decltype(PH)::value;
                        if ((last_ph + t) % P1 == ph) stage_rot(PH, PbIC<1>{});
                    });
                    store_staged(f, P + 1, 2 * P + 1);
                }
                ++f;
            }
            return;
        } else {
        while (s < rg.s_end) {
            pb_static_for<0, P1>([&](auto RC) {
                constexpr int R = decltype(RC)::value;
                if (s < rg.s_end) {
                    if (s > rg.s_begin) {       // the function that left the span range: row/column 0 of the previous phase
                        retire_rot(PbIC<(R + P) % P1>{}, f);
                        ++f;
                    }
                    double xc[Q][NOPS];
                    ld.next(s, xc);
                    pb_static_for<0, Q>([&](auto GQ) { node(RC, GQ, xc, s); });
                    ++s;
                }
            });
        }
        }
        // flush: the functions still in the window, starting in the phase of the last span
        const int last_phase = (rg.s_end - rg.s_begin - 1) % P1;
        pb_static_for<0, P1>([&](auto PH) {
            constexpr int ph = decltype(PH)::value;
            if (ph == last_phase) {
                pb_static_for<0, P1>([&](auto T) {
                    constexpr int t = decltype(T)::value;
                    if (f < prm.N) retire_rot(PbIC<(ph + t) % P1>{}, f);
                    ++f;
                });
            }
        });
        return;
    }

    for (int s = rg.s_begin; s < rg.s_end; ++s) {
        const int fs = tb.first[s];
        while (f < fs) { retire_shift(f); ++f; }

        double xc[Q][NOPS];
        ld.next(s, xc);

#pragma unroll
        for (int gq = 0; gq < Q; ++gq) {
            const double* Vn = Vt + (long long)(s * Q + gq) * (2 * P1);
            double D[2][P1];
#pragma unroll
            for (int a = 0; a < P1; ++a) { D[0][a] = Vn[a]; D[1][a] = Vn[P1 + a]; }

            pb_static_for<0, NOUT>([&](auto O) {
                constexpr int o = decltype(O)::value;
                constexpr bool by_fu = pb_group_by_fu<Plan>(o);
                pb_static_for<0, 2>([&](auto FL) {
                    constexpr int fl = decltype(FL)::value;     // the flag shared by the group
                    constexpr int cnt = by_fu ? pb_count_fu<Plan>(o, fl) : pb_count_ft<Plan>(o, fl);
                    if constexpr (cnt > 0) {
                        // y[c] = sum over the group's ops of D_other[c] * x
                        double y[P1];
                        constexpr int lead = pb_first_in_group<Plan>(o, fl, by_fu);
                        pb_static_for<0, NOPS>([&](auto I) {
                            constexpr int i = decltype(I)::value;
                            constexpr PbOp op = Plan::op(i);
                            if constexpr (op.out == o && (by_fu ? op.fu : op.ft) == fl) {
                                constexpr int other = by_fu ? op.ft : op.fu;
                                const double xv = xc[gq][i];
                                if constexpr (i == lead) {
#pragma unroll
                                    for (int c = 0; c < P1; ++c) y[c] = D[other][c] * xv;
                                } else {
#pragma unroll
                                    for (int c = 0; c < P1; ++c) y[c] = fma(D[other][c], xv, y[c]);
                                }
                            }
                        });
                        if constexpr (by_fu) {   // y indexed by the test function a
#pragma unroll
                            for (int a = 0; a < P1; ++a)
#pragma unroll
                                for (int b = 0; b < P1; ++b) acc[o][a][b] = fma(y[a], D[fl][b], acc[o][a][b]);
                        } else {                 // y indexed by the trial function b
#pragma unroll
                            for (int a = 0; a < P1; ++a)
#pragma unroll
                                for (int b = 0; b < P1; ++b) acc[o][a][b] = fma(D[fl][a], y[b], acc[o][a][b]);
                        }
                    }
                });
            });
        }
    }
    // flush: everything still in the window is complete now
#pragma unroll 1
    for (int k = 0; k < P1; ++k) {
        if (f < prm.N) retire_shift(f);
        ++f;
    }
}

#if defined(__CUDACC__)
// ---- TMA bulk copy of the 1D basis table into shared memory ---------------------------------
PB_D uint32_t pb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

PB_D void pb_tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    // one elected thread: arm the barrier with the byte count and launch the bulk copy
    const uint32_t b = pb_smem_u32(bar), d = pb_smem_u32(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
        "l"(gsrc), "r"(bytes), "r"(b)
        : "memory");
}
PB_D void pb_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pb_smem_u32(smem)), "l"(gmem) : "memory");
}
PB_D void pb_cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pb_smem_u32(smem)), "l"(gmem) : "memory");
}
PB_D void pb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PB_D void pb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PB_D void pb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PB_D void pb_mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t b = pb_smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(b),
        "r"(phase)
        : "memory");
}

template <int OFF> PB_D void pb_cp_async8_at(uint32_t smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1], 8;" ::"r"(smem), "l"(gmem), "n"(OFF) : "memory");
}

// spans are issued in order, so the global addresses are running pointers (no multiply per copy)
// and the shared-memory destinations are a stage base plus compile-time offsets
template <class Plan, int Q, int NST, int NTHR = 128>
struct PbAsyncLoader {
    static constexpr int NOPS = Plan::NOPS;
    static constexpr bool ROLLED = false;
    static constexpr int STAGE = Q * NOPS * NTHR;       // doubles per stage of the block ring
    const double* src[NOPS];
    bool has[NOPS];
    long long sc;
    int s_end;
    double* ring;       // this thread's column of the block ring: [NST][Q][NOPS][NTHR]
    int st, s_next;
    uint32_t ring_u32;
    long long scb;
    const char* p[NOPS];
    PB_D void issue(int stage) {
        if (s_next < s_end) {
            const uint32_t d = ring_u32 + (uint32_t)stage * (uint32_t)(STAGE * sizeof(double));
            pb_static_for<0, NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                if (has[i]) {
                    const char* a = p[i];
                    pb_static_for<0, Q>([&](auto GQ) {
                        constexpr int gq = decltype(GQ)::value;
                        pb_cp_async8_at<(gq * NOPS + i) * NTHR * (int)sizeof(double)>(d, a);
                        a += scb;
                    });
                    p[i] = a;
                }
            });
        }
        ++s_next;
        pb_cp_async_commit();
    }
    PB_D void prime(int s_begin) {
        st = 0;
        s_next = s_begin;
        scb = sc * (long long)sizeof(double);
        ring_u32 = pb_smem_u32(ring);
        pb_static_for<0, NOPS>([&](auto I) {
            constexpr int i = decltype(I)::value;
            p[i] = reinterpret_cast<const char*>(src[i]) + (long long)s_begin * Q * scb;
        });
#pragma unroll
        for (int k = 0; k < NST - 1; ++k) issue(k);
    }
    PB_D void next(int, double (&xc)[Q][NOPS]) {
        issue((st + NST - 1) % NST);
        pb_cp_async_wait<NST - 1>();
        const double* r = ring + st * STAGE;
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
            pb_static_for<0, NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                xc[gq][i] = has[i] ? r[(gq * NOPS + i) * NTHR] : 0.0;
            });
        st = (st + 1) % NST;
    }
};

// Generic stage kernel.  Dynamic shared memory holds the walk-axis table slice
// [s_begin*Q, s_end*Q) x 2 x (P+1) doubles when `use_smem` is set; otherwise the table is read
// through the read-only path from global memory (axes too long for 227 KB).
template <class Plan, int P, int Q, int MINB, int NPF>
__global__ void __launch_bounds__(128, MINB) pb_walk_kernel(const __grid_constant__ PbWalkParams prm, int use_smem) {
    extern __shared__ __align__(128) unsigned char pb_smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const double* Vt = prm.V2;
    const PbWalkRange rg = pb_walk_range(prm, blockIdx.y);
    if (use_smem) {
        double* sV = reinterpret_cast<double*>(pb_smem_raw);
        const long long first_node = (long long)rg.s_begin * Q;
        const uint32_t bytes = (uint32_t)((long long)(rg.s_end - rg.s_begin) * Q * 2 * (P + 1) * sizeof(double));
        if (threadIdx.x == 0) {
            pb_mbar_init(&bar, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // bulk copies are limited in size; issue in 32 KB pieces on the same barrier
            const char* g = reinterpret_cast<const char*>(prm.V2 + first_node * 2 * (P + 1));
            char* d = reinterpret_cast<char*>(sV);
            uint32_t done = 0;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_u32(&bar)), "r"(bytes)
                         : "memory");
            while (done < bytes) {
                const uint32_t piece = (bytes - done) < 32768u ? (bytes - done) : 32768u;
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        pb_smem_u32(d + done)),
                    "l"(g + done), "r"(piece), "r"(pb_smem_u32(&bar))
                    : "memory");
                done += piece;
            }
        }
        pb_mbar_wait(&bar, 0);
        Vt = sV - first_node * 2 * (P + 1);     // so that Vt[(s*Q+gq)*2*(P+1)] addresses the slice
    }
    // the small integer tables (first active function per span, retire table) go to shared memory
    // as well: every span of every thread reads them, and with most of the L1 carved out as shared
    // memory they would otherwise be re-fetched from L2 inside the dependent chain of the walk
    const size_t vbytes = use_smem ? ((size_t)(rg.s_end - rg.s_begin) * Q * 2 * (P + 1) * sizeof(double) + 127) & ~size_t(127) : 0;
    const int nsp = rg.s_end - rg.s_begin;
    const int f_lo = rg.f_lo, f_hi = rg.f_hi;                            // functions retired by this walk
    int* s_first = reinterpret_cast<int*>(pb_smem_raw + vbytes);
    int* s_ret = s_first + ((nsp + 3) & ~3);
    const size_t ibytes = (((size_t)((nsp + 3) & ~3) + (size_t)(f_hi - f_lo) * (2 * P + 1)) * sizeof(int) + 127) & ~size_t(127);
    PbWalkTables tb;
    tb.V = Vt;
    if (use_smem) {
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
        __syncthreads();
        tb.first = s_first - rg.s_begin;
        tb.ret_mu = s_ret - (long long)f_lo * (2 * P + 1);
    } else {
        tb.first = prm.first;
        tb.ret_mu = prm.ret_mu;
    }
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < prm.nthreads) {
        if constexpr (NPF >= 2) {
            // NPF doubles as the ring depth of the asynchronous loader
            PbAsyncLoader<Plan, Q, NPF> ld;
            ld.ring = reinterpret_cast<double*>(pb_smem_raw + vbytes + (use_smem ? ibytes : 0)) + threadIdx.x;
            if (use_smem && prm.regular) pb_walk_line_impl<Plan, P, Q, true, true>(prm, rg, tid, tb, ld);
            else if (use_smem) pb_walk_line_impl<Plan, P, Q, true, false>(prm, rg, tid, tb, ld);
            else pb_walk_line_impl<Plan, P, Q, false, false>(prm, rg, tid, tb, ld);
        } else {
            PbRegLoader<Plan, Q> ld;
            if (use_smem) pb_walk_line_impl<Plan, P, Q, true, false>(prm, rg, tid, tb, ld);
            else pb_walk_line_impl<Plan, P, Q, false, false>(prm, rg, tid, tb, ld);
        }
    }
}
#endif

// =================================================================================================
// Final-stage variant: the walk axis is the fastest axis in memory ("lane per span").
//
// When the contracted node axis is contiguous (last tensor axis), the thread-per-line mapping of
// pb_walk_kernel reads one 32-byte sector per lane and load instruction.  Here a WARP owns the line
// instead: lane l takes span sb+l, loads its Q nodes (coalesced across the warp), and accumulates
// the local (P+1)x(P+1) block L[a][b] of the pairs (s+a, s+b).  A band entry collects the blocks
// of up to P+1 neighbouring spans; that sum is done with warp shuffles:
//     out(m, m+d) = sum_{t=0..P-d} L_{span m-t}[t][t+d]          (shfl_up by t)
// Lane l finishes the pairs "owned" by function m = f0 + l (the same retire set as in the walk
// kernel), so ret_mu is reused.  Batches of 32 spans overlap by P spans because the first P lanes
// of a batch lack their left neighbours.  Requires single interior knots (first[s+1] = first[s]+1).
// =================================================================================================

// local block of one span: L[a][b] = sum_gq sum_ops D_ft[a] D_fu[b] x   (single output plan)
template <class Plan, int P, int Q>
PB_HD void pb_span_block(const double (&x)[Q][Plan::NOPS], const double (&D)[Q][2][P + 1], double (&L)[P + 1][P + 1]) {
    constexpr int P1 = P + 1;
    constexpr int NOPS = Plan::NOPS;
    constexpr bool by_fu = pb_group_by_fu<Plan>(0);
#pragma unroll
    for (int a = 0; a < P1; ++a)
#pragma unroll
        for (int b = 0; b < P1; ++b) L[a][b] = 0.0;
#pragma unroll
    for (int gq = 0; gq < Q; ++gq) {
        pb_static_for<0, 2>([&](auto FL) {
            constexpr int fl = decltype(FL)::value;
            constexpr int cnt = by_fu ? pb_count_fu<Plan>(0, fl) : pb_count_ft<Plan>(0, fl);
            if constexpr (cnt > 0) {
                double y[P1];
                constexpr int lead = pb_first_in_group<Plan>(0, fl, by_fu);
                pb_static_for<0, NOPS>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr PbOp op = Plan::op(i);
                    if constexpr ((by_fu ? op.fu : op.ft) == fl) {
                        constexpr int other = by_fu ? op.ft : op.fu;
                        const double xv = x[gq][i];
                        if constexpr (i == lead) {
#pragma unroll
                            for (int c = 0; c < P1; ++c) y[c] = D[gq][other][c] * xv;
                        } else {
#pragma unroll
                            for (int c = 0; c < P1; ++c) y[c] = fma(D[gq][other][c], xv, y[c]);
                        }
                    }
                });
                if constexpr (by_fu) {
#pragma unroll
                    for (int a = 0; a < P1; ++a)
#pragma unroll
                        for (int b = 0; b < P1; ++b) L[a][b] = fma(y[a], D[gq][fl][b], L[a][b]);
                } else {
#pragma unroll
                    for (int a = 0; a < P1; ++a)
#pragma unroll
                        for (int b = 0; b < P1; ++b) L[a][b] = fma(D[gq][fl][a], y[b], L[a][b]);
                }
            }
        });
    }
}

// line decode shared by the warp kernel and its emulation
struct PbLineOffsets { long long in, in_tr, out, out_tr; bool keep, mirror; };
template <class Plan>
PB_HD PbLineOffsets pb_decode_line(const PbWalkParams& prm, long long line) {
    PbLineOffsets o;
    const int x = (int)(line % prm.X);
    const long long t2 = line / prm.X;
    const int v = (int)(t2 % prm.V);
    const int u = (int)(t2 / prm.V) + prm.u_begin;
    o.keep = true;
    o.mirror = false;
    if (prm.u_pair_i != nullptr) {
        const int ui = prm.u_pair_i[u], uj = prm.u_pair_j[u];
        o.keep = pb_keep(prm.u_mode[0], ui, uj, prm.u_lo, prm.u_hi);
        if (prm.mirror) {
            // symmetric form: of a line and its transposed partner only the "upper" one is computed
            // (if the partner row belongs to this slab) and its result is written to both
            const bool partner_owned = (uj >= prm.u_lo && uj < prm.u_hi);
            int vi = 0, vj = 0;
            if (prm.v_pair_i != nullptr) { vi = prm.v_pair_i[v]; vj = prm.v_pair_j[v]; }
            const bool upper = ui < uj || (ui == uj && vi <= vj);
            const bool self = (ui == uj && vi == vj);
            o.keep = o.keep && (!partner_owned || upper);
            o.mirror = partner_owned && upper && !self;
        }
    }
    o.in = (long long)(u - prm.u_base_in) * prm.in_su + (long long)v * prm.in_sv + (long long)x * prm.in_sx;
    o.in_tr = o.in;
    if (Plan::HAS_TR) {
        const int ut = prm.tr_u ? prm.tr_u[u] : u;
        const int vt = prm.tr_v ? prm.tr_v[v] : v;
        o.in_tr = (long long)(ut - prm.u_base_in) * prm.in_su + (long long)vt * prm.in_sv + (long long)x * prm.in_sx;
    }
    o.out = (long long)(u - prm.u_base_out) * prm.out_su + (long long)v * prm.out_sv + (long long)x * prm.out_sx;
    o.out_tr = o.out;
    if (o.mirror) {
        const int ut = prm.tr_u ? prm.tr_u[u] : u;
        const int vt = prm.tr_v ? prm.tr_v[v] : v;
        o.out_tr = (long long)(ut - prm.u_base_out) * prm.out_su + (long long)vt * prm.out_sv + (long long)x * prm.out_sx;
    }
    return o;
}

PB_HD int pb_lane_batches(int nspans, int P) {
    const int rest = nspans + P - 32;
    return 1 + (rest > 0 ? (rest + (32 - P) - 1) / (32 - P) : 0);
}

// sequential emulation of one (line, batch) of the warp kernel
template <class Plan, int P, int Q>
PB_HD void pb_lane_span_seq(const PbWalkParams& prm, long long line, int batch) {
    constexpr int P1 = P + 1, NOPS = Plan::NOPS;
    const PbLineOffsets lo = pb_decode_line<Plan>(prm, line);
    if (!lo.keep) return;
    const int sb = prm.s_begin + batch * (32 - P);
    const int f0 = prm.first[prm.s_begin];
    double Ls[32][P1][P1];
    for (int lane = 0; lane < 32; ++lane) {
        const int s = sb + lane;
        double x[Q][NOPS], D[Q][2][P1];
        for (int gq = 0; gq < Q; ++gq) {
            for (int i = 0; i < NOPS; ++i) x[gq][i] = 0.0;
            for (int a = 0; a < P1; ++a) D[gq][0][a] = D[gq][1][a] = 0.0;
        }
        if (s < prm.s_end) {
            for (int gq = 0; gq < Q; ++gq) {
                pb_static_for<0, NOPS>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    x[gq][i] = prm.in[i] ? prm.in[i][(Plan::op(i).tr ? lo.in_tr : lo.in) + (long long)(s * Q + gq)] : 0.0;
                });
                const double* Vn = prm.V2 + (long long)(s * Q + gq) * 2 * P1;
                for (int a = 0; a < P1; ++a) { D[gq][0][a] = Vn[a]; D[gq][1][a] = Vn[P1 + a]; }
            }
        }
        pb_span_block<Plan, P, Q>(x, D, Ls[lane]);
    }
    for (int lane = 0; lane < 32; ++lane) {
        if (batch > 0 && lane < P) continue;
        const int m = f0 + (sb - prm.s_begin) + lane;
        if (m >= prm.N) continue;
        const int* rm = prm.ret_mu + (long long)m * (2 * P + 1);
        for (int k = 0; k <= 2 * P; ++k) {
            if (rm[k] < 0) continue;
            const int d = (k <= P) ? k : k - P;
            double sum = 0.0;
            for (int t = 0; t <= P - d; ++t)
                if (lane - t >= 0) sum += (k <= P) ? Ls[lane - t][t][t + d] : Ls[lane - t][t + d][t];
            prm.out[0][lo.out + (long long)(rm[k] - prm.mu_base)] = sum;
            if (lo.mirror) {        // entry (i,j) of this line is entry (j,i) of the transposed line
                const int kt = (k == 0) ? 0 : (k <= P ? k + P : k - P);
                prm.out[0][lo.out_tr + (long long)(rm[kt] - prm.mu_base)] = sum;
            }
        }
    }
}

#if defined(__CUDACC__)
// grid.x: groups of `lines_per_warp` lines, 4 warps per block; grid.y: batch
template <class Plan, int P, int Q>
__global__ void __launch_bounds__(128) pb_lane_span_kernel(const __grid_constant__ PbWalkParams prm, int lines_per_warp) {
    constexpr int P1 = P + 1, NOPS = Plan::NOPS;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int batch = blockIdx.y;
    const long long line0 = warp * lines_per_warp;
    if (line0 >= prm.nthreads) return;
    const long long line1 = (line0 + lines_per_warp < prm.nthreads) ? line0 + lines_per_warp : prm.nthreads;

    const int sb = prm.s_begin + batch * (32 - P);
    const int s = sb + lane;
    const bool active = s < prm.s_end;
    const int m = prm.first[prm.s_begin] + (sb - prm.s_begin) + lane;
    const bool writer = (batch == 0 || lane >= P) && m < prm.N;

    double D[Q][2][P1];
#pragma unroll
    for (int gq = 0; gq < Q; ++gq) {
        const double* Vn = prm.V2 + (long long)((active ? s : prm.s_begin) * Q + gq) * 2 * P1;
#pragma unroll
        for (int a = 0; a < P1; ++a) {
            D[gq][0][a] = active ? __ldg(Vn + a) : 0.0;
            D[gq][1][a] = active ? __ldg(Vn + P1 + a) : 0.0;
        }
    }
    int mu[2 * P + 1];
#pragma unroll
    for (int k = 0; k <= 2 * P; ++k) mu[k] = writer ? __ldg(prm.ret_mu + (long long)m * (2 * P + 1) + k) : -1;

    const long long node0 = (long long)(active ? s : prm.s_begin) * Q;
    double xn[Q][NOPS];
    auto load_line = [&](long long line, PbLineOffsets& lo) {
        lo = pb_decode_line<Plan>(prm, line);
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
            pb_static_for<0, NOPS>([&](auto I) {
                constexpr int i = decltype(I)::value;
                xn[gq][i] = (active && lo.keep && prm.in[i]) ? prm.in[i][(Plan::op(i).tr ? lo.in_tr : lo.in) + node0 + gq] : 0.0;
            });
    };
    PbLineOffsets lo_next;
    load_line(line0, lo_next);
    for (long long line = line0; line < line1; ++line) {
        const PbLineOffsets lo = lo_next;
        double x[Q][NOPS];
#pragma unroll
        for (int gq = 0; gq < Q; ++gq)
            pb_static_for<0, NOPS>([&](auto I) { constexpr int i = decltype(I)::value; x[gq][i] = xn[gq][i]; });
        if (line + 1 < line1) load_line(line + 1, lo_next);
        if (!lo.keep) continue;             // warp-uniform
        double L[P1][P1];
        pb_span_block<Plan, P, Q>(x, D, L);
        // gather the blocks of the left neighbours
#pragma unroll
        for (int k = 0; k <= 2 * P; ++k) {
            const int d = (k <= P) ? k : k - P;
            double sum = (k <= P) ? L[0][d] : L[d][0];
#pragma unroll
            for (int t = 1; t <= P; ++t) {
                if (t <= P - d) {
                    const double vsh = __shfl_up_sync(0xffffffffu, (k <= P) ? L[t][t + d] : L[t + d][t], t);
                    if (lane >= t) sum += vsh;
                }
            }
            if (mu[k] >= 0) prm.out[0][lo.out + (long long)(mu[k] - prm.mu_base)] = sum;
        }
    }
}
// -------------------------------------------------------------------------------------------------
// v2 of the warp-per-line kernel: asynchronous staging through shared memory.
//   * loads: every line segment (32 spans x Q nodes per input term) is copied global -> shared with
//     cp.async in 16-byte pieces (8-byte for odd Q) that are contiguous across the warp, NST line
//     segments ahead of their use; no registers are tied up by data in flight.  For Q = 4 the
//     16-byte pieces are XOR-swizzled so that the per-lane 32-byte reads are bank-conflict free.
//   * stores: the 2P+1 finished entries of every lane are first placed at their band offset in a
//     per-warp shared buffer and then written by consecutive lanes, i.e. as full 32-byte sectors.
// -------------------------------------------------------------------------------------------------

template <int P, int Q> struct PbLaneCfg {
    static constexpr int SEG = 32 * Q;                      // doubles per term and line segment
    static constexpr int OUTSLOTS = (32 + P) * (2 * P + 1); // staging slots for finished entries
    static constexpr int OUTPAD = (OUTSLOTS + 31) / 32 * 32;
    static constexpr int DQ = 64;           // per warp: decoded lines (<= lines per warp)
    static constexpr int LOSLOTS = 4 * DQ;  // in, in_tr, out, out_tr offsets of the decoded lines
};

template <class Plan, int P, int Q, int NST>
__global__ void __launch_bounds__(128, (P >= 4 ? 2 : 3)) pb_lane_span_kernel_v2(const __grid_constant__ PbWalkParams prm, int lines_per_warp) {
    constexpr int P1 = P + 1, NOPS = Plan::NOPS;
    using Cfg = PbLaneCfg<P, Q>;
    constexpr int SEG = Cfg::SEG;
    constexpr bool VEC = (Q % 2 == 0);                      // 16-byte pieces
    constexpr int STAGE = NOPS * SEG;                       // doubles per pipeline stage
    extern __shared__ __align__(16) double pb_lane_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* ring = pb_lane_smem + (size_t)wib * (NST * STAGE + Cfg::OUTPAD + Cfg::LOSLOTS);
    double* obuf = ring + NST * STAGE;
    long long* dq = reinterpret_cast<long long*>(obuf + Cfg::OUTPAD);

    const long long warp = (long long)blockIdx.x * 4 + wib;
    const int batch = blockIdx.y;
    const long long line0 = warp * lines_per_warp;
    if (line0 >= prm.nthreads) return;
    const long long line1 = (line0 + lines_per_warp < prm.nthreads) ? line0 + lines_per_warp : prm.nthreads;

    const int sb = prm.s_begin + batch * (32 - P);
    const int s = sb + lane;
    const bool active = s < prm.s_end;
    const int m = prm.first[prm.s_begin] + (sb - prm.s_begin) + lane;
    const bool writer = (batch == 0 || lane >= P) && m < prm.N;
    const long long seg_node0 = (long long)sb * Q;                       // first node of the segment
    const long long seg_nodes = (long long)(pb_min(prm.s_end, sb + 32) - sb) * Q;   // valid nodes in it

    double D[Q][2][P1];
#pragma unroll
    for (int gq = 0; gq < Q; ++gq) {
        const double* Vn = prm.V2 + (long long)((active ? s : prm.s_begin) * Q + gq) * 2 * P1;
#pragma unroll
        for (int a = 0; a < P1; ++a) {
            D[gq][0][a] = active ? __ldg(Vn + a) : 0.0;
            D[gq][1][a] = active ? __ldg(Vn + P1 + a) : 0.0;
        }
    }
    // band offsets of this lane's entries relative to the smallest one written by the warp
    int mu[2 * P + 1];
    int mu_lo = 0x7fffffff, mu_hi = -1;
#pragma unroll
    for (int k = 0; k <= 2 * P; ++k) {
        mu[k] = writer ? __ldg(prm.ret_mu + (long long)m * (2 * P + 1) + k) : -1;
        if (mu[k] >= 0) { mu_lo = pb_min(mu_lo, mu[k]); mu_hi = pb_max(mu_hi, mu[k]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mu_lo = pb_min(mu_lo, __shfl_xor_sync(0xffffffffu, mu_lo, o));
        mu_hi = pb_max(mu_hi, __shfl_xor_sync(0xffffffffu, mu_hi, o));
    }
    if (mu_hi < 0) return;                                              // nothing to write in this batch
    const int nslots = mu_hi - mu_lo + 1;                               // <= OUTSLOTS
    // which staging slots does this warp own?  (bit j of `mine`: slot lane + 32 j)
    for (int t = lane; t < Cfg::OUTPAD; t += 32) obuf[t] = 0.0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k <= 2 * P; ++k)
        if (mu[k] >= 0) obuf[mu[k] - mu_lo] = 1.0;
    __syncwarp();
    unsigned mine = 0;
#pragma unroll
    for (int j = 0; j < Cfg::OUTPAD / 32; ++j)
        if (obuf[lane + 32 * j] != 0.0) mine |= 1u << j;
    __syncwarp();

    // regions of the ring that no copy ever writes (spans past the end of the axis, absent terms)
    // must read as zero
    for (int t = lane; t < NST * STAGE; t += 32) ring[t] = 0.0;
    __syncwarp();

    // asynchronous copy of one line segment into ring stage `st`
    // this lane's share of a line segment: piece index -> offsets in global / shared memory
    constexpr int NPIECE = VEC ? (SEG / 2 + 31) / 32 : (SEG + 31) / 32;
    int goff[NPIECE], soff[NPIECE];
#pragma unroll
    for (int h = 0; h < NPIECE; ++h) {
        const int c = lane + 32 * h;
        if constexpr (VEC) {
            const int pc = (Q == 4) ? (c ^ ((c >> 3) & 1)) : c;
            goff[h] = (c < SEG / 2 && 2 * c < seg_nodes) ? 2 * c : -1;
            soff[h] = 2 * pc;
        } else {
            goff[h] = (c < SEG && c < seg_nodes) ? c : -1;
            soff[h] = c;
        }
    }
    // Lines of this warp: decoded 32 at a time, one line per lane, so that the table look-ups of the
    // decode (pair tables, transposed indices) cost one memory latency per 32 lines instead of a
    // dependent chain per line; lines that are filtered out (slab filter, mirrored half of a
    // symmetric form) are dropped here and never enter the copy pipeline.
    int nkept = 0;
    for (long long base = line0; base < line1; base += 32) {
        const long long ln = base + lane;
        PbLineOffsets lo;
        lo.keep = false;
        if (ln < line1) lo = pb_decode_line<Plan>(prm, ln);
        const unsigned bal = __ballot_sync(0xffffffffu, lo.keep);
        if (lo.keep) {
            const int slot = nkept + __popc(bal & ((1u << lane) - 1u));
            dq[slot] = lo.in;
            dq[Cfg::DQ + slot] = lo.in_tr;
            dq[2 * Cfg::DQ + slot] = lo.out;
            dq[3 * Cfg::DQ + slot] = lo.mirror ? lo.out_tr : -1;
        }
        nkept += __popc(bal);
    }
    __syncwarp();
    if (nkept == 0) return;

    auto issue = [&](int k, int st) {
        const long long in = dq[k], in_tr = dq[Cfg::DQ + k];
        double* dst = ring + (size_t)st * STAGE;
        pb_static_for<0, NOPS>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if (prm.in[i]) {
                const double* src = prm.in[i] + (Plan::op(i).tr ? in_tr : in) + seg_node0;
#pragma unroll
                for (int h = 0; h < NPIECE; ++h) {
                    if (goff[h] >= 0) {
                        if constexpr (VEC) pb_cp_async16(dst + i * SEG + soff[h], src + goff[h]);
                        else pb_cp_async8(dst + i * SEG + soff[h], src + goff[h]);
                    }
                }
            }
        });
        pb_cp_async_commit();
    };

#pragma unroll
    for (int j = 0; j < NST - 1; ++j) {
        if (j < nkept) issue(j, j);
        else pb_cp_async_commit();
    }
    int st = 0;
    for (int k = 0; k < nkept; ++k) {
        __syncwarp();                                   // everyone is done with the stage refilled below
        if (k + NST - 1 < nkept) issue(k + NST - 1, (st + NST - 1) % NST);
        else pb_cp_async_commit();
        pb_cp_async_wait<NST - 1>();
        __syncwarp();
        const long long out = dq[2 * Cfg::DQ + k], out_tr = dq[3 * Cfg::DQ + k];
        const double* src = ring + (size_t)st * STAGE;
        double x[Q][NOPS];
        pb_static_for<0, NOPS>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (VEC) {
#pragma unroll
                for (int h = 0; h < Q / 2; ++h) {
                    const int c = lane * (Q / 2) + h;
                    const int pc = (Q == 4) ? (c ^ ((c >> 3) & 1)) : c;
                    const double2 v = *reinterpret_cast<const double2*>(src + i * SEG + 2 * pc);
                    x[2 * h][i] = v.x;
                    x[2 * h + 1][i] = v.y;
                }
            } else {
#pragma unroll
                for (int gq = 0; gq < Q; ++gq) x[gq][i] = src[i * SEG + lane * Q + gq];
            }
        });
        double L[P1][P1];
        pb_span_block<Plan, P, Q>(x, D, L);
        double val[2 * P + 1];
#pragma unroll
        for (int kk = 0; kk <= 2 * P; ++kk) {
            const int d = (kk <= P) ? kk : kk - P;
            double sum = (kk <= P) ? L[0][d] : L[d][0];
#pragma unroll
            for (int t = 1; t <= P; ++t) {
                if (t <= P - d) {
                    const double vsh = __shfl_up_sync(0xffffffffu, (kk <= P) ? L[t][t + d] : L[t + d][t], t);
                    if (lane >= t) sum += vsh;
                }
            }
            val[kk] = sum;
            if (mu[kk] >= 0) obuf[mu[kk] - mu_lo] = sum;
        }
        __syncwarp();
        double* dst = prm.out[0] + out + (long long)(mu_lo - prm.mu_base);
#pragma unroll
        for (int j = 0; j < Cfg::OUTPAD / 32; ++j)
            if ((mine >> j) & 1u) dst[lane + 32 * j] = obuf[lane + 32 * j];
        if (out_tr >= 0) {                              // warp-uniform
            // entry (i,j) of this line is entry (j,i) of the transposed line: same staging
            // slots, values swapped between the pair and its transpose
            __syncwarp();
#pragma unroll
            for (int kk = 0; kk <= 2 * P; ++kk) {
                const int kt = (kk == 0) ? 0 : (kk <= P ? kk + P : kk - P);
                if (mu[kk] >= 0) obuf[mu[kk] - mu_lo] = val[kt];
            }
            __syncwarp();
            double* dst_t = prm.out[0] + out_tr + (long long)(mu_lo - prm.mu_base);
#pragma unroll
            for (int j = 0; j < Cfg::OUTPAD / 32; ++j)
                if ((mine >> j) & 1u) dst_t[lane + 32 * j] = obuf[lane + 32 * j];
        }
        st = (st + 1) % NST;
    }
    (void)nslots;
}
#endif
