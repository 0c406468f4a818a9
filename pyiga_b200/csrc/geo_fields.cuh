// K2 — geometry Jacobian and per-Gauss-point coefficient fields, fused.
//
// Replaces, in one kernel, the reference's
//   * `BSplineFunc.grid_eval / grid_jacobian` (pyiga/bspline.py:874-921): d (+1) sparse mode
//     products of the control net with collocation matrices via `apply_tprod`
//     (pyiga/tensor.py:97-128),
//   * the NURBS quotient rule `_nurbs_jacobian` (pyiga/geometry.py:17-25, 116-123),
//   * `precompute_fields` of the assembler classes (pyiga/assemblers.pyx:1389-1449 stiffness 3D,
//     :1223-1249 mass 3D, :234-275 / :86-110 in 2D).
// One thread per Gauss point (lanes along the last grid axis).  The Jacobian is built from the
// compact 1D geometry tables (values / derivatives of the p_g+1 active geometry basis functions at
// every node) and never stored; only the fields the form needs go to HBM, in SoA layout
//     F[c][g0][g1][g2]
// so that the contraction kernels read them coalesced.
//
// Conventions (same as the reference): J[i][j] = d geo_i / d xi_j, with xi_0 the *last* tensor axis.
#pragma once
#include "common.cuh"

struct PbGeoDev {
    int sdim, dim;          // parameter / physical dimension
    int nc;                 // stored components = dim (+1 if rational: premultiplied coords + weight)
    int rational;
    int pg[PB_MAXDIM];      // degrees of the geometry basis
    int Ng[PB_MAXDIM];      // control net size
    const int* gfirst[PB_MAXDIM];   // [G_k]  first active geometry function at node g_k
    const double* GV[PB_MAXDIM];    // [G_k][2][pg_k+1]
    const double* coeffs;           // [Ng0][Ng1][Ng2][nc]
};

struct PbFieldParams {
    int dim;
    int G[PB_MAXDIM];
    const double* gw[PB_MAXDIM];    // Gauss weights per axis
    PbGeoDev geo;
    const double* jac_in;           // optional: Jacobians evaluated by the caller, [pts][dim][dim]
    const double* val_in;           // optional: geometry values, [pts][dim]
    double* fields;                 // [nf][pts]
    long long npts;                 // points of the whole grid (stride between fields)
    long long pt_begin, pt_end;     // linear range of points evaluated by this launch
    int nf;
    // general first-order scalar forms (PbProgGeneral): physical coefficient terms and output map
    const double* inputs[PB_MAXPHYS];  // coefficient arrays on the Gauss grid (null: constant 1)
    int nphys;
    struct { int bt, bu, input; double scale; } phys[PB_MAXPHYS];   // slots: 0 value, 1+a d/dx_a
    struct { int bp, ap; } outmap[PB_MAXPHYS];                      // parametric slots of field f
};

template <int I> struct PbInt { static constexpr int value = I; };

struct PbPoint {        // what a field program sees at one Gauss point
    double J[3][3];     // Jacobian numerators: the Jacobian is J / jden
    double jden;        // 1 unless the producer leaves the quotient rule's 1/W^2 to the program (RAT)
    double x[3];        // physical coordinates
    double gw;          // product of the 1D Gauss weights
    long long idx;      // linear point index
};

// geometry value and Jacobian at the Gauss point (g0,g1,g2)
template <int DIM>
PB_HD void pb_geo_eval(const PbGeoDev& geo, const int* g, PbPoint& pt) {
    double val[4] = {0, 0, 0, 0};
    double dv[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // dv[c][k]: derivative along tensor axis k
    const int nc = geo.nc;
    const int f0 = geo.gfirst[0][g[0]], f1 = geo.gfirst[1][g[1]];
    const double* T0 = geo.GV[0] + (long long)g[0] * 2 * (geo.pg[0] + 1);
    const double* T1 = geo.GV[1] + (long long)g[1] * 2 * (geo.pg[1] + 1);
    if constexpr (DIM == 2) {
        for (int a0 = 0; a0 <= geo.pg[0]; ++a0) {
            const double w0 = T0[a0], d0 = T0[geo.pg[0] + 1 + a0];
            double sv[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
            for (int a1 = 0; a1 <= geo.pg[1]; ++a1) {
                const double w1 = T1[a1], d1 = T1[geo.pg[1] + 1 + a1];
                const double* c = geo.coeffs + ((long long)(f0 + a0) * geo.Ng[1] + (f1 + a1)) * nc;
                for (int k = 0; k < nc; ++k) { sv[k] = fma(c[k], w1, sv[k]); sd[k] = fma(c[k], d1, sd[k]); }
            }
            for (int k = 0; k < nc; ++k) {
                val[k] = fma(w0, sv[k], val[k]);
                dv[k][1] = fma(w0, sd[k], dv[k][1]);
                dv[k][0] = fma(d0, sv[k], dv[k][0]);
            }
        }
    } else {
        const int f2 = geo.gfirst[2][g[2]];
        const double* T2 = geo.GV[2] + (long long)g[2] * 2 * (geo.pg[2] + 1);
        for (int a0 = 0; a0 <= geo.pg[0]; ++a0) {
            const double w0 = T0[a0], d0 = T0[geo.pg[0] + 1 + a0];
            for (int a1 = 0; a1 <= geo.pg[1]; ++a1) {
                const double w1 = T1[a1], d1 = T1[geo.pg[1] + 1 + a1];
                double sv[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
                const double* c = geo.coeffs + (((long long)(f0 + a0) * geo.Ng[1] + (f1 + a1)) * geo.Ng[2] + f2) * nc;
                for (int a2 = 0; a2 <= geo.pg[2]; ++a2) {
                    const double w2 = T2[a2], d2 = T2[geo.pg[2] + 1 + a2];
                    for (int k = 0; k < nc; ++k) {
                        sv[k] = fma(c[a2 * nc + k], w2, sv[k]);
                        sd[k] = fma(c[a2 * nc + k], d2, sd[k]);
                    }
                }
                const double w01 = w0 * w1, w0d1 = w0 * d1, d0w1 = d0 * w1;
                for (int k = 0; k < nc; ++k) {
                    val[k] = fma(w01, sv[k], val[k]);
                    dv[k][2] = fma(w01, sd[k], dv[k][2]);
                    dv[k][1] = fma(w0d1, sv[k], dv[k][1]);
                    dv[k][0] = fma(d0w1, sv[k], dv[k][0]);
                }
            }
        }
    }
    // J[i][j], j = DIM-1-k  (x is the last tensor axis)
    if (geo.rational) {
        const double W = val[geo.dim];
        const double iW2 = 1.0 / (W * W);
        for (int i = 0; i < geo.dim; ++i) {
            pt.x[i] = val[i] / W;
            for (int k = 0; k < DIM; ++k)
                pt.J[i][DIM - 1 - k] = (dv[i][k] * W - val[i] * dv[geo.dim][k]) * iW2;
        }
    } else {
        for (int i = 0; i < geo.dim; ++i) {
            pt.x[i] = val[i];
            for (int k = 0; k < DIM; ++k) pt.J[i][DIM - 1 - k] = dv[i][k];
        }
    }
}

// 1 / x for the field programs of the fused stage 1: the hardware's reciprocal seed (20+ bits) and three
// Newton steps — full double precision up to the last bit or two (the IEEE division's exact rounding
// costs a ~90-cycle dependent chain per Gauss point, which the two resident warps per scheduler of that
// kernel cannot hide).  x is a positive finite number here (|det J| of a regular map).
PB_HD double pb_rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// ---- field programs ---------------------------------------------------------------------------
template <int DIM> PB_HD double pb_det(const double (&J)[3][3]) {
    if constexpr (DIM == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
    else return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
              - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
              + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

// inverse by cofactors, like the generated reference code (pyiga/assemblers.pyx:1430-1441)
template <int DIM> PB_HD void pb_inv(const double (&J)[3][3], double det, double (&I)[3][3]) {
    const double t = 1.0 / det;
    if constexpr (DIM == 2) {
        I[0][0] = t * J[1][1];  I[0][1] = -t * J[0][1];
        I[1][0] = -t * J[1][0]; I[1][1] = t * J[0][0];
    } else {
        I[0][0] = t * (J[1][1] * J[2][2] - J[1][2] * J[2][1]);
        I[0][1] = -t * (J[0][1] * J[2][2] - J[0][2] * J[2][1]);
        I[0][2] = t * (J[0][1] * J[1][2] - J[0][2] * J[1][1]);
        I[1][0] = -t * (J[1][0] * J[2][2] - J[1][2] * J[2][0]);
        I[1][1] = t * (J[0][0] * J[2][2] - J[0][2] * J[2][0]);
        I[1][2] = -t * (J[0][0] * J[1][2] - J[0][2] * J[1][0]);
        I[2][0] = t * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        I[2][1] = -t * (J[0][0] * J[2][1] - J[0][1] * J[2][0]);
        I[2][2] = t * (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
    }
}

// adjugate (transposed cofactors): inverse = A / det
template <int DIM> PB_HD void pb_adj(const double (&J)[3][3], double (&A)[3][3]) {
    if constexpr (DIM == 2) {
        A[0][0] = J[1][1];  A[0][1] = -J[0][1];
        A[1][0] = -J[1][0]; A[1][1] = J[0][0];
    } else {
        A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
        A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
        A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
        A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    }
}

// programs are instantiated with RAT = true where the point carries an unnormalised Jacobian
// (J / jden, rational geometries on the row path) and fold the denominator into their own scaling
template <int DIM, bool RAT> PB_HD void pb_point_normalize(PbPoint& pt) {
    if constexpr (RAT) {
        const double r = 1.0 / pt.jden;
        for (int i = 0; i < DIM; ++i)
            for (int j = 0; j < DIM; ++j) pt.J[i][j] *= r;
        pt.jden = 1.0;
    }
}

// mass: W = GaussWeight * |det J|                      (pyiga/assemblers.pyx:1223-1249)
template <int DIM> struct PbProgMass {
    static constexpr int NF = 1;
    static constexpr bool NEED_X = false;
    template <bool RAT> PB_HD static void run(const PbFieldParams&, PbPoint& pt, double* f) { point<RAT>(pt, f); }
    template <bool RAT, bool FAST = false> PB_HD static void point(PbPoint& pt, double* f) {
        const double d = pt.gw * fabs(pb_det<DIM>(pt.J));
        if constexpr (RAT) {
            const double r = FAST ? pb_rcp_fast(pt.jden) : 1.0 / pt.jden;
            f[0] = d * (DIM == 2 ? r * r : r * r * r);
        } else {
            f[0] = d;
        }
    }
};

// stiffness: B = W * J^-1 J^-T, symmetric-packed upper triangle, x,y,z index order
//                                                      (pyiga/assemblers.pyx:1389-1449, vform.py:28-34)
template <int DIM> struct PbProgStiffness {
    static constexpr int NF = DIM * (DIM + 1) / 2;
    static constexpr bool NEED_X = false;
    template <bool RAT> PB_HD static void run(const PbFieldParams&, PbPoint& pt, double* f) { point<RAT>(pt, f); }
    template <bool RAT, bool FAST = false> PB_HD static void point(PbPoint& pt, double* f) {
        // with J = N / jden:  W J^-1 J^-T = gw jden^(2-DIM) / |det N| * adj(N) adj(N)^T  -- one division
        double A[3][3];
        pb_adj<DIM>(pt.J, A);
        // Laplace expansion along the first row with the cofactors that are already there
        double det = pt.J[0][0] * A[0][0];
        for (int m = 1; m < DIM; ++m) det = fma(pt.J[0][m], A[m][0], det);
        const double den = (RAT && DIM == 3) ? pt.jden * fabs(det) : fabs(det);
        const double sc = FAST ? pt.gw * pb_rcp_fast(den) : pt.gw / den;
        int k = 0;
        for (int a = 0; a < DIM; ++a)
            for (int b = a; b < DIM; ++b) {
                double s = A[a][0] * A[b][0];
                for (int m = 1; m < DIM; ++m) s = fma(A[a][m], A[b][m], s);
                f[k++] = sc * s;
            }
    }
};

// General scalar bilinear form with at most first derivatives on u and v:
//     a(u,v) = int sum_t c_t(x) d^{bt} v d^{bu} u dx,   slots: 0 = value, 1+a = d/dx_a (physical)
// pulled back to the parameter domain: d/dx_a = sum_j Jinv[j][a] d/dxi_j, xi_j = tensor axis DIM-1-j.
// Output field f holds the coefficient of the parametric slot pair outmap[f]:
//     C[bp][ap] = W * sum_t c_t T[bt][bp] T[bu][ap],  T[0][0] = 1,  T[1+a][1+k] = Jinv[DIM-1-k][a].
// This is what the reference's generated precompute_fields computes for such forms
// (pyiga/codegen/cython.py:673-701 on the finalized VForm, pyiga/vform.py:705-731).
template <int DIM> struct PbProgGeneral {
    static constexpr int NF = PB_MAXPHYS;
    static constexpr bool NEED_X = false;
    template <bool RAT> PB_HD static void run(const PbFieldParams& prm, PbPoint& pt, double* f) {
        pb_point_normalize<DIM, RAT>(pt);
        const double det = pb_det<DIM>(pt.J);
        const double W = pt.gw * fabs(det);
        double I[3][3];
        pb_inv<DIM>(pt.J, det, I);
        double T[DIM + 1][DIM + 1];
        for (int a = 0; a <= DIM; ++a)
            for (int b = 0; b <= DIM; ++b) T[a][b] = 0.0;
        T[0][0] = 1.0;
        for (int a = 0; a < DIM; ++a)
            for (int k = 0; k < DIM; ++k) T[1 + a][1 + k] = I[DIM - 1 - k][a];
        for (int c = 0; c < prm.nf; ++c) f[c] = 0.0;
        for (int t = 0; t < prm.nphys; ++t) {
            const int in = prm.phys[t].input;
            const double cv = W * prm.phys[t].scale * (in >= 0 ? prm.inputs[in][pt.idx] : 1.0);
            for (int c = 0; c < prm.nf; ++c) {
                // linear forms have no trial function: slot -1 on both sides
                const double tu = prm.outmap[c].ap < 0 ? 1.0 : T[prm.phys[t].bu][prm.outmap[c].ap];
                f[c] = fma(cv, T[prm.phys[t].bt][prm.outmap[c].bp] * tu, f[c]);
            }
        }
    }
};

// raw geometry data (debug / host callbacks): J row-major (DIM*DIM), then x (DIM)
template <int DIM> struct PbProgGeoRaw {
    static constexpr int NF = DIM * DIM + DIM;
    static constexpr bool NEED_X = true;
    template <bool RAT> PB_HD static void run(const PbFieldParams&, PbPoint& pt, double* f) {
        pb_point_normalize<DIM, RAT>(pt);
        for (int i = 0; i < DIM; ++i)
            for (int j = 0; j < DIM; ++j) f[i * DIM + j] = pt.J[i][j];
        for (int i = 0; i < DIM; ++i) f[DIM * DIM + i] = pt.x[i];
    }
};

template <int DIM, class Prog>
PB_HD void pb_fields_point(const PbFieldParams& prm, long long idx) {
    int g[3];
    long long r = idx;
    for (int k = DIM - 1; k >= 0; --k) { g[k] = (int)(r % prm.G[k]); r /= prm.G[k]; }
    PbPoint pt;
    pt.idx = idx;
    pt.jden = 1.0;
    pt.gw = 1.0;
    for (int k = 0; k < DIM; ++k) pt.gw *= prm.gw[k][g[k]];
    if (prm.jac_in) {
        for (int i = 0; i < DIM; ++i)
            for (int j = 0; j < DIM; ++j) pt.J[i][j] = prm.jac_in[(idx * DIM + i) * DIM + j];
        for (int i = 0; i < DIM; ++i) pt.x[i] = prm.val_in ? prm.val_in[idx * DIM + i] : 0.0;
    } else {
        pb_geo_eval<DIM>(prm.geo, g, pt);
    }
    double f[Prog::NF];
    Prog::template run<false>(prm, pt, f);
    for (int c = 0; c < Prog::NF && c < prm.nf; ++c) prm.fields[(long long)c * prm.npts + idx] = f[c];
}

// ---- row-wise evaluation (the production path for spline geometries) --------------------------
// All points of one grid row (fixed g0 [,g1]; the last axis runs) share the contraction of the
// control net with the basis functions of the leading axes.  A block first reduces the net to
//     Y[i_last][c][v] = sum_{a0[,a1]} coeffs[f0+a0][f1+a1][i_last][c] * w_v(a0, a1)
// (v selects which leading axis carries the derivative: 0 = none, 1 = axis 0, 2 = axis 1) in shared
// memory, then every thread finishes its point with p_last+1 terms.  This is the sum-factorised
// form of `apply_tprod` (pyiga/tensor.py:97-128) for one row and makes K2 a streaming kernel.
template <int DIM> struct PbRowVariants { static constexpr int NV = DIM; };

// Y entry for (i_last, c): the NV = DIM variants
template <int DIM>
PB_HD void pb_geo_row_partial(const PbGeoDev& geo, const int* g, int i_last, int c, double* y) {
    const int nc = geo.nc;
    const int f0 = geo.gfirst[0][g[0]];
    const double* T0 = geo.GV[0] + (long long)g[0] * 2 * (geo.pg[0] + 1);
    if constexpr (DIM == 2) {
        double s0 = 0.0, s1 = 0.0;
        for (int a0 = 0; a0 <= geo.pg[0]; ++a0) {
            const double cv = geo.coeffs[((long long)(f0 + a0) * geo.Ng[1] + i_last) * nc + c];
            s0 = fma(cv, T0[a0], s0);
            s1 = fma(cv, T0[geo.pg[0] + 1 + a0], s1);
        }
        y[0] = s0; y[1] = s1;
    } else {
        const int f1 = geo.gfirst[1][g[1]];
        const double* T1 = geo.GV[1] + (long long)g[1] * 2 * (geo.pg[1] + 1);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int a0 = 0; a0 <= geo.pg[0]; ++a0) {
            const double w0 = T0[a0], d0 = T0[geo.pg[0] + 1 + a0];
            double t0 = 0.0, t1 = 0.0;      // sum over a1 with value / derivative weights
            for (int a1 = 0; a1 <= geo.pg[1]; ++a1) {
                const double cv = geo.coeffs[(((long long)(f0 + a0) * geo.Ng[1] + (f1 + a1)) * geo.Ng[2] + i_last) * nc + c];
                t0 = fma(cv, T1[a1], t0);
                t1 = fma(cv, T1[geo.pg[1] + 1 + a1], t1);
            }
            s0 = fma(w0, t0, s0);       // no derivative on the leading axes
            s1 = fma(d0, t0, s1);       // derivative on axis 0
            s2 = fma(w0, t1, s2);       // derivative on axis 1
        }
        y[0] = s0; y[1] = s1; y[2] = s2;
    }
}

// finish one point of the row from Y ([Ng_last][NC][DIM]).  PGL: degree of the geometry on the last
// axis when known at compile time (unrolled), -1 for a run-time loop.  `row_idx`: linear index of the
// row's first point, `gw_row`: product of the Gauss weights of the leading axes.
template <int DIM, int NC, class Prog, int PGL>
PB_HD void pb_fields_row_point(const PbFieldParams& prm, int gl, const double* Y, long long row_idx, double gw_row) {
    const PbGeoDev& geo = prm.geo;
    constexpr int L = DIM - 1;
    const int pgl = PGL >= 0 ? PGL : geo.pg[L];
    const int fl = geo.gfirst[L][gl];
    const double* TL = geo.GV[L] + (long long)gl * 2 * (pgl + 1);
    double val[NC], dv[NC][DIM];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        val[c] = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) dv[c][k] = 0.0;
    }
    auto term = [&](int a) {
        const double w = TL[a], d = TL[pgl + 1 + a];
        const double* yp = Y + (long long)(fl + a) * NC * DIM;
        double y[NC * DIM];
        if constexpr ((NC * DIM) % 2 == 0) {        // 16-byte aligned entries: vector loads
#pragma unroll
            for (int h = 0; h < NC * DIM / 2; ++h) {
                const double2 v = *reinterpret_cast<const double2*>(yp + 2 * h);
                y[2 * h] = v.x;
                y[2 * h + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int h = 0; h < NC * DIM; ++h) y[h] = yp[h];
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            val[c] = fma(w, y[c * DIM + 0], val[c]);
            dv[c][L] = fma(d, y[c * DIM + 0], dv[c][L]);        // derivative on the last axis
            dv[c][0] = fma(w, y[c * DIM + 1], dv[c][0]);        // derivative on axis 0
            if constexpr (DIM == 3) dv[c][1] = fma(w, y[c * DIM + 2], dv[c][1]);
        }
    };
    if constexpr (PGL >= 0) {
#pragma unroll
        for (int a = 0; a <= PGL; ++a) term(a);
    } else {
        for (int a = 0; a <= pgl; ++a) term(a);
    }
    PbPoint pt;
    const long long idx = row_idx + gl;
    pt.idx = idx;
    pt.gw = gw_row * prm.gw[L][gl];
    constexpr int GD = DIM;     // square geometry maps: dim == sdim
    constexpr bool RAT = (NC == GD + 1);
    if constexpr (RAT) {
        // quotient rule without the division: J = (V' W - V W') / W^2, the program folds 1 / W^2 in
        const double W = val[GD];
        pt.jden = W * W;
#pragma unroll
        for (int i = 0; i < GD; ++i) {
            if constexpr (Prog::NEED_X) pt.x[i] = val[i] / W;
#pragma unroll
            for (int k = 0; k < DIM; ++k) pt.J[i][DIM - 1 - k] = dv[i][k] * W - val[i] * dv[GD][k];
        }
    } else {
        pt.jden = 1.0;
#pragma unroll
        for (int i = 0; i < GD; ++i) {
            pt.x[i] = val[i];
#pragma unroll
            for (int k = 0; k < DIM; ++k) pt.J[i][DIM - 1 - k] = dv[i][k];
        }
    }
    double f[Prog::NF];
    Prog::template run<RAT>(prm, pt, f);
    double* o = prm.fields + idx;
    const int nf = prm.nf;
#pragma unroll
    for (int c = 0; c < Prog::NF; ++c) {
        if (c < nf) *o = f[c];
        o += prm.npts;
    }
}

// linear index of the first point of a row and the Gauss weight of its leading axes
template <int DIM>
PB_HD void pb_row_origin(const PbFieldParams& prm, long long row, int* g, long long& row_idx, double& gw_row) {
    g[0] = g[1] = g[2] = 0;
    if constexpr (DIM == 2) {
        g[0] = (int)row;
        gw_row = prm.gw[0][g[0]];
    } else {
        g[0] = (int)(row / prm.G[1]);
        g[1] = (int)(row % prm.G[1]);
        gw_row = prm.gw[0][g[0]] * prm.gw[1][g[1]];
    }
    row_idx = row * prm.G[DIM - 1];
}

// one whole row, sequentially (host emulation) — `Y` is scratch of Ng_last*NC*DIM doubles
template <int DIM, int NC, class Prog>
PB_HD void pb_fields_row_seq(const PbFieldParams& prm, long long row, double* Y) {
    int g[3];
    long long row_idx;
    double gw_row;
    pb_row_origin<DIM>(prm, row, g, row_idx, gw_row);
    const int NgL = prm.geo.Ng[DIM - 1];
    for (int i = 0; i < NgL; ++i)
        for (int c = 0; c < NC; ++c) pb_geo_row_partial<DIM>(prm.geo, g, i, c, Y + ((long long)i * NC + c) * DIM);
    for (int gl = 0; gl < prm.G[DIM - 1]; ++gl) pb_fields_row_point<DIM, NC, Prog, -1>(prm, gl, Y, row_idx, gw_row);
}

#if defined(__CUDACC__)
// One block per group of PB_K2_ROWS consecutive grid rows (same g0 in 3D); dynamic shared memory:
// PB_K2_ROWS * Ng_last*NC*DIM doubles.  Threads own fixed positions on the last axis and visit the
// rows of the group one after the other, so the block synchronises once per group.
#define PB_K2_ROWS 8
template <int DIM, int NC, class Prog>
__global__ void __launch_bounds__(128) pb_fields_row_kernel(const __grid_constant__ PbFieldParams prm, long long row_begin,
                                                            long long row_end) {
    extern __shared__ __align__(16) double pb_Y[];
    const long long row0 = row_begin + (long long)blockIdx.x * PB_K2_ROWS;
    const int nrows = (int)((row_end - row0) < PB_K2_ROWS ? (row_end - row0) : PB_K2_ROWS);
    const int NgL = prm.geo.Ng[DIM - 1];
    const int ysz = NgL * NC * DIM;
    for (int e = threadIdx.x; e < nrows * NgL * NC; e += blockDim.x) {
        const int r = e / (NgL * NC), t = e % (NgL * NC);
        const long long row = row0 + r;
        int g[3] = {0, 0, 0};
        if constexpr (DIM == 2) g[0] = (int)row;
        else { g[0] = (int)(row / prm.G[1]); g[1] = (int)(row % prm.G[1]); }
        pb_geo_row_partial<DIM>(prm.geo, g, t / NC, t % NC, pb_Y + (long long)r * ysz + (long long)t * DIM);
    }
    __syncthreads();
    const int GL = prm.G[DIM - 1];
    auto rows = [&](auto PGLc) {
        constexpr int PGL = decltype(PGLc)::value;
        for (int r = 0; r < nrows; ++r) {
            int g[3];
            long long row_idx;
            double gw_row;
            pb_row_origin<DIM>(prm, row0 + r, g, row_idx, gw_row);
            const double* Y = pb_Y + (long long)r * ysz;
            for (int gl = threadIdx.x; gl < GL; gl += blockDim.x)
                pb_fields_row_point<DIM, NC, Prog, PGL>(prm, gl, Y, row_idx, gw_row);
        }
    };
    // the loop over the geometry's basis functions of the last axis is unrolled for the usual degrees
    switch (prm.geo.pg[DIM - 1]) {
        case 1: rows(PbInt<1>{}); break;
        case 2: rows(PbInt<2>{}); break;
        case 3: rows(PbInt<3>{}); break;
        default: rows(PbInt<-1>{}); break;
    }
}
#endif

#if defined(__CUDACC__)
template <int DIM, class Prog>
__global__ void __launch_bounds__(256) pb_fields_kernel(const __grid_constant__ PbFieldParams prm) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = prm.pt_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < prm.pt_end; idx += stride)
        pb_fields_point<DIM, Prog>(prm, idx);
}
#endif
