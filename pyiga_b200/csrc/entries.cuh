// Per-entry quadrature — the `multi_entries` / `entry` operator of the assembler protocol.
//
// Restates the reference's entry-wise path (pyiga/genericasm.pxi:691-700 `multi_entries_chunk`,
// pyiga/assemblers.pyx:1499-1540 `entry_impl`, :1455-1494 `combine`): unravel (I,J) into per-axis
// indices, intersect the supports of test function i and trial function j on every axis (empty ->
// the entry stays 0), then sum the integrand over the Gauss points of the joint support.
// One thread per requested entry.  It serves scattered index lists (hierarchical / low-rank
// callers of the protocol) and is the general fall-back for configurations the sum-factorised
// pipeline has no instantiation for; full-matrix assembly goes through walk.cuh instead.
//
// The integrand is a list of terms  C_t(g) * d^{bt} v_i(g) * d^{bu} u_j(g)  where slot 0 is the
// function value and slot 1+k the derivative along tensor axis k; slots >= PB_SLOT_EXT carry one
// derivative order per axis (common.cuh) and are evaluated from the three-row tables V3u / V3v.
#pragma once
#include "common.cuh"

struct PbTerm { int field; int bt; int bu; };

struct PbEntryParams {
    int dim;
    PbAxis ax[PB_MAXDIM];
    const double* fields;       // [nf][G0][G1][G2]
    long long npts;
    int nterms;
    PbTerm terms[PB_MAXTERMS];
    int ext;                    // some term has a second or mixed derivative slot: orders from ot / ou
    unsigned char ot[PB_MAXTERMS][PB_MAXDIM], ou[PB_MAXTERMS][PB_MAXDIM];   // derivative order per axis (test / trial)
    const unsigned long long* ij;   // [n][2]
    long long n;
    double* out;                // [n]
};

// value, first and (when the axis carries the three-row table) second derivative of active function a at node g
PB_HD void pb_entry_rows(const double* V2, const double* V3, int g, int p, int a, double* r) {
    if (V3) {
        const double* T = V3 + (long long)g * 3 * (p + 1);
        r[0] = T[a]; r[1] = T[p + 1 + a]; r[2] = T[2 * (p + 1) + a];
    } else {
        const double* T = V2 + (long long)g * 2 * (p + 1);
        r[0] = T[a]; r[1] = T[p + 1 + a]; r[2] = 0.0;
    }
}

template <int DIM>
PB_HD double pb_entry(const PbEntryParams& prm, unsigned long long I, unsigned long long J) {
    int i[3] = {0, 0, 0}, j[3] = {0, 0, 0};
    for (int k = DIM - 1; k >= 0; --k) {
        i[k] = (int)(I % (unsigned long long)prm.ax[k].Nv); I /= (unsigned long long)prm.ax[k].Nv;
        j[k] = (int)(J % (unsigned long long)prm.ax[k].Nu); J /= (unsigned long long)prm.ax[k].Nu;
    }
    if (I != 0 || J != 0) return 0.0;                         // index outside the matrix
    int sa[3] = {0, 0, 0}, sb[3] = {1, 1, 1};
    for (int k = 0; k < DIM; ++k) {
        const PbAxis& A = prm.ax[k];
        sa[k] = pb_max(A.supp_v[2 * i[k]], A.supp_u[2 * j[k]]);
        sb[k] = pb_min(A.supp_v[2 * i[k] + 1], A.supp_u[2 * j[k] + 1]);
        if (sa[k] >= sb[k]) return 0.0;                      // no joint support
    }
    const PbAxis& A0 = prm.ax[0];
    const PbAxis& A1 = prm.ax[1];
    const PbAxis& A2 = prm.ax[DIM - 1];
    double r = 0.0;
    for (int s0 = sa[0]; s0 < sb[0]; ++s0) {
        const int av0 = i[0] - A0.first_v[s0], au0 = j[0] - A0.first_u[s0];
        for (int q0 = 0; q0 < A0.q; ++q0) {
            const int g0 = s0 * A0.q + q0;
            double v0[3], u0[3];
            pb_entry_rows(A0.Vv, A0.V3v, g0, A0.pv, av0, v0);
            pb_entry_rows(A0.Vu, A0.V3u, g0, A0.pu, au0, u0);
            for (int s1 = sa[1]; s1 < sb[1]; ++s1) {
                const int av1 = i[1] - A1.first_v[s1], au1 = j[1] - A1.first_u[s1];
                for (int q1 = 0; q1 < A1.q; ++q1) {
                    const int g1 = s1 * A1.q + q1;
                    double v1[3], u1[3];
                    pb_entry_rows(A1.Vv, A1.V3v, g1, A1.pv, av1, v1);
                    pb_entry_rows(A1.Vu, A1.V3u, g1, A1.pu, au1, u1);
                    if constexpr (DIM == 2) {
                        if (prm.ext) {
                            const long long pt = (long long)g0 * A1.G + g1;
                            for (int t = 0; t < prm.nterms; ++t)
                                r += prm.fields[(long long)prm.terms[t].field * prm.npts + pt]
                                     * (v0[prm.ot[t][0]] * v1[prm.ot[t][1]]) * (u0[prm.ou[t][0]] * u1[prm.ou[t][1]]);
                            continue;
                        }
                        const double vt[3] = {v0[0] * v1[0], v0[1] * v1[0], v0[0] * v1[1]};
                        const double ut[3] = {u0[0] * u1[0], u0[1] * u1[0], u0[0] * u1[1]};
                        const long long pt = (long long)g0 * A1.G + g1;
                        for (int t = 0; t < prm.nterms; ++t)
                            r += prm.fields[(long long)prm.terms[t].field * prm.npts + pt]
                                 * vt[prm.terms[t].bt] * ut[prm.terms[t].bu];
                    } else {
                        for (int s2 = sa[2]; s2 < sb[2]; ++s2) {
                            const int av2 = i[2] - A2.first_v[s2], au2 = j[2] - A2.first_u[s2];
                            for (int q2 = 0; q2 < A2.q; ++q2) {
                                const int g2 = s2 * A2.q + q2;
                                double v2[3], u2[3];
                                pb_entry_rows(A2.Vv, A2.V3v, g2, A2.pv, av2, v2);
                                pb_entry_rows(A2.Vu, A2.V3u, g2, A2.pu, au2, u2);
                                if (prm.ext) {
                                    const long long pt = ((long long)g0 * A1.G + g1) * A2.G + g2;
                                    for (int t = 0; t < prm.nterms; ++t)
                                        r += prm.fields[(long long)prm.terms[t].field * prm.npts + pt]
                                             * (v0[prm.ot[t][0]] * v1[prm.ot[t][1]] * v2[prm.ot[t][2]])
                                             * (u0[prm.ou[t][0]] * u1[prm.ou[t][1]] * u2[prm.ou[t][2]]);
                                    continue;
                                }
                                const double vt[4] = {v0[0] * v1[0] * v2[0], v0[1] * v1[0] * v2[0],
                                                      v0[0] * v1[1] * v2[0], v0[0] * v1[0] * v2[1]};
                                const double ut[4] = {u0[0] * u1[0] * u2[0], u0[1] * u1[0] * u2[0],
                                                      u0[0] * u1[1] * u2[0], u0[0] * u1[0] * u2[1]};
                                const long long pt = ((long long)g0 * A1.G + g1) * A2.G + g2;
                                for (int t = 0; t < prm.nterms; ++t)
                                    r += prm.fields[(long long)prm.terms[t].field * prm.npts + pt]
                                         * vt[prm.terms[t].bt] * ut[prm.terms[t].bu];
                            }
                        }
                    }
                }
            }
        }
    }
    return r;
}

// entry e of the MLB slab that starts at band index mu0_begin on axis 0
template <int DIM>
PB_HD double pb_entry_mlb(const PbEntryParams& prm, long long mu0_begin, long long e) {
    long long r = e;
    int mu[3] = {0, 0, 0};
    for (int k = DIM - 1; k >= 1; --k) { mu[k] = (int)(r % prm.ax[k].M); r /= prm.ax[k].M; }
    mu[0] = (int)(r + mu0_begin);
    unsigned long long I = 0, J = 0;
    for (int k = 0; k < DIM; ++k) {
        I = I * (unsigned long long)prm.ax[k].Nv + (unsigned long long)prm.ax[k].pair_i[mu[k]];
        J = J * (unsigned long long)prm.ax[k].Nu + (unsigned long long)prm.ax[k].pair_j[mu[k]];
    }
    return pb_entry<DIM>(prm, I, J);
}

// Partial-row assembly: entry e of requested row I.  The row's pattern is the Cartesian product of
// the per-axis column ranges [jmin_k[i_k], jmin_k[i_k] + nb_k), enumerated with the last axis
// fastest, i.e. in increasing column order (what MLStructure.nonzeros_for_rows lists row by row,
// pyiga/mlmatrix.py:150-185); the value is the same per-entry quadrature as multi_entries.
template <int DIM, class IdxT>
PB_HD void pb_row_entry(const PbEntryParams& prm, unsigned long long I, long long base, int e, IdxT* indices, double* values) {
    int i[3] = {0, 0, 0}, nb[3] = {1, 1, 1}, jm[3] = {0, 0, 0};
    unsigned long long r = I;
    for (int k = DIM - 1; k >= 0; --k) {
        i[k] = (int)(r % (unsigned long long)prm.ax[k].Nv);
        r /= (unsigned long long)prm.ax[k].Nv;
        nb[k] = prm.ax[k].row_start[i[k] + 1] - prm.ax[k].row_start[i[k]];
        jm[k] = prm.ax[k].jmin[i[k]];
    }
    int t = e;
    unsigned long long J = 0, mul = 1;
    for (int k = DIM - 1; k >= 0; --k) {
        J += mul * (unsigned long long)(jm[k] + t % nb[k]);
        t /= nb[k];
        mul *= (unsigned long long)prm.ax[k].Nu;
    }
    indices[base + e] = (IdxT)J;
    values[base + e] = pb_entry<DIM>(prm, I, J);
}

#if defined(__CUDACC__)
// one block per requested row
template <int DIM, class IdxT>
__global__ void __launch_bounds__(128) pb_rows_fill_kernel(const __grid_constant__ PbEntryParams prm,
                                                           const long long* __restrict__ rows,
                                                           const IdxT* __restrict__ indptr, long long nrows,
                                                           IdxT* __restrict__ indices, double* __restrict__ values) {
    for (long long r = blockIdx.x; r < nrows; r += gridDim.x) {
        const long long base = (long long)indptr[r];
        const int cnt = (int)((long long)indptr[r + 1] - base);
        for (int e = threadIdx.x; e < cnt; e += blockDim.x)
            pb_row_entry<DIM, IdxT>(prm, (unsigned long long)rows[r], base, e, indices, values);
    }
}

template <int DIM>
__global__ void __launch_bounds__(128) pb_entries_kernel(const __grid_constant__ PbEntryParams prm) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < prm.n; e += stride)
        prm.out[e] = pb_entry<DIM>(prm, prm.ij[2 * e], prm.ij[2 * e + 1]);
}

// all entries of the band pattern, rows of axis 0 in [row_begin,row_end): writes the MLB slab
template <int DIM>
__global__ void __launch_bounds__(128) pb_entries_mlb_kernel(const __grid_constant__ PbEntryParams prm,
                                                             long long mu0_begin, long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
        prm.out[e] = pb_entry_mlb<DIM>(prm, mu0_begin, e);
}
#endif
