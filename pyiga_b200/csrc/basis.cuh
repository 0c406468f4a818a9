// K1 — B-spline basis functions and derivatives at a list of nodes.
//
// Replaces the reference's per-node Cython loop `active_deriv` -> `bspline_active_deriv_single`
// (pyiga/bspline_cy.pyx:42-145, Piegl & Tiller "The NURBS Book" algorithm A2.3) and the span
// search `pyx_findspan` (pyiga/bspline_cy.pyx:13-27).  One thread per node; the triangular
// table lives in local registers.  Output is the compact layout the other kernels consume:
//     first[g]              index of the first active function at node g   (span - p)
//     values[g][r][a]       r-th derivative (r = 0..nd-1) of function first[g]+a,  a = 0..p
#pragma once
#include "common.cuh"

// knot span index i with kv[i] <= u < kv[i+1]; the right end point belongs to the last span
PB_HD int pb_findspan(const double* kv, int nk, int p, double u) {
    if (u >= kv[nk - p - 1]) return nk - p - 2;
    int a = 0, b = nk - 1;
    while (b - a > 1) {
        const int c = a + (b - a) / 2;
        if (kv[c] > u) b = c; else a = c;
    }
    return a;
}

// all derivatives up to order nd-1 of the p+1 functions active at u;  out[r*(p+1) + a]
PB_HD void pb_basis_derivs(const double* kv, int p, int span, double u, int nd, double* out) {
    double ndu[PB_MAXP + 1][PB_MAXP + 1];
    double left[PB_MAXP + 1], right[PB_MAXP + 1];
    ndu[0][0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        left[j - 1] = u - kv[span + 1 - j];
        right[j - 1] = kv[span + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r] + left[j - r - 1];            // knot differences (lower triangle)
            const double t = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r] * t;                  // basis values of degree j (upper triangle)
            saved = left[j - r - 1] * t;
        }
        ndu[j][j] = saved;
    }
    for (int a = 0; a <= p; ++a) out[a] = ndu[a][p];
    if (nd <= 1) return;

    double rowa[PB_MAXP + 1], rowb[PB_MAXP + 1];
    for (int r = 0; r <= p; ++r) {
        double* a1 = rowa;
        double* a2 = rowb;
        a1[0] = 1.0;
        double fac = (double)p;
        for (int k = 1; k < nd; ++k) {
            double d = 0.0;
            if (k <= p) {
                const int rk = r - k, pk = p - k;
                if (r >= k) {
                    a2[0] = a1[0] / ndu[pk + 1][rk];
                    d = a2[0] * ndu[rk][pk];
                }
                const int j1 = (rk >= -1) ? 1 : -rk;
                const int j2 = (r - 1 <= pk) ? k - 1 : p - r;
                for (int j = j1; j <= j2; ++j) {
                    a2[j] = (a1[j] - a1[j - 1]) / ndu[pk + 1][rk + j];
                    d += a2[j] * ndu[rk + j][pk];
                }
                if (r <= pk) {
                    a2[k] = -a1[k - 1] / ndu[pk + 1][r];
                    d += a2[k] * ndu[r][pk];
                }
                out[k * (p + 1) + r] = d * fac;
                fac *= (double)pk;
                double* t = a1; a1 = a2; a2 = t;
            } else {
                out[k * (p + 1) + r] = 0.0;     // derivative order above the degree
            }
        }
    }
}

PB_HD void pb_basis_node(const double* kv, int nk, int p, const double* nodes, int nd, int* first,
                         double* values, int g) {
    const double u = nodes[g];
    const int span = pb_findspan(kv, nk, p, u);
    double buf[3 * (PB_MAXP + 1)];
    pb_basis_derivs(kv, p, span, u, nd, buf);
    if (first) first[g] = span - p;
    for (int k = 0; k < nd * (p + 1); ++k) values[(long long)g * nd * (p + 1) + k] = buf[k];
}

// several independent K1 jobs (the axes of a space, the axes of a geometry) in one launch
#define PB_BASIS_MAXJOBS 6
struct PbBasisBatch {
    int njobs;
    const double* kv[PB_BASIS_MAXJOBS];
    int nk[PB_BASIS_MAXJOBS], p[PB_BASIS_MAXJOBS];
    const double* nodes[PB_BASIS_MAXJOBS];
    int m[PB_BASIS_MAXJOBS], nd[PB_BASIS_MAXJOBS];
    int* first[PB_BASIS_MAXJOBS];
    double* values[PB_BASIS_MAXJOBS];
};

#if defined(__CUDACC__)
__global__ void pb_basis_batch_kernel(const __grid_constant__ PbBasisBatch b) {
    const int j = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < b.m[j]) pb_basis_node(b.kv[j], b.nk[j], b.p[j], b.nodes[j], b.nd[j], b.first[j], b.values[j], g);
}

__global__ void pb_basis_kernel(const double* __restrict__ kv, int nk, int p, const double* __restrict__ nodes,
                                int m, int nd, int* __restrict__ first, double* __restrict__ values) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < m) pb_basis_node(kv, nk, p, nodes, nd, first, values, g);
}
#endif
