/* CPU restatement (plain C + OpenMP) of the reference's per-entry quadrature.
 *
 * TEST / BASELINE INFRASTRUCTURE — never linked into or loaded by the pyiga_b200 package.
 *
 * Follows pyiga/genericasm.pxi:691-700 (multi_entries_chunk), pyiga/assemblers.pyx:1499-1540
 * (entry_impl: intersect the Gauss-index supports per axis, slice tables and fields) and
 * :1455-1494 (combine: triple loop over the joint support; mass :1255-1322; 2D :281-349, :116-172).
 * Inputs are the reference's own data layout: dense basis tables C_k[N_k][G_k][2] (value,
 * derivative), supports ms_k[N_k][2] in Gauss-node indices, fields[G0][G1][G2][nf].
 * Threads split the entry list into contiguous chunks like pyiga's thread pool.
 */
#include <stddef.h>
#include <omp.h>

static double entry3(int form, const size_t* N, const size_t* G, const long* const* ms, const double* const* C,
                     const double* fields, int nf, size_t I, size_t J) {
    size_t i[3], j[3];
    long a[3], b[3];
    for (int k = 2; k >= 0; --k) { i[k] = I % N[k]; I /= N[k]; j[k] = J % N[k]; J /= N[k]; }
    for (int k = 0; k < 3; ++k) {
        long lo = ms[k][2 * j[k]] > ms[k][2 * i[k]] ? ms[k][2 * j[k]] : ms[k][2 * i[k]];
        long hi = ms[k][2 * j[k] + 1] < ms[k][2 * i[k] + 1] ? ms[k][2 * j[k] + 1] : ms[k][2 * i[k] + 1];
        if (lo >= hi) return 0.0;
        a[k] = lo; b[k] = hi;
    }
    const double* u0 = C[0] + (j[0] * G[0]) * 2, *v0 = C[0] + (i[0] * G[0]) * 2;
    const double* u1 = C[1] + (j[1] * G[1]) * 2, *v1 = C[1] + (i[1] * G[1]) * 2;
    const double* u2 = C[2] + (j[2] * G[2]) * 2, *v2 = C[2] + (i[2] * G[2]) * 2;
    double r = 0.0;
    for (long g0 = a[0]; g0 < b[0]; ++g0)
        for (long g1 = a[1]; g1 < b[1]; ++g1)
            for (long g2 = a[2]; g2 < b[2]; ++g2) {
                const double* f = fields + (((size_t)g0 * G[1] + g1) * G[2] + g2) * nf;
                if (form == 1) {
                    r += f[0] * (u0[2 * g0] * u1[2 * g1] * u2[2 * g2]) * (v0[2 * g0] * v1[2 * g1] * v2[2 * g2]);
                } else {
                    const double dux = u0[2 * g0] * u1[2 * g1] * u2[2 * g2 + 1];
                    const double duy = u0[2 * g0] * u1[2 * g1 + 1] * u2[2 * g2];
                    const double duz = u0[2 * g0 + 1] * u1[2 * g1] * u2[2 * g2];
                    const double dvx = v0[2 * g0] * v1[2 * g1] * v2[2 * g2 + 1];
                    const double dvy = v0[2 * g0] * v1[2 * g1 + 1] * v2[2 * g2];
                    const double dvz = v0[2 * g0 + 1] * v1[2 * g1] * v2[2 * g2];
                    r += (f[0] * dux + f[1] * duy + f[2] * duz) * dvx + (f[1] * dux + f[3] * duy + f[4] * duz) * dvy
                         + (f[2] * dux + f[4] * duy + f[5] * duz) * dvz;
                }
            }
    return r;
}

static double entry2(int form, const size_t* N, const size_t* G, const long* const* ms, const double* const* C,
                     const double* fields, int nf, size_t I, size_t J) {
    size_t i[2], j[2];
    long a[2], b[2];
    for (int k = 1; k >= 0; --k) { i[k] = I % N[k]; I /= N[k]; j[k] = J % N[k]; J /= N[k]; }
    for (int k = 0; k < 2; ++k) {
        long lo = ms[k][2 * j[k]] > ms[k][2 * i[k]] ? ms[k][2 * j[k]] : ms[k][2 * i[k]];
        long hi = ms[k][2 * j[k] + 1] < ms[k][2 * i[k] + 1] ? ms[k][2 * j[k] + 1] : ms[k][2 * i[k] + 1];
        if (lo >= hi) return 0.0;
        a[k] = lo; b[k] = hi;
    }
    const double* u0 = C[0] + (j[0] * G[0]) * 2, *v0 = C[0] + (i[0] * G[0]) * 2;
    const double* u1 = C[1] + (j[1] * G[1]) * 2, *v1 = C[1] + (i[1] * G[1]) * 2;
    double r = 0.0;
    for (long g0 = a[0]; g0 < b[0]; ++g0)
        for (long g1 = a[1]; g1 < b[1]; ++g1) {
            const double* f = fields + ((size_t)g0 * G[1] + g1) * nf;
            if (form == 1) {
                r += f[0] * (u0[2 * g0] * u1[2 * g1]) * (v0[2 * g0] * v1[2 * g1]);
            } else {
                const double dux = u0[2 * g0] * u1[2 * g1 + 1], duy = u0[2 * g0 + 1] * u1[2 * g1];
                const double dvx = v0[2 * g0] * v1[2 * g1 + 1], dvy = v0[2 * g0 + 1] * v1[2 * g1];
                r += (f[0] * dux + f[1] * duy) * dvx + (f[1] * dux + f[2] * duy) * dvy;
            }
        }
    return r;
}

/* form: 1 = mass, 2 = stiffness.  ij: n x 2 (row, column). */
void oracle_multi_entries(int dim, int form, const size_t* N, const size_t* G, const long* const* ms,
                          const double* const* C, const double* fields, int nf, const size_t* ij, size_t n,
                          double* out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long e = 0; e < (long)n; ++e)
        out[e] = dim == 3 ? entry3(form, N, G, ms, C, fields, nf, ij[2 * e], ij[2 * e + 1])
                          : entry2(form, N, G, ms, C, fields, nf, ij[2 * e], ij[2 * e + 1]);
}
