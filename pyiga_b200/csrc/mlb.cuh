// Multi-level banded (MLB) matrices on the device: CSR export, matvec, Kronecker matvec.
//
// The MLB tensor `data[mu0][mu1][mu2]` stores A[I,J] for (i_k, j_k) = bidx_k[mu_k]
// (pyiga/mlmatrix.py:201-269).  Because every bidx_k is sorted by i then j, row i_k owns the
// contiguous band range [row_start_k[i_k], row_start_k[i_k+1]) and column j_k = jmin_k[i_k] + offset;
// no index arrays are needed.
//   * CSR export replaces `ml_nonzero_3d` + scipy COO->CSR (+ mirror) of the reference
//     (pyiga/mlmatrix_cy.pyx:257-289, pyiga/assemble.py:743-754): indptr is a closed form of the
//     per-axis prefix sums, indices/data are one permutation pass over the MLB tensor.
//   * matvec replaces the serial `ml_matvec_2d/3d` (pyiga/mlmatrix_cy.pyx:224-325) with one thread
//     per matrix row reading its band rows coalesced.
//   * Kronecker matvec replaces `apply_kronecker` for dense factors (pyiga/kronecker.py:15-34).
#pragma once
#include "common.cuh"

struct PbMlbParams {
    int dim;
    int Nv[PB_MAXDIM], Nu[PB_MAXDIM], M[PB_MAXDIM];
    const int* row_start[PB_MAXDIM];
    const int* jmin[PB_MAXDIM];
    const int* pair_i[PB_MAXDIM];
    const int* pair_j[PB_MAXDIM];
    int row0_begin, row0_end;       // slab of rows on axis 0
    const double* data;             // MLB slab: slot 0 <-> mu0 = row_start[0][row0_begin]
};

// number of stored entries before row (i0,i1,i2), relative to the slab start
PB_HD long long pb_csr_row_offset(const PbMlbParams& p, const int* i) {
    const long long m1 = (p.dim > 1) ? p.M[1] : 1, m2 = (p.dim > 2) ? p.M[2] : 1;
    const long long nb0 = p.row_start[0][i[0] + 1] - p.row_start[0][i[0]];
    long long off = (long long)(p.row_start[0][i[0]] - p.row_start[0][p.row0_begin]) * m1 * m2;
    if (p.dim == 1) return off;
    const long long nb1 = p.row_start[1][i[1] + 1] - p.row_start[1][i[1]];
    if (p.dim == 2) return off + nb0 * p.row_start[1][i[1]];
    return off + nb0 * ((long long)p.row_start[1][i[1]] * m2 + nb1 * p.row_start[2][i[2]]);
}

template <class IdxT>
PB_HD void pb_csr_indptr_row(const PbMlbParams& p, long long nrows, IdxT* indptr, long long r) {
    if (r == nrows) {
        const long long m1 = (p.dim > 1) ? p.M[1] : 1, m2 = (p.dim > 2) ? p.M[2] : 1;
        indptr[r] = (IdxT)((long long)(p.row_start[0][p.row0_end] - p.row_start[0][p.row0_begin]) * m1 * m2);
        return;
    }
    int i[3] = {0, 0, 0};
    long long t = r;
    for (int k = p.dim - 1; k >= 1; --k) { i[k] = (int)(t % p.Nv[k]); t /= p.Nv[k]; }
    i[0] = (int)t + p.row0_begin;
    indptr[r] = (IdxT)pb_csr_row_offset(p, i);
}

// one MLB element: column index + value go to their CSR slot
template <class IdxT>
PB_HD void pb_csr_fill_elem(const PbMlbParams& p, IdxT* indices, double* values, long long e) {
    const int mu0_base = p.row_start[0][p.row0_begin];
    long long t = e;
    int mu[3] = {0, 0, 0};
    for (int k = p.dim - 1; k >= 1; --k) { mu[k] = (int)(t % p.M[k]); t /= p.M[k]; }
    mu[0] = (int)t + mu0_base;
    int i[3] = {0, 0, 0};
    long long J = 0, pos = 0;
    for (int k = 0; k < p.dim; ++k) {
        i[k] = p.pair_i[k][mu[k]];
        const int nb = p.row_start[k][i[k] + 1] - p.row_start[k][i[k]];
        J = J * p.Nu[k] + p.pair_j[k][mu[k]];
        pos = pos * nb + (mu[k] - p.row_start[k][i[k]]);
    }
    pos += pb_csr_row_offset(p, i);
    indices[pos] = (IdxT)J;
    values[pos] = p.data[e];
}

// CSR export, one matrix row: entry e = (k0*nb1 + k1)*nb2 + k2 of row (i0,i1,i2) comes from
// data[mu0 = rs0[i0]+k0][mu1 = rs1[i1]+k1][mu2 = rs2[i2]+k2].  Written so that consecutive e are
// consecutive CSR slots (a warp fills a row with coalesced stores; the reads are runs of nb2).
template <class IdxT>
PB_HD void pb_csr_fill_row_entry(const PbMlbParams& p, const int* i, const int* rs, const int* nb, const int* jm,
                                 long long rowoff, int e, IdxT* indices, double* values) {
    const int mu0_base = p.row_start[0][p.row0_begin];
    if (p.dim == 2) {
        const int k1 = e % nb[1], k0 = e / nb[1];
        const long long src = (long long)(rs[0] + k0 - mu0_base) * p.M[1] + rs[1] + k1;
        if (indices) indices[rowoff + e] = (IdxT)((long long)(jm[0] + k0) * p.Nu[1] + jm[1] + k1);
        values[rowoff + e] = p.data[src];
    } else {
        const int k2 = e % nb[2], t = e / nb[2];
        const int k1 = t % nb[1], k0 = t / nb[1];
        const long long src = ((long long)(rs[0] + k0 - mu0_base) * p.M[1] + rs[1] + k1) * p.M[2] + rs[2] + k2;
        if (indices) indices[rowoff + e] = (IdxT)(((long long)(jm[0] + k0) * p.Nu[1] + jm[1] + k1) * p.Nu[2] + jm[2] + k2);
        values[rowoff + e] = p.data[src];
    }
    (void)i;
}

PB_HD void pb_csr_row_tables(const PbMlbParams& p, long long r, int* i, int* rs, int* nb, int* jm) {
    long long t = r;
    for (int k = p.dim - 1; k >= 1; --k) { i[k] = (int)(t % p.Nv[k]); t /= p.Nv[k]; }
    i[0] = (int)t + p.row0_begin;
    for (int k = 0; k < p.dim; ++k) {
        rs[k] = p.row_start[k][i[k]];
        nb[k] = p.row_start[k][i[k] + 1] - rs[k];
        jm[k] = p.jmin[k][i[k]];
    }
}

// y[I] = sum_J A[I,J] x[J] for one row of the slab.  `x` starts at trial index j0 = x_j0_begin on
// axis 0 (halo layout of the slab-distributed operator); y starts at row0_begin.
PB_HD void pb_mlb_matvec_row(const PbMlbParams& p, const double* __restrict__ x, int x_j0_begin,
                             double* __restrict__ y, long long r) {
    int i[3] = {0, 0, 0};
    long long t = r;
    for (int k = p.dim - 1; k >= 1; --k) { i[k] = (int)(t % p.Nv[k]); t /= p.Nv[k]; }
    i[0] = (int)t + p.row0_begin;
    int rs[3] = {0, 0, 0}, nb[3] = {1, 1, 1}, jm[3] = {0, 0, 0};
    for (int k = 0; k < p.dim; ++k) {
        rs[k] = p.row_start[k][i[k]];
        nb[k] = p.row_start[k][i[k] + 1] - rs[k];
        jm[k] = p.jmin[k][i[k]];
    }
    rs[0] -= p.row_start[0][p.row0_begin];
    jm[0] -= x_j0_begin;
    double acc = 0.0;
    if (p.dim == 2) {
        for (int k0 = 0; k0 < nb[0]; ++k0) {
            const double* d = p.data + (long long)(rs[0] + k0) * p.M[1] + rs[1];
            const double* xx = x + (long long)(jm[0] + k0) * p.Nu[1] + jm[1];
            for (int k1 = 0; k1 < nb[1]; ++k1) acc = fma(d[k1], xx[k1], acc);
        }
    } else {
        for (int k0 = 0; k0 < nb[0]; ++k0)
            for (int k1 = 0; k1 < nb[1]; ++k1) {
                const double* d = p.data + ((long long)(rs[0] + k0) * p.M[1] + rs[1] + k1) * p.M[2] + rs[2];
                const double* xx = x + ((long long)(jm[0] + k0) * p.Nu[1] + jm[1] + k1) * p.Nu[2] + jm[2];
                for (int k2 = 0; k2 < nb[2]; ++k2) acc = fma(d[k2], xx[k2], acc);
            }
    }
    y[r] = acc;
}

// mode-k product with a dense factor: y[a, i, c] = sum_j A[i, j] x[a, j, c]
PB_HD void pb_modek_elem(const double* __restrict__ A, int m, int n, const double* __restrict__ x, long long inner,
                         double* __restrict__ y, long long e) {
    const long long c = e % inner;
    const long long t = e / inner;
    const int i = (int)(t % m);
    const long long a = t / m;
    const double* xr = x + a * n * inner + c;
    const double* Ar = A + (long long)i * n;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc = fma(Ar[j], xr[(long long)j * inner], acc);
    y[e] = acc;
}

#if defined(__CUDACC__)
// indptr[r] for local rows r = 0..nrows (inclusive), nrows = (row0_end-row0_begin)*Nv1*Nv2
template <class IdxT>
__global__ void pb_csr_indptr_kernel(const __grid_constant__ PbMlbParams p, long long nrows, IdxT* indptr) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nrows) pb_csr_indptr_row<IdxT>(p, nrows, indptr, r);
}
template <class IdxT>
__global__ void pb_csr_fill_kernel(const __grid_constant__ PbMlbParams p, long long count, IdxT* indices,
                                   double* values) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride)
        pb_csr_fill_elem<IdxT>(p, indices, values, e);
}
// one warp per matrix row
template <class IdxT>
__global__ void __launch_bounds__(256) pb_csr_fill_rows_kernel(const __grid_constant__ PbMlbParams p, long long nrows,
                                                               IdxT* indices, double* values) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        int i[3] = {0, 0, 0}, rs[3] = {0, 0, 0}, nb[3] = {1, 1, 1}, jm[3] = {0, 0, 0};
        pb_csr_row_tables(p, r, i, rs, nb, jm);
        const long long rowoff = pb_csr_row_offset(p, i);
        const int len = nb[0] * nb[1] * nb[2];
        for (int e = lane; e < len; e += 32) pb_csr_fill_row_entry<IdxT>(p, i, rs, nb, jm, rowoff, e, indices, values);
    }
}
__global__ void __launch_bounds__(256) pb_mlb_matvec_kernel(const __grid_constant__ PbMlbParams p, long long nrows,
                                                            const double* __restrict__ x, int x_j0_begin,
                                                            double* __restrict__ y) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows) pb_mlb_matvec_row(p, x, x_j0_begin, y, r);
}
__global__ void pb_modek_kernel(const double* __restrict__ A, int m, int n, const double* __restrict__ x,
                                long long outer, long long inner, double* __restrict__ y) {
    const long long total = outer * m * inner;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
        pb_modek_elem(A, m, n, x, inner, y, e);
}
#endif
