// CSR utilities on the device: row/column restriction (elimination of constrained dofs), matvec,
// gather / scatter of vectors.
//
// The reference eliminates Dirichlet dofs by multiplying with 0/1 selection matrices built from an
// identity matrix (RestrictedLinearSystem, pyiga/assemble.py:575-652):
//     A_r = R_free_v A R_free^T,     b_r = R_free_v (b - A R_elim^T values).
// Here the same result is produced by one counting pass over the column indices of the kept rows, a
// prefix sum, and one compaction pass (order inside a row is preserved, so sorted rows stay sorted),
// and the right-hand side by one CSR matvec.  All of it is HBM-bound integer/byte work:
// 4 B/nnz (count) + 12 B/nnz read + 12 B/nnz written per kept entry (fill).
#pragma once
#include "common.cuh"

// ---- per-row bodies (also run sequentially by the host emulation) --------------------------------
template <class IT>
PB_HD long long pb_csr_count_row(const IT* indptr, const IT* indices, const int* colmap, long long row) {
    long long c = 0;
    for (IT e = indptr[row]; e < indptr[row + 1]; ++e) c += colmap[indices[e]] >= 0 ? 1 : 0;
    return c;
}

template <class IT>
PB_HD void pb_csr_fill_row(const IT* indptr, const IT* indices, const double* values, const int* colmap, long long row,
                           IT dst, IT* indices_new, double* values_new) {
    for (IT e = indptr[row]; e < indptr[row + 1]; ++e) {
        const int c = colmap[indices[e]];
        if (c >= 0) {
            indices_new[dst] = (IT)c;
            values_new[dst] = values[e];
            ++dst;
        }
    }
}

template <class IT>
PB_HD double pb_csr_row_dot(const IT* indptr, const IT* indices, const double* values, const double* x, long long row) {
    double s = 0.0;
    for (IT e = indptr[row]; e < indptr[row + 1]; ++e) s = fma(values[e], x[indices[e]], s);
    return s;
}

#if defined(__CUDACC__)
// one warp per kept row; counts[r] = surviving entries of old row rows[r]
template <class IT>
__global__ void __launch_bounds__(256) pb_csr_restrict_count_kernel(long long nnew, const int* __restrict__ rows,
                                                                    const IT* __restrict__ indptr,
                                                                    const IT* __restrict__ indices,
                                                                    const int* __restrict__ colmap, IT* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nnew; r += nwarps) {
        const long long row = rows[r];
        const IT e0 = indptr[row], e1 = indptr[row + 1];
        int c = 0;
        for (IT e = e0 + lane; e < e1; e += 32) c += colmap[indices[e]] >= 0 ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) counts[r] = (IT)c;
    }
}

// one warp per kept row; entries are compacted 32 at a time with a ballot, which keeps their order
template <class IT>
__global__ void __launch_bounds__(256) pb_csr_restrict_fill_kernel(long long nnew, const int* __restrict__ rows,
                                                                   const IT* __restrict__ indptr,
                                                                   const IT* __restrict__ indices,
                                                                   const double* __restrict__ values,
                                                                   const int* __restrict__ colmap,
                                                                   const IT* __restrict__ indptr_new,
                                                                   IT* __restrict__ indices_new,
                                                                   double* __restrict__ values_new) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nnew; r += nwarps) {
        const long long row = rows[r];
        const IT e0 = indptr[row], e1 = indptr[row + 1];
        IT dst = indptr_new[r];
        for (IT b = e0; b < e1; b += 32) {
            const IT e = b + lane;
            int c = -1;
            double v = 0.0;
            if (e < e1) {
                c = colmap[indices[e]];
                v = values[e];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, c >= 0);
            if (c >= 0) {
                const IT pos = dst + (IT)__popc(bal & ((1u << lane) - 1u));
                indices_new[pos] = (IT)c;
                values_new[pos] = v;
            }
            dst += (IT)__popc(bal);
        }
    }
}

// ---- in-place inclusive prefix sum over a[0..n) in three launches (2048 items per block) -----------
#define PB_SCAN_ITEMS 8
#define PB_SCAN_BLOCK (256 * PB_SCAN_ITEMS)
template <class IT>
__device__ __forceinline__ IT pb_block_scan_256(IT v, IT* total) {   // inclusive scan of one value per thread
    __shared__ long long wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const IT t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();                        // wsum may still be read from a previous call
    if (lane == 31) wsum[w] = (long long)v;
    __syncthreads();
    IT off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < w) off += (IT)wsum[k];
        tot += (IT)wsum[k];
    }
    *total = tot;
    return v + off;
}

template <class IT>
__global__ void __launch_bounds__(256) pb_scan_sums_kernel(const IT* __restrict__ a, long long n, IT* __restrict__ sums) {
    const long long base = (long long)blockIdx.x * PB_SCAN_BLOCK + (long long)threadIdx.x * PB_SCAN_ITEMS;
    IT s = 0;
#pragma unroll
    for (int k = 0; k < PB_SCAN_ITEMS; ++k)
        if (base + k < n) s += a[base + k];
    IT total;
    pb_block_scan_256<IT>(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// exclusive scan of the block sums by a single block
template <class IT>
__global__ void __launch_bounds__(256) pb_scan_top_kernel(IT* __restrict__ sums, long long nb) {
    IT carry = 0;
    for (long long b0 = 0; b0 < nb; b0 += 256) {
        const long long i = b0 + threadIdx.x;
        const IT v = i < nb ? sums[i] : (IT)0;
        IT total;
        const IT inc = pb_block_scan_256<IT>(v, &total);
        if (i < nb) sums[i] = carry + inc - v;
        carry += total;
    }
}

template <class IT>
__global__ void __launch_bounds__(256) pb_scan_apply_kernel(IT* __restrict__ a, long long n, const IT* __restrict__ sums) {
    const long long base = (long long)blockIdx.x * PB_SCAN_BLOCK + (long long)threadIdx.x * PB_SCAN_ITEMS;
    IT v[PB_SCAN_ITEMS];
    IT s = 0;
#pragma unroll
    for (int k = 0; k < PB_SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? a[base + k] : (IT)0;
        s += v[k];
    }
    IT total;
    const IT inc = pb_block_scan_256<IT>(s, &total);
    IT run = sums[blockIdx.x] + inc - s;
#pragma unroll
    for (int k = 0; k < PB_SCAN_ITEMS; ++k) {
        run += v[k];
        if (base + k < n) a[base + k] = run;
    }
}

// y_out[r] = (y_in ? y_in[r] : 0) + alpha * (A x)[r],  one warp per row
template <class IT>
__global__ void __launch_bounds__(256) pb_csr_matvec_kernel(long long nrows, const IT* __restrict__ indptr,
                                                            const IT* __restrict__ indices,
                                                            const double* __restrict__ values,
                                                            const double* __restrict__ x, const double* __restrict__ y_in,
                                                            double alpha, double* __restrict__ y_out) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const IT e0 = indptr[r], e1 = indptr[r + 1];
        double s = 0.0;
        for (IT e = e0 + lane; e < e1; e += 32) s = fma(values[e], x[indices[e]], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y_out[r] = (y_in ? y_in[r] : 0.0) + alpha * s;
    }
}

__global__ void __launch_bounds__(256) pb_gather_kernel(long long n, const int* __restrict__ idx, const double* __restrict__ in,
                                                        double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = in[idx[k]];
}
__global__ void __launch_bounds__(256) pb_scatter_kernel(long long n, const int* __restrict__ idx, const double* __restrict__ in,
                                                         double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[idx[k]] = in[k];
}
#endif
