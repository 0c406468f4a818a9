// Sum-factorised contraction for linear forms (arity 1): load vectors F[i] = int sum_b f_b d^b v_i.
//
// Replaces `BaseAssembler*.assemble_vector` (pyiga/genericasm.pxi:129-145, 762-778), which calls
// `entry_impl(I, NULL)` — a full quadrature over the support — for every basis function.  Same idea
// as walk.cuh with a window of p+1 single functions instead of (p+1)^2 pairs: a thread owns a line
// and walks the node axis; function f is complete once the walk leaves its support.
//     out[f] = sum_g ( D_0[f,g] * x0(g) + D_1[f,g] * x1(g) )        (x1 or x0 may be absent)
// Stages contract axis 0, 1, 2 in turn; vectors are tiny compared with matrices, so this kernel is
// not tuned beyond coalescing the non-final stages.
#pragma once
#include "common.cuh"

struct PbWalk1Params {
    long long nthreads;
    int X;                          // tid -> (u, x), x fastest
    const double* in0;              // term contracted with values      (or null)
    const double* in1;              // term contracted with derivatives (or null)
    long long in_su, in_sx, in_sc;
    double* out;
    long long out_su, out_sx, out_sf;   // out_sf: stride of the function index
    int n, N;                       // spans, functions on the axis
    const int* first;               // [n]
    const double* V2;               // [G][2][P+1]
};

template <int P, int Q>
PB_HD void pb_walk1_line(const PbWalk1Params& prm, long long tid) {
    constexpr int P1 = P + 1;
    const int x = (int)(tid % prm.X);
    const long long u = tid / prm.X;
    const long long off_in = u * prm.in_su + (long long)x * prm.in_sx;
    const long long off_out = u * prm.out_su + (long long)x * prm.out_sx;
    double acc[P1];
#pragma unroll
    for (int a = 0; a < P1; ++a) acc[a] = 0.0;
    int f = prm.first[0];
    for (int s = 0; s < prm.n; ++s) {
        const int fs = prm.first[s];
        while (f < fs) {
            prm.out[off_out + (long long)f * prm.out_sf] = acc[0];
#pragma unroll
            for (int a = 0; a < P; ++a) acc[a] = acc[a + 1];
            acc[P] = 0.0;
            ++f;
        }
#pragma unroll
        for (int gq = 0; gq < Q; ++gq) {
            const long long g = (long long)s * Q + gq;
            const double x0 = prm.in0 ? prm.in0[off_in + g * prm.in_sc] : 0.0;
            const double x1 = prm.in1 ? prm.in1[off_in + g * prm.in_sc] : 0.0;
            const double* Vn = prm.V2 + g * (2 * P1);
#pragma unroll
            for (int a = 0; a < P1; ++a) acc[a] = fma(Vn[a], x0, fma(Vn[P1 + a], x1, acc[a]));
        }
    }
#pragma unroll 1
    for (int k = 0; k < P1; ++k) {
        if (f < prm.N) prm.out[off_out + (long long)f * prm.out_sf] = acc[0];
#pragma unroll
        for (int a = 0; a < P; ++a) acc[a] = acc[a + 1];
        acc[P] = 0.0;
        ++f;
    }
}

#if defined(__CUDACC__)
template <int P, int Q>
__global__ void __launch_bounds__(128) pb_walk1_kernel(const __grid_constant__ PbWalk1Params prm) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < prm.nthreads) pb_walk1_line<P, Q>(prm, tid);
}
#endif

typedef int (*PbWalk1Launch)(const PbWalk1Params* prm, void* stream);
extern "C" __attribute__((visibility("default"))) void pb200_register_walk1(int P, int Q, PbWalk1Launch fn);
PbWalk1Launch pb_find_walk1(int P, int Q);
