// DFMA throughput of one B200 as a function of resident warps per SM and independent chains per thread.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_ilp fp64_ilp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(int iters, double a, double* sink) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, 1e-30);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) *sink = s;
}

template <int ILP>
void run(int warps_per_sm, int sms, double* sink) {
    const int threads = 128, blocks_per_sm = warps_per_sm / 4;
    const int iters = 200000 / ILP * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<sms * blocks_per_sm, threads>>>(iters / 10, 1.0000001, sink);
    cudaEventRecord(e0);
    k<ILP><<<sms * blocks_per_sm, threads>>>(iters, 1.0000001, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * ILP * (double)iters * threads * sms * blocks_per_sm;
    // cycles per DFMA per warp: time * clock / (iters * ILP)
    printf("warps/SM %2d  ILP %2d  %7.2f TFLOP/s  (%.2f ns per dependent step)\n", warps_per_sm, ILP, fl / (ms * 1e-3) / 1e12,
           ms * 1e6 / iters);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* sink;
    cudaMalloc(&sink, 8);
    for (int w : {4, 8, 12, 16, 32, 64}) {
        run<1>(w, p.multiProcessorCount, sink);
        run<2>(w, p.multiProcessorCount, sink);
        run<4>(w, p.multiProcessorCount, sink);
        run<8>(w, p.multiProcessorCount, sink);
        run<16>(w, p.multiProcessorCount, sink);
    }
    return 0;
}
