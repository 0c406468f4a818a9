// Do FP64 tensor-core MMAs (mma.sync m8n8k4 f64) and vector DFMAs share a pipe on B200?
// Runs DMMA alone, DFMA alone and both interleaved; prints TFLOP/s of each.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void k(int iters, double a, double* sink) {
    double c[8][2];
    double x[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; x[i] = threadIdx.x * 1e-9 + i; }
    const double bb = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c[i], a, bb);
#pragma unroll
        for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, 1e-30);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + x[i];
    if (s == 123.456) *sink = s;
}

template <int NM, int NF>
void run(const char* name, int sms, double* sink) {
    const int threads = 256, blocks = sms * 4, iters = 100000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF><<<blocks, threads>>>(iters / 10, 1.0000001, sink);
    cudaEventRecord(e0);
    k<NM, NF><<<blocks, threads>>>(iters, 1.0000001, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double fl_mma = 2.0 * 8 * 8 * 4 * NM * (double)iters * warps;
    const double fl_fma = 2.0 * 32 * NF * (double)iters * warps;
    printf("%-28s %7.3f ms   DMMA %6.2f TFLOP/s   DFMA %6.2f TFLOP/s   sum %6.2f\n", name, ms, fl_mma / (ms * 1e-3) / 1e12,
           fl_fma / (ms * 1e-3) / 1e12, (fl_mma + fl_fma) / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* sink;
    cudaMalloc(&sink, 8);
    run<8, 0>("DMMA only (8 chains)", p.multiProcessorCount, sink);
    run<0, 8>("DFMA only (8 chains)", p.multiProcessorCount, sink);
    run<8, 8>("8 DMMA + 8 DFMA per iter", p.multiProcessorCount, sink);
    run<4, 8>("4 DMMA + 8 DFMA per iter", p.multiProcessorCount, sink);
    run<8, 4>("8 DMMA + 4 DFMA per iter", p.multiProcessorCount, sink);
    return 0;
}
