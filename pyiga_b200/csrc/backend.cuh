// Thin runtime layer under api.cu.
//
// The product build (nvcc) maps these calls to the CUDA runtime and launches the kernels.
// Defining PB_EMULATE (done only by tests/emu/build_emu.py, with a plain C++ compiler) maps
// "device" memory to host memory and runs every kernel body as a sequential loop over its thread
// index.  The emulation exists so that the index logic of the kernels can be checked in a
// container without a GPU; it is test infrastructure, is never built into libpyiga_b200.so and is
// never loaded by the pyiga_b200 package.
#pragma once
#include <cstdlib>
#include <cstring>

#ifdef PB_EMULATE
typedef void* pbStream;
typedef int pbError;
#define pbSuccess 0
static inline pbError pbSetDevice(int) { return 0; }
static inline pbError pbMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline pbError pbFree(void* p) { std::free(p); return 0; }
static inline pbError pbMemcpyH2D(void* d, const void* s, size_t n, pbStream) { std::memcpy(d, s, n); return 0; }
static inline pbError pbStreamSync(pbStream) { return 0; }
static inline pbError pbLastError() { return 0; }
static inline const char* pbErrorString(pbError e) { return e ? "allocation failed" : "ok"; }
// run body(i) for i in [0, n)
template <class F> static inline void pb_emu_for(long long n, F&& body) { for (long long i = 0; i < n; ++i) body(i); }
#else
#include <cuda_runtime.h>
typedef cudaStream_t pbStream;
typedef cudaError_t pbError;
#define pbSuccess cudaSuccess
static inline pbError pbSetDevice(int d) { return cudaSetDevice(d); }
static inline pbError pbMalloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1); }
static inline pbError pbFree(void* p) { return cudaFree(p); }
static inline pbError pbMemcpyH2D(void* d, const void* s, size_t n, pbStream st) {
    return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st);
}
static inline pbError pbStreamSync(pbStream st) { return cudaStreamSynchronize(st); }
static inline pbError pbLastError() { return cudaGetLastError(); }
static inline const char* pbErrorString(pbError e) { return cudaGetErrorString(e); }
#endif
