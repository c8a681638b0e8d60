// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when a kernel header is compiled as plain C++
// for the emulated CTA of tests/host/cta_emul.h (g++ -I tests/host/shim ...).
#pragma once
#include <cmath>
#include <cstdio>
#include "../cta_emul.h"

#define __CUDACC__ 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))

struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

#define threadIdx (cta::thread_idx())
#define blockIdx (cta::block_idx())
#define blockDim (cta::block_dim())
#define gridDim (cta::grid_dim())
#define AXS_DYNAMIC_SMEM(name) unsigned char* const name = cta::g.dyn_smem

static inline void __syncthreads() { cta::syncthreads(); }
static inline void __syncwarp() { cta::syncwarp(); }
static inline unsigned __ballot_sync(unsigned, bool p) { return cta::ballot(p); }
static inline double __shfl_sync(unsigned, double v, int src) { return cta::shfl(v, src); }
static inline double __shfl_xor_sync(unsigned, double v, int m) { return cta::shfl(v, (int)(cta::g.cur & 31) ^ m); }
static inline float __shfl_xor_sync(unsigned, float v, int m) { return (float)cta::shfl((double)v, (int)(cta::g.cur & 31) ^ m); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int atomicMax(int* p, int v) { const int old = *p; if (v > old) *p = v; return old; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline long long clock64() { static long long c = 0; return c += 7; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline double atomicAdd(double* p, double v) { const double old = *p; *p = old + v; return old; }
static inline float __double2float_rn(double x) { return (float)x; }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)cta::shfl((double)v, (int)(cta::g.cur & 31) ^ m); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)cta::shfl((double)v, src); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long old = *p; if (v > old) *p = v; return old; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; return old; }
