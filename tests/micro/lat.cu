// Microbenchmark: dependent-op latencies on the target GPU (cycles).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = threadIdx.x * 1.0;
    __syncthreads();
    double x = out[0], y = out[1];
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // rsqrt chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { x = rsqrt(x + 1.5); x = rsqrt(x + 1.5); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // shuffle chain (double = 2 shfl)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // LDS dependent chain
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { idx = (int)sm[idx & 1023] & 1023; idx = (int)sm[idx] & 1023; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // barrier
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { __syncthreads(); __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // STS -> barrier -> LDS round trip
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { sm[threadIdx.x] = x; __syncthreads(); x += sm[(threadIdx.x + 33) % blockDim.x]; __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // double divide chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { x = y / (x + 2.0); x = y / (x + 2.0); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // membar
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { sm[threadIdx.x] = x; __threadfence_block(); sm[threadIdx.x] = x + 1; __threadfence_block(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    // fp32 rsqrt + FFMA chain
    float fx = (float)x;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { fx = rsqrtf(fx + 1.5f); fx = rsqrtf(fx + 1.5f); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    out[2 + threadIdx.x] = x + idx + fx;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 4096 * 8); cudaMalloc(&c, 16 * 8);
    double h[2] = {0.999, 1e-3}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    const int n = 2000;
    for (int threads : {32, 128, 256, 512}) {
        k<<<1, threads>>>(d, c, n); cudaDeviceSynchronize();
        long long hc[16]; cudaMemcpy(hc, c, 16 * 8, cudaMemcpyDeviceToHost);
        printf("threads %3d: DFMA %.1f  rsqrt(+add) %.1f  shfl64 %.1f  LDS-dep %.1f  bar %.1f  sts-bar-lds-bar %.1f  ddiv %.1f  sts+membar %.1f  rsqrtf %.1f\n", threads,
               hc[0] / (4.0 * n), hc[1] / (2.0 * n), hc[2] / (2.0 * n), hc[3] / (2.0 * n), hc[4] / (2.0 * n), hc[5] / (1.0 * n), hc[6] / (2.0 * n), hc[7] / (2.0 * n), hc[8] / (2.0 * n));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
