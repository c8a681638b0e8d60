// Microbenchmark: FP64 FMA issue throughput per warp / per SM (cycles per warp-instruction).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
    double a0 = out[0], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double y = out[1];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        a0 = fma(a0, y, y); a1 = fma(a1, y, y); a2 = fma(a2, y, y); a3 = fma(a3, y, y);
        a4 = fma(a4, y, y); a5 = fma(a5, y, y); a6 = fma(a6, y, y); a7 = fma(a7, y, y);
        a0 = fma(a0, y, y); a1 = fma(a1, y, y); a2 = fma(a2, y, y); a3 = fma(a3, y, y);
        a4 = fma(a4, y, y); a5 = fma(a5, y, y); a6 = fma(a6, y, y); a7 = fma(a7, y, y);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[2 + blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, (1 << 20) * 8); cudaMalloc(&c, 16 * 8);
    double h[2] = {0.999, 1e-3}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    const int n = 4000;
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
        k<<<148, threads>>>(d, c, n); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, threads>>>(d, c, n); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        double flops = 148.0 * threads * 16.0 * n * 2;
        printf("threads/SM %4d: %.2f cycles per warp-DFMA (per warp), chip %.2f TFLOP/s\n", threads, hc / (16.0 * n), flops / (ms * 1e-3) / 1e12);
    }
    return 0;
}
