// Microbenchmark: the systolic "front" step (rotg by one lane -> shuffle -> apply) in isolation.
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 cplx;
__device__ __forceinline__ cplx mk(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) { return mk(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))); }
__device__ __forceinline__ cplx shfl_c(cplx v, int s) { v.x = __shfl_sync(0xffffffffu, v.x, s); v.y = __shfl_sync(0xffffffffu, v.y, s); return v; }
__global__ void k(cplx* H, long long* cyc, int nsteps, int variant) {
    extern __shared__ cplx sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = H[i];
    __syncthreads();
    if (warp > 0) {               // optional interference: spinning warps
        volatile cplx* v = sm;
        if (variant == 1) { while (v[8191].x < 1e300) { } }
        return;
    }
    cplx carry = sm[lane], g = sm[64 + lane], low = sm[128 + lane];
    long long t0 = clock64();
    if (variant < 2) {
    for (int k = 0; k < nsteps; ++k) {
            const int src = k & 31;
            cplx a = mk(1, 0), b = mk(0, 0);
            if (lane == src) {
                double r2 = carry.x * carry.x + carry.y * carry.y + g.x * g.x + g.y * g.y;
                double ir = rsqrt(r2);
                a = mk(carry.x * ir, -carry.y * ir); b = mk(g.x * ir, -g.y * ir);
            }
            a = shfl_c(a, src); b = shfl_c(b, src);
            if (lane != src) {
                cplx lw = low;
                low = sm[256 + ((k * 32 + lane) & 4095)];
                sm[4400 + lane] = cfma(a, carry, cmul(b, lw));
                carry = csub(cmulc(lw, a), cmulc(carry, b));
            } else {
                sm[4500 + (k & 63)] = a; sm[4600 + (k & 63)] = b;
                carry = mk(carry.x * 0.5 + 1.0, carry.y);
            }
        }
    } else {
        for (int k = 0; k < nsteps; ++k) {
            const int src = k & 31;
            // broadcast the generator's carry, every lane forms the rotation redundantly
            const cplx cs = shfl_c(carry, src);
            const cplx gs = sm[64 + src];                    // broadcast LDS, independent of the chain
            double r2 = cs.x * cs.x + cs.y * cs.y + gs.x * gs.x + gs.y * gs.y;
            double ir = (variant == 3) ? r2 * 1e-3 : rsqrt(r2);
            const cplx a = mk(cs.x * ir, -cs.y * ir), b = mk(gs.x * ir, -gs.y * ir);
            cplx lw = low;
            low = sm[256 + ((k * 32 + lane) & 4095)];
            cplx top = cfma(a, carry, cmul(b, lw));
            cplx nc = csub(cmulc(lw, a), cmulc(carry, b));
            if (lane != src) { sm[4400 + lane] = top; carry = nc; }
            else { sm[4500 + (k & 63)] = a; sm[4600 + (k & 63)] = b; carry = mk(carry.x * 0.5 + 1.0, carry.y); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; sm[8191].x = 1e301; }
    H[lane] = carry;
    if (lane == 0) ((volatile cplx*)sm)[8191].x = 1e301;
}
int main() {
    cplx* d; long long* c; cudaMalloc(&d, 8192 * 16); cudaMalloc(&c, 64);
    cplx* h = new cplx[8192]; for (int i = 0; i < 8192; ++i) h[i] = make_double2(1.0 + i % 7, 0.5 - i % 3);
    cudaMemcpy(d, h, 8192 * 16, cudaMemcpyHostToDevice);
    const int n = 4000;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    for (int variant = 0; variant < 4; ++variant)
        for (int threads : {32, 256}) {
            cudaMemcpy(d, h, 8192 * 16, cudaMemcpyHostToDevice);
            k<<<1, threads, 8192 * 16>>>(d, c, n, variant); cudaDeviceSynchronize();
            long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
            printf("variant %d threads %d: %.1f cycles per front step  (%s)\n", variant, threads, hc / (double)n, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
