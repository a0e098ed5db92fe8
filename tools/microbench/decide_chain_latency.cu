#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, long long seed, double dseed) {
    long long q = seed + threadIdx.x;
    long long t0, t1;
    double acc = dseed;
    // I2F.F64.S64 dependent chain
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) { acc = __ll2double_rn(q) + acc * 0.0; q = (long long) acc + 1; }   // I2F + F2I chain
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 64;
    // DFMA chain
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) acc = __fma_rn(acc, 1.0000001, 0.5);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0) / 64;
    // I2F.S64 only (independent of previous result except via integer add)
    long long s = q;
    double sum = 0;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) { double d = __ll2double_rn(s); s += __double_as_longlong(d) & 1; sum += d; }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0) / 64;
    // 32-bit split conversion
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) { double d = __fma_rn((double) (int) (s >> 32), 4294967296.0, (double) (unsigned) s); s += __double_as_longlong(d) & 1; sum += d; }
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0) / 64;
    // shared atomicAdd 64
    __shared__ unsigned long long sh[4];
    if (threadIdx.x == 0) sh[0] = 0;
    __syncthreads();
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) if (threadIdx.x == 0) atomicAdd(&sh[0], (unsigned long long) s);
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0) / 64;
    // REDUX
    unsigned r = (unsigned) s;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) r = __reduce_add_sync(0xffffffffu, r) + 1;
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = (t1 - t0) / 64;
    // LDS dependent chain
    __shared__ int idx[64];
    idx[threadIdx.x & 63] = (threadIdx.x + 1) & 63;
    __syncthreads();
    int j = threadIdx.x & 63;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) j = idx[j];
    t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0) / 64;
    // S2R CgaCtaId
    unsigned cg = 0;
    t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 64; i++) { unsigned v; asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(v)); cg += v; }
    t1 = clock64();
    if (threadIdx.x == 0) out[7] = (t1 - t0) / 64;
    if (threadIdx.x == 0) out[8] = (long long) (acc + sum) + r + j + cg + sh[0];
}
int main() {
    long long* d; cudaMalloc(&d, 128); 
    for (int it = 0; it < 2; it++) k<<<1, 32>>>(d, 12345, 1.5);
    long long h[9]; cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost);
    printf("cycles per iteration: I2F.S64+F2I chain %lld, DFMA %lld, I2F.F64.S64 loop %lld, split32 conversion loop %lld, shared atomicAdd64 %lld, REDUX %lld, LDS chain %lld, S2R cga %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    return 0;
}
