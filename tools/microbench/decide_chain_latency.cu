// Dependent-issue latencies (SM cycles per operation, one warp alone on its SM) of the instructions on the LOCAL kernel's slot
// chain: what a single warp pays for each link.  nvcc -arch=sm_100a -O3 -o decide_chain_latency decide_chain_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define TIME(slot, ...)                                      \
    t0 = clock64();                                          \
    _Pragma("unroll 1") for (int i = 0; i < 64; i++) { __VA_ARGS__; } \
    t1 = clock64();                                          \
    if (threadIdx.x == 0) out[slot] = (t1 - t0) * 100 / 64;

__global__ void k(long long* out, long long seed, double dseed, const long long* __restrict__ chase, const long long* chase2) {
    long long q = seed + threadIdx.x, t0, t1;
    double acc = dseed + threadIdx.x, acc2 = dseed * 3, acc3 = dseed * 5;
    float f = (float) dseed;
    unsigned r = (unsigned) seed + threadIdx.x;
    __shared__ long long sh[64];
    sh[threadIdx.x & 63] = threadIdx.x;
    __syncthreads();
    TIME(0, acc = __fma_rn(acc, 1.0000001, 0.5))
    TIME(1, acc = __dadd_rn(acc, 0.5))
    TIME(2, acc = __dmul_rn(acc, 1.0000001))
    TIME(3, acc = __ll2double_rn(__double_as_longlong(acc) >> 12))                        // I2F.F64.S64 (+ shift)
    TIME(4, q = __double2ll_rn(__longlong_as_double(q | 0x4000000000000000LL)) & 0xFFFFF)   // F2I.S64.F64 (+ ops)
    TIME(5, f = __double2float_rn((double) f + 1.0))                                       // F2F.F64.F32 + DADD + F2F.F32.F64
    TIME(6, r = __reduce_add_sync(0xffffffffu, r) + 1)                                     // REDUX dependent
    TIME(7, { unsigned a = __reduce_add_sync(0xffffffffu, r), b = __reduce_add_sync(0xffffffffu, r + 1), c = __reduce_add_sync(0xffffffffu, r + 2),
              d = __reduce_add_sync(0xffffffffu, r + 3), e2 = __reduce_add_sync(0xffffffffu, r + 4), g = __reduce_add_sync(0xffffffffu, r + 5); r = a ^ b ^ c ^ d ^ e2 ^ g; })   // 6 independent REDUX
    TIME(8, r = __shfl_xor_sync(0xffffffffu, r, 1) + 1)
    TIME(9, r += __any_sync(0xffffffffu, r & 1))
    TIME(10, q = sh[q & 63])                                                                // LDS chain
    TIME(11, q = chase[q & 1023])                                                           // LDG.CONSTANT (nc) chain, L2/L1 hits
    TIME(12, { long long v; asm volatile("ld.global.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(chase2 + (q & 1023))); q = v; })   // LDG bypassing L1: L2 hit
    TIME(13, { acc = __fma_rn(acc, 1.0000001, 0.5); acc2 = __fma_rn(acc2, 1.0000001, 0.5); acc3 = __fma_rn(acc3, 1.0000001, 0.5); })   // 3 independent DFMA chains
    TIME(14, f = __fsqrt_rn(f + 1.0f))
    TIME(15, acc = __dsqrt_rn(acc + 1.0))
    TIME(16, f = __fmaf_rn(f, 1.0001f, 0.5f))
    TIME(17, q = q * 3 + 1)                                                                 // 64-bit IMAD chain
    TIME(18, { asm volatile("bar.arrive 1, 64;"); if (threadIdx.x >= 32) asm volatile("bar.sync 1, 64;"); else asm volatile("bar.sync 1, 64;"); r++; })
    if (threadIdx.x == 0) out[31] = (long long) (acc + acc2 + acc3 + f) + r + q;
}
int main() {
    long long *d, *c, *c2;
    cudaMalloc(&d, 256); cudaMalloc(&c, 8192); cudaMalloc(&c2, 8192);
    long long hc[1024];
    for (int i = 0; i < 1024; i++) hc[i] = (i * 37 + 11) & 1023;
    cudaMemcpy(c, hc, 8192, cudaMemcpyHostToDevice); cudaMemcpy(c2, hc, 8192, cudaMemcpyHostToDevice);
    for (int it = 0; it < 3; it++) k<<<1, 32>>>(d, 12345, 1.5, c, c2);
    long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA", "DADD", "DMUL", "I2F.F64.S64+SHF", "F2I.S64.F64+2 ops", "F2F f->d, DADD, F2F d->f", "REDUX dependent", "6 independent REDUX", "SHFL+IADD", "VOTE.ANY+IADD",
                           "LDS chain", "LDG nc chain", "LDG L1-bypass chain (L2 hit)", "3 independent DFMA", "fsqrt_rn+FADD", "dsqrt_rn+DADD", "FFMA", "IMAD64", "bar.arrive+bar.sync (1 warp)"};
    for (int i = 0; i < 19; i++) printf("%-34s %6.2f cycles\n", names[i], h[i] / 100.0);
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
