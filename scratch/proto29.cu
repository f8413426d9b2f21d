
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>
struct fe29 { uint32_t v[9]; };
#define M29 0x1fffffffu
__host__ __device__ __forceinline__ uint32_t qm(int i) {
    return i==0?0x187cfd47U:i==1?0x10460b6U:i==2?0x1c72a34fU:i==3?0x2d522d0U:i==4?0x1585d978U:i==5?0x2db40c0U:i==6?0xa6e141U:i==7?0xe5c2634U:0x30644eU;
}
#define QINV 0x4866389U
// a: limbs < 2^30 (lazy), b: limbs < 2^29.  result: limbs < 2^29 (top limb small), value < 2^255, = a*b*2^-261 mod p
__host__ __device__ __forceinline__ fe29 mul29(const fe29 &a, const fe29 &b) {
    uint64_t t[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) t[j] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] += (uint64_t)a.v[j] * b.v[i];
        uint32_t q = ((uint32_t)t[0] * QINV) & M29;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] += (uint64_t)q * qm(j);
        uint64_t c = t[0] >> 29;
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] = t[j + 1];
        t[9] = 0;
        t[0] += c;
    }
    fe29 r;
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        uint64_t v = t[j] + c;
        r.v[j] = (uint32_t)v & M29;
        c = v >> 29;
    }
    r.v[8] += (uint32_t)(c << 29);   // keep any excess in the top limb (value < 2^255 so c == 0)
    return r;
}
template <int ILP> __global__ void __launch_bounds__(256) probe(fe29 *out, uint32_t iters) {
    fe29 x[ILP], y[ILP];
    for (int k = 0; k < ILP; ++k) for (int j = 0; j < 9; ++j) { x[k].v[j] = (threadIdx.x * 2654435761u + j * 40503u + k) & M29; y[k].v[j] = (blockIdx.x * 97u + j * 7919u + 3 * k + 1) & M29; }
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = mul29(x[k], y[k]);
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k) for (int j = 0; j < 9; ++j) s ^= x[k].v[j];
    if (s == 0x12345678u) out[0] = x[0];
}
int main(int argc, char **argv) {
    if (argc > 1) {   // host check: read 18 hex limbs, print product limbs
        fe29 a, b;
        for (int j = 0; j < 9; ++j) a.v[j] = strtoul(argv[1 + j], 0, 16);
        for (int j = 0; j < 9; ++j) b.v[j] = strtoul(argv[10 + j], 0, 16);
        fe29 r = mul29(a, b);
        for (int j = 0; j < 9; ++j) printf("%x ", r.v[j]);
        printf("\n");
        return 0;
    }
    fe29 *d; cudaMalloc(&d, 64);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ilp = 1; ilp <= 2; ++ilp) {
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            unsigned blocks = prop.multiProcessorCount * 8, iters = 4096 / ilp;
            cudaEventRecord(e0);
            if (ilp == 1) probe<1><<<blocks, 256>>>(d, iters); else probe<2><<<blocks, 256>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double rate = (double)blocks * 256 * iters * ilp / (ms * 1e-3);
            if (rep && rate > best) best = rate;
        }
        printf("mul29 ilp%d: %.2f G mul/s (err %s)\n", ilp, best / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
