// Dependent-chain latencies of the instructions the ScreenPressor symbol chain is made of, one warp, sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat tools/microbench/lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
template <int OP>
__global__ void k(uint32_t *out, long long *cyc, uint32_t seed)
{
    __shared__ uint32_t sm[64];
    sm[threadIdx.x] = threadIdx.x * 4 % 128; sm[threadIdx.x + 32] = threadIdx.x;
    __syncwarp();
    uint32_t v = seed + threadIdx.x, w = seed;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
#pragma unroll
        for (int i = 0; i < N / 16; i++) {
            if (OP == 0) v = v * 3u + w;                                                        // IMAD
            if (OP == 1) v = __shfl_sync(0xffffffffu, v, (v >> 3) & 31) + 1;                    // SHFL + IADD
            if (OP == 2) v = __popc(__ballot_sync(0xffffffffu, v & 1)) + v;                     // VOTE + POPC + IADD
            if (OP == 3) v = __reduce_max_sync(0xffffffffu, v) + threadIdx.x;                   // REDUX.MAX + IADD
            if (OP == 4) v = __clz((int)v) + v + 1;                                             // FLO + 2 IADD
            if (OP == 5) v = sm[(v & 31)] + 1;                                                  // LDS + IADD (address dependent)
            if (OP == 6) v = __umulhi(v, w) + 7u;                                               // IMAD.HI
            if (OP == 7) v = __ballot_sync(0xffffffffu, v & 1) + v;                             // VOTE + IADD
            if (OP == 8) v = __reduce_add_sync(0xffffffffu, v) + threadIdx.x;                   // REDUX.SUM
            if (OP == 9) { if (v & 4) v += 3; else v = v * 5 + 1; }                             // (likely predicated)
            if (OP == 10) v = __byte_perm(v, w, 0x0123) + 1;                                    // PRMT
            if (OP == 11) v = __funnelshift_l(v, w, v) + 1;                                     // SHF
            if (OP == 12) v = min(v, w) + 1;                                                    // VIMNMX
            if (OP == 13) v = __match_any_sync(0xffffffffu, v & 3) + v;                         // MATCH
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
// data-dependent uniform branch cost
__global__ void kbr(uint32_t *out, long long *cyc, uint32_t seed)
{
    uint32_t v = seed, a = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N; it++) {
        v = v * 1664525u + 1013904223u;
        if (__ballot_sync(0xffffffffu, (v >> 16) & 1)) { a += v; a ^= a >> 3; a *= 3; a += 11; a ^= a << 2; a += v >> 5; a *= 7; a ^= 0x55; }
        else { a -= v; a ^= a >> 5; a *= 5; a += 13; a ^= a << 3; a += v >> 7; a *= 9; a ^= 0xAA; }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[20] = t1 - t0;
}
__global__ void knobr(uint32_t *out, long long *cyc, uint32_t seed)
{
    uint32_t v = seed, a = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N; it++) {
        v = v * 1664525u + 1013904223u;
        a += v; a ^= a >> 3; a *= 3; a += 11; a ^= a << 2; a += v >> 5; a *= 7; a ^= 0x55;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[21] = t1 - t0;
}
int main()
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 32 * 8);
    for (int r = 0; r < 2; r++) {
        k<0><<<1, 32>>>(out, cyc, 3); k<1><<<1, 32>>>(out, cyc, 3); k<2><<<1, 32>>>(out, cyc, 3); k<3><<<1, 32>>>(out, cyc, 3);
        k<4><<<1, 32>>>(out, cyc, 3); k<5><<<1, 32>>>(out, cyc, 3); k<6><<<1, 32>>>(out, cyc, 3); k<7><<<1, 32>>>(out, cyc, 3);
        k<8><<<1, 32>>>(out, cyc, 3); k<9><<<1, 32>>>(out, cyc, 3); k<10><<<1, 32>>>(out, cyc, 3); k<11><<<1, 32>>>(out, cyc, 3);
        k<12><<<1, 32>>>(out, cyc, 3); k<13><<<1, 32>>>(out, cyc, 3);
        kbr<<<1, 32>>>(out, cyc, 3); knobr<<<1, 32>>>(out, cyc, 3);
        cudaDeviceSynchronize();
    }
    const char *names[] = {"IMAD", "SHFL.IDX+IADD", "VOTE+POPC+IADD", "REDUX.MAX+IADD", "FLO+2 IADD", "LDS+IADD", "IMAD.HI+IADD", "VOTE+IADD",
                           "REDUX.SUM+IADD", "if/else small", "PRMT+IADD", "SHF+IADD", "VIMNMX+IADD", "MATCH+IADD"};
    for (int i = 0; i < 14; i++) printf("%-16s %6.1f cycles per dependent step\n", names[i], (double)cyc[i] / N);
    printf("loop with a data-dependent uniform branch (8-op bodies) %6.1f cycles per iteration; same work without branch %6.1f\n",
           (double)cyc[20] / N, (double)cyc[21] / N);
    return 0;
}
