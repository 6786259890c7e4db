// epi_probe.cu -- issue cost of the fused requantisation epilogue on the SM's integer pipes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o epi_probe epi_probe.cu && ./epi_probe
// One CTA per SM, W warps; every thread requantises 16 int32 values per step (values come from a
// register recurrence so nothing is hoisted), REPS steps.  Prints cycles per step per warp and the
// SM-wide elements per clock for W = 4, 8, 16 warps and four variants of the math:
//   0: the library's epilogue16_plain_u8 (add, shift, tie test + fix, saturating pack)
//   1: no tie fix (round half up)          2: no pack (xor-reduce instead)
//   3: tie fix through  t' = y + half - 1 + bit_n(y)  (no predicate)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 step16(const int32_t (&v)[16], const int32_t *bh, int n) {
    const uint32_t mask = (1u << n) - 1u;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t t = (uint32_t)v[4 * q + j] + (uint32_t)bh[4 * q + j];
            if (MODE == 4) {
                // tie with an odd quotient <=> the low n+1 bits of t are exactly 2^n: step t down by one
                uint32_t tt = t;
                if ((t & (2u * mask + 1u)) == mask + 1u) tt = t - 1u;
                r[j] = (int32_t)tt >> n;
            } else if (MODE == 3) {
                const uint32_t b = (t >> n) & 1u;
                r[j] = (int32_t)(t + (mask >> 1) + b) >> n;
            } else {
                r[j] = (int32_t)t >> n;
                if (MODE != 1) { if ((t & mask) == 0u) r[j] &= ~1; }
            }
        }
        if (MODE == 2) {
            w[q] = (uint32_t)(r[0] ^ r[1] ^ r[2] ^ r[3]);
        } else {
            uint32_t hi;
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r[3]), "r"(r[2]), "r"(0));
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w[q]) : "r"(r[1]), "r"(r[0]), "r"(hi));
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long *out, int reps, int n) {
    __shared__ int32_t bias[64];
    __shared__ uint4 sink[512];
    if (threadIdx.x < 64) bias[threadIdx.x] = threadIdx.x * 977 + 12345;
    __syncthreads();
    int32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 131 + i * 7919;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        int32_t b[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const int4 x = *reinterpret_cast<const int4 *>(bias + ((r * 16 + i) & 63));
            b[i] = x.x; b[i + 1] = x.y; b[i + 2] = x.z; b[i + 3] = x.w;
        }
        const uint4 o = step16<MODE>(v, b, n);
        sink[threadIdx.x] = o;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += (int32_t)o.x + i;      // recurrence: next step depends on this one
        acc.x ^= o.y;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0 + (acc.x == 0x12345678u);
}

template <int MODE>
void run(int warps, long long *dev) {
    const int reps = 2000;
    probe<MODE><<<148, warps * 32>>>(dev, reps, 9);
    cudaDeviceSynchronize();
    probe<MODE><<<148, warps * 32>>>(dev, reps, 9);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_step = (double)mx / reps;
    printf("mode %d  warps %2d : %7.1f cycles / step / warp   %6.2f elements / clk / SM\n", MODE, warps, per_step,
           warps * 32 * 16 / per_step);
}

int main() {
    long long *dev;
    cudaMalloc(&dev, 148 * 16 * sizeof(long long));
    for (int w : {4, 8, 16}) { run<0>(w, dev); run<1>(w, dev); run<2>(w, dev); run<3>(w, dev); run<4>(w, dev); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
