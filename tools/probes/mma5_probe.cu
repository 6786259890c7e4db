// mma5_probe.cu -- cost of each piece of a pipeline-stage boundary in the MMA issue loop (nothing is hidden:
// mma4_probe shows the tensor pipe does not run ahead of the issuing thread).  Bursts of 24 M128 N64 K32 MMAs;
// between bursts, selected by `mode` bits:
//   1  tcgen05.commit to a ring barrier (stage release)          2  mbarrier try_wait on a barrier completed long ago
//   4  tcgen05.fence::after_thread_sync                          8  whole warp loops, elect_one per burst (else lane 0 alone)
//   16 descriptors recomputed from a loop-carried slot index      32 __syncwarp after the burst
//   128 a plain ld.volatile.shared of a flag word instead of the mbarrier wait (is it generic shared-memory latency?)
//   256 a second warp does the mbarrier wait; the MMA warp meets it on a named barrier (bar.sync 1, 64): no shared-memory op
//       on the issuing warp's path
//   64 the wait of burst k+1 is a non-blocking mbarrier.test_wait issued in the middle of burst k (blocking wait only if that failed)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma5_probe mma5_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../f8net_b200/csrc/umma_common.cuh"

using namespace f8u;

constexpr int N = 64, G = 24, RING = 4;

template <int MODE>
__global__ void __launch_bounds__(64, 1) k(int iters, long long *out, int rnd) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, ring[RING];
    __shared__ uint32_t tslot;
    const uint32_t a_base = f8::smem_u32(smem), b_base = a_base + 48 * 1024;
    for (int i = threadIdx.x; i < 112 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + (uint32_t)blockIdx.x * 40503u;      // rnd: pseudo-random bytes (full toggle rate)
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        ((uint32_t *)smem)[i] = rnd == 1 ? h : (rnd == 2 ? 0u : 0x01010101u * (i & 3));
    }
    if (threadIdx.x == 0) {
        mbar_init(f8::smem_u32(&bar), 1);
        for (int s = 0; s < RING; ++s) mbar_init(f8::smem_u32(&ring[s]), 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32 && ((MODE & 8) || lane == 0)) {
        const uint32_t idesc = instr_desc(false, N);
        constexpr uint32_t hi_a = (512u >> 4) | (1u << 14) | (4u << 29), hi_b = (128u >> 4) | (1u << 14);
        const uint32_t lbo_b = ((uint32_t)(N * 16) >> 4) << 16;
        int slot = 0, phase = 0;
        uint32_t a_lo = ((a_base & 0x3ffffu) >> 4) | (1u << 16), b_lo = ((b_base & 0x3ffffu) >> 4) | lbo_b;
        uint32_t peeked = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (MODE & 2) {            // wait for the release of this slot RING bursts ago (completed long ago)
                if (it >= RING && !peeked) mbar_wait(f8::smem_u32(&ring[slot]), (uint32_t)(phase ^ 1));
            }
            if (MODE & 128) {
                uint32_t v;
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(f8::smem_u32(&tslot)) : "memory");
                if (v == 0xdeadbeefu) break;
            }
            if (MODE & 256) asm volatile("bar.sync 1, 64;" ::: "memory");
            if (MODE & 4) tc_fence_after();
            if (MODE & 16) {
                a_lo = (((a_base + (uint32_t)slot * 4096u) & 0x3ffffu) >> 4) | (1u << 16);
                b_lo = (((b_base + (uint32_t)slot * 3u * 4096u) & 0x3ffffu) >> 4) | lbo_b;
            }
            if (!(MODE & 8) || elect_one()) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    umma_i8_lohi(tmem + (uint32_t)((j & 3) * N), a_lo + (uint32_t)((j % 6) * 4), hi_a, b_lo + (uint32_t)((j & 1) * 2 * N), hi_b,
                                 idesc, 1);
                    if ((MODE & 64) && j == 3) {
                        const int ns = slot + 1 == RING ? 0 : slot + 1;
                        const uint32_t np = (uint32_t)((slot + 1 == RING ? phase ^ 1 : phase) ^ 1);
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(peeked) : "r"(f8::smem_u32(&ring[ns])), "r"(np) : "memory");
                        if (it + 1 < RING) peeked = 1;
                    }
                }
                if (MODE & 1) umma_commit(f8::smem_u32(&ring[slot]));
            }
            if (MODE & 32) __syncwarp();
            if (++slot == RING) { slot = 0; phase ^= 1; }
        }
        if (lane == 0) {
            umma_commit(f8::smem_u32(&bar));
            mbar_wait(f8::smem_u32(&bar), 0);
            out[blockIdx.x] = clock64() - t0;
        }
    }
    if ((MODE & 256) && threadIdx.x >= 32) {
        // scout warp: waits for each stage's barrier (released RING bursts ago by the MMA warp's commits), then meets the MMA warp
        int slot = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (it >= RING) mbar_wait(f8::smem_u32(&ring[slot]), (uint32_t)(phase ^ 1));
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (++slot == RING) { slot = 0; phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
static long long *dout;

template <int MODE>
void run(int rnd = 0) {
    const int iters = 500;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    for (int rep = 0; rep < 2; ++rep) {
        k<MODE><<<148, 64, 130 * 1024>>>(iters, dout, rnd);
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    printf("data=%s mode %3d [%s%s%s%s%s%s]: %7.1f cycles/burst (back-to-back %d) -> +%.0f per stage boundary\n",
           rnd == 1 ? "random" : (rnd == 2 ? "zeros" : "pattern"), MODE, MODE & 1 ? "commit " : "",
           MODE & 2 ? "wait " : "", MODE & 4 ? "fence " : "", MODE & 8 ? "warp+elect " : "", MODE & 16 ? "descs " : "",
           MODE & 32 ? "syncwarp " : (MODE & 128 ? "ld.shared " : (MODE & 256 ? "scout+bar.sync " : "")), (double)mx / iters, G * 48, (double)mx / iters - G * 48);
}

int main() {
    CK(cudaMalloc(&dout, 256 * sizeof(long long)));
    run<0>(); run<1>(); run<3>(); run<7>(); run<16>(); run<8>(); run<9>(); run<11>(); run<15>(); run<31>(); run<63>(); run<23>(); run<128 + 1>(); run<128 + 9>(); run<256 + 9>(); run<256 + 8 + 1 + 4 + 16>();
    run<0>(1); run<0>(2); run<9>(1); run<256 + 8 + 1 + 4 + 16>(1); run<31>(1);
    return 0;
}
