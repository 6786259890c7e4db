// mma3_probe.cu -- what sets the issue interval of tcgen05.mma.kind::i8 (M = 128, K = 32) for N = 16..256?
// The kernels measure ~74 cycles per N = 64 MMA and ~90 per N = 32 MMA where the operand-fetch model says
// max(N/2, (128+N)/4) = 48 / 40.  Variables: accumulators rotated (accs), MMAs in a row on one accumulator
// (run), same / distinct B descriptor per MMA, A layout (no swizzle chunk-major vs 64-byte swizzle), a
// second warp streaming tcgen05.ld from the other TMEM half, and bulk copies writing shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma3_probe mma3_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../f8net_b200/csrc/umma_common.cuh"

using namespace f8u;

template <int N, int ACCS, int RUN>
__global__ void __launch_bounds__(192, 1) k(int sameb, int swz, int bg, int iters, long long *out,
                                            const uint8_t *gsrc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, cbar;
    __shared__ uint32_t tslot;
    __shared__ volatile int stop;
    const uint32_t a_base = f8::smem_u32(smem), b_base = a_base + 48 * 1024, scratch = a_base + 112 * 1024;
    for (int i = threadIdx.x; i < 112 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); mbar_init(f8::smem_u32(&cbar), 1); fence_barrier_init(); stop = 0; }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(false, N);
        // A: swz ? [slot][64 B] SWIZZLE_64B (SBO = 512, start advanced by taps) : chunk-major [chunk][256 rows][16]
        // B: chunk-major [chunk][N rows][16], K step = two chunks further
        uint64_t da[8], db[8];
        for (int j = 0; j < 8; ++j) {
            const uint32_t ao = swz ? (uint32_t)(j >> 1) * 64u * 3u + (uint32_t)(j & 1) * 32u : (uint32_t)(j & 3) * 2u * 256u * 16u + (uint32_t)(j >> 2) * 64u;
            da[j] = swz ? (smem_desc(a_base + ao, 16, 512) | (4ull << 61)) : smem_desc(a_base + ao, 256 * 16, 128);
            const uint32_t bo = sameb ? 0u : (uint32_t)j * 2u * N * 16u;
            db[j] = smem_desc(b_base + bo, N * 16, 128);
        }
        constexpr int NACC = 512 / N < ACCS ? 512 / N : ACCS;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) umma_i8(tmem + (uint32_t)(((j / RUN) % NACC) * N), da[j], db[j], idesc, 1);
        }
        umma_commit(f8::smem_u32(&bar));
        mbar_wait(f8::smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
        stop = 1;
    } else if (warp >= 1 && warp <= 4 && (bg & 1)) {
        // background tcgen05.ld: four warps (one per lane quadrant) reading 16 columns at a time
        const uint32_t base = tmem + ((uint32_t)(((warp - 1) & 3) * 32) << 16);
        int acc = 0, c = 0;
        while (!stop) {
            int32_t v[16];
            tmem_ld16(base + (uint32_t)c, v);
            tmem_ld_wait();
            acc += v[0] ^ v[15];
            c = (c + 16) & 255;
        }
        if (acc == 0x12345678) out[200] = acc;
    } else if (warp == 5 && (bg & 2) && (threadIdx.x & 31) == 0) {
        // background bulk copies global -> shared (as a weight loader does), 8 KB at a time
        uint32_t ph = 0;
        while (!stop) {
            mbar_expect_tx(f8::smem_u32(&cbar), 8192);
            mbar_arrive(f8::smem_u32(&cbar));
            for (int j = 0; j < 4; ++j) bulk_g2s(scratch + j * 2048, gsrc + ((blockIdx.x * 4 + j) & 63) * 2048, 2048, f8::smem_u32(&cbar));
            mbar_wait(f8::smem_u32(&cbar), ph);
            ph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
static long long *dout;
static uint8_t *gsrc;

template <int N, int ACCS, int RUN>
void run_n() {
    const int iters = 1000;
    CK(cudaFuncSetAttribute(k<N, ACCS, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    struct Cfg { int accs, run, sameb, swz, bg; };
    const Cfg cfgs[] = {{ACCS, RUN, 0, 0, 0}, {ACCS, RUN, 1, 0, 0}, {ACCS, RUN, 0, 1, 0}, {ACCS, RUN, 0, 1, 1}, {ACCS, RUN, 0, 1, 2},
                        {ACCS, RUN, 0, 1, 3}};
    for (const Cfg &c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            k<N, ACCS, RUN><<<148, 192, 130 * 1024>>>(c.sameb, c.swz, c.bg, iters, dout, gsrc);
            CK(cudaDeviceSynchronize());
        }
        std::vector<long long> h(148);
        CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("N=%3d accs=%d run=%d sameB=%d A=%s bg(ld=%d,copy=%d): %6.1f cycles/MMA (model %g)\n", N, c.accs, c.run, c.sameb,
               c.swz ? "swz64" : "plain", c.bg & 1, (c.bg >> 1) & 1, (double)mx / (iters * 8.0), N / 2.0 > (128 + N) / 4.0 ? N / 2.0 : (128 + N) / 4.0);
    }
}

int main() {
    CK(cudaMalloc(&dout, 256 * sizeof(long long)));
    CK(cudaMalloc(&gsrc, 64 * 2048));
    CK(cudaMemset(gsrc, 1, 64 * 2048));
    run_n<16, 1, 1>(); run_n<16, 2, 1>(); run_n<16, 4, 2>();
    run_n<32, 1, 1>(); run_n<32, 2, 1>(); run_n<32, 4, 2>(); run_n<32, 8, 1>();
    run_n<64, 1, 1>(); run_n<64, 2, 1>(); run_n<64, 4, 2>(); run_n<64, 4, 1>();
    run_n<128, 1, 1>(); run_n<128, 2, 1>(); run_n<128, 2, 2>(); run_n<128, 4, 2>();
    run_n<256, 1, 1>(); run_n<256, 2, 2>();
    return 0;
}
