// epi2_probe.cu -- settles the epilogue floor of the tcgen05 kernels:
//   (A) TMEM read rate: W warps (1 / 2 / 4 per lane quadrant) issue tcgen05.ld.32x32b.{x16,x32,x64}
//       back to back, D loads in flight before one tcgen05.wait::ld, no arithmetic, with and
//       without a second warp keeping the tensor pipe busy (M128 N128 K32 MMAs into the other 256
//       TMEM columns).
//   (B) MMA issue rate for narrow N (16 / 32 / 64; M = 128 needs N % 16 == 0) at M = 128, single accumulator and the
//       depthwise "diagonal" pattern (two N = 32 MMAs per tap on disjoint column halves).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi2_probe epi2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../f8net_b200/csrc/umma_common.cuh"

using namespace f8u;

#define LD_ASM_16(v, taddr)                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr) : "memory")
#define LD_ASM_32(v, taddr)                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr) : "memory")

// X = registers per load (16 | 32), D = loads in flight per wait (1 | 2 | 4; X*D <= 64 registers)
template <int X, int D>
__global__ void __launch_bounds__(544, 1) ld_kernel(int warps, int iters, int with_mma, long long *out, int *sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 128 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
    if (warp == 0) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 16) {
        // background MMAs: M128 N128 K32, no-swizzle operands, accumulators in columns 256..511
        if (with_mma && threadIdx.x == 16 * 32) {
            const uint32_t a_base = f8::smem_u32(smem), b_base = a_base + 128 * 128;
            const uint32_t idesc = instr_desc(false, 128);
            uint32_t phase = 0;
            while (!stop) {
                for (int k = 0; k < 8; ++k)
                    umma_i8(tmem + 256 + (uint32_t)((k & 1) * 128), smem_desc(a_base + (k & 3) * 2 * 2048, 2048, 128),
                            smem_desc(b_base + (k & 3) * 2 * 2048, 2048, 128), idesc, 1);
                umma_commit(f8::smem_u32(&bar));
                mbar_wait(f8::smem_u32(&bar), phase);
                phase ^= 1;
            }
        }
    } else if (warp < warps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int acc = 0;
        asm volatile("bar.sync 1, %0;" ::"r"(warps * 32) : "memory");
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            int32_t v[D][X];
            const uint32_t col0 = (uint32_t)((((warp >> 2) * 64 + it * X * D)) & 255);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const uint32_t col = (col0 + (uint32_t)(d * X)) & 255u;
                if (X == 16) LD_ASM_16(v[d], base + col);
                else LD_ASM_32(v[d], base + col);
            }
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < D; ++d) acc += v[d][0] ^ v[d][X - 1];
        }
        const long long t1 = clock64();
        asm volatile("bar.sync 1, %0;" ::"r"(warps * 32) : "memory");
        if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
        if (acc == 0x12345678) *sink = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// MMA issue rate: pattern 0 = one accumulator of N columns; 1 = alternate two disjoint N-column
// accumulators (the depthwise diagonal pair); 2 = pattern 1 with the A start address advanced by
// 32 B for the second half (as the DW kernel does)
template <int N>
__global__ void __launch_bounds__(64, 1) mma_kernel(int pattern, int iters, int swz, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a_base = f8::smem_u32(smem), b_base = a_base + 32 * 1024;
        const uint32_t idesc = instr_desc(false, N);
        // A: 64-byte-swizzled [slot][64 B] patch (as conv3x3) when swz, else no-swizzle [chunk][rows][16].
        // Every descriptor is built before the timed loop (the issue loop is MMA instructions only).
        uint64_t da[4], da2[4];
        for (int k = 0; k < 4; ++k) {
            const uint32_t off = (uint32_t)k * 64;
            da[k] = swz ? (smem_desc(a_base + off, 16, 512) | (4ull << 61)) : smem_desc(a_base + off, 2048, 128);
            const uint32_t off2 = off + (pattern == 2 ? 32u : 0u);
            da2[k] = swz ? (smem_desc(a_base + off2, 16, 512) | (4ull << 61)) : smem_desc(a_base + off2, 2048, 128);
        }
        const uint64_t b0 = smem_desc(b_base, 64 * 16, 128), b1 = smem_desc(b_base + 2 * 64 * 16 + 512, 64 * 16, 128);
        const long long t0 = clock64();
        if (pattern == 0) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tmem + (uint32_t)((it & 1) * 256), da[k], b0, idesc, 1);
            }
        } else {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_i8(tmem + (uint32_t)((it & 1) * 256), da[k], b0, idesc, 1);
                    umma_i8(tmem + (uint32_t)((it & 1) * 256 + N), da2[k], b1, idesc, 1);
                }
            }
        }
        umma_commit(f8::smem_u32(&bar));
        mbar_wait(f8::smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static long long *dout;
static int *sink;

template <int X, int D>
void run_ld() {
    const int iters = 4000;
    CK(cudaFuncSetAttribute(ld_kernel<X, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    for (int with_mma = 0; with_mma < 2; ++with_mma)
        for (int warps : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                ld_kernel<X, D><<<148, 544, 70 * 1024>>>(warps, iters, with_mma, dout, sink);
                CK(cudaDeviceSynchronize());
            }
            std::vector<long long> h(148);
            CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            const double bytes = (double)iters * D * X * 32 * 4 * warps;
            printf("ld 32x32b.x%-2d inflight=%d warps=%2d (%d per quadrant) mma=%d: %7.1f B/clk/SM  %6.1f cycles per wait-group per warp\n", X,
                   D, warps, warps / 4, with_mma, bytes / mx, (double)mx / iters);
        }
}

template <int N>
void run_mma() {
    const int iters = 2000;
    CK(cudaFuncSetAttribute(mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    for (int swz = 0; swz < 2; ++swz)
        for (int pattern = 0; pattern < 3; ++pattern) {
            if (pattern && 2 * N > 256) continue;
            for (int rep = 0; rep < 2; ++rep) {
                mma_kernel<N><<<148, 64, 70 * 1024>>>(pattern, iters, swz, dout);
                CK(cudaDeviceSynchronize());
            }
            std::vector<long long> h(148);
            CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto v : h) mx = v > mx ? v : mx;
            const double n_mma = (double)iters * 4 * (pattern ? 2 : 1);
            printf("mma M128 N=%3d K32 A=%s pattern=%d: %6.1f cycles per MMA (tensor floor %d, smem floor %.0f)\n", N,
                   swz ? "swizzle64" : "noswizzle", pattern, mx / n_mma, N / 2, (128 + N) / 4.0);
        }
}

int main() {
    CK(cudaMalloc(&dout, 148 * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    run_ld<16, 1>();
    run_ld<16, 2>();
    run_ld<16, 4>();
    run_ld<32, 1>();
    run_ld<32, 2>();
    run_mma<16>();
    run_mma<32>();
    run_mma<64>();
    run_mma<128>();
    return 0;
}
