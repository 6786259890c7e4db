// mma_probe.cu -- micro-benchmark + semantics probe for tcgen05.mma.kind::i8 operand fetch.
//   (1) rate: cycles per K=32 MMA (M=128, N in {64,128,256}) for shared-memory operand layouts
//       SWIZZLE_NONE / 32B / 64B / 128B (K-major), all 148 SMs busy.
//   (2) semantics: with a swizzled patch stored by ABSOLUTE address bits, does a descriptor whose
//       start address is shifted by d rows (not a multiple of the swizzle atom) read rows m+d?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../f8net_b200/csrc/umma_common.cuh"

using namespace f8u;

// layout modes: 0 none, 1 = 32B swizzle, 2 = 64B, 3 = 128B
__host__ __device__ inline uint32_t row_bytes_of(int mode) { return mode == 0 ? 16u : (16u << mode); }
__host__ __device__ inline uint64_t layout_bits(int mode) {
    // sm_100 descriptor bits [61,64): 0 none, 6 = 32B, 4 = 64B, 2 = 128B
    const uint64_t t = mode == 0 ? 0 : (mode == 1 ? 6 : (mode == 2 ? 4 : 2));
    return t << 61;
}
// byte offset (relative to a 1024-aligned base) of K byte k of row r; swizzled modes store a
// [rows][row_bytes] image with 16-byte chunks XORed by absolute address bits
__host__ __device__ inline uint32_t elem_off(int mode, int rows, int r, int k) {
    if (mode == 0) return (uint32_t)((k >> 4) * rows * 16 + r * 16 + (k & 15));
    const uint32_t rb = row_bytes_of(mode);
    uint32_t a = (uint32_t)r * rb + (uint32_t)k;
    const uint32_t xmask = (rb / 16 - 1);               // 1, 3, 7 chunks
    a ^= ((a >> 7) & xmask) << 4;
    return a;
}
__device__ inline uint64_t make_desc(int mode, uint32_t addr, int rows) {
    uint32_t lbo, sbo;
    if (mode == 0) { lbo = (uint32_t)rows * 16; sbo = 128; }
    else { lbo = 16; sbo = 8 * row_bytes_of(mode); }
    return smem_desc(addr, lbo, sbo) | layout_bits(mode);
}

// ---------------------------------------------------------------- (1) rate
template <int N>
__global__ void __launch_bounds__(160, 1) rate_kernel(int mode, int iters, int kper, int dshift, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t a_base = f8::smem_u32(smem);
    const uint32_t b_base = a_base + 160 * 128;           // A: 160 rows x 128 B of K
    for (int i = threadIdx.x; i < (160 + N) * 128 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(false, N);
        // K step of 32 bytes: none: two chunks further (2 * LBO); swizzled: +32 B inside the row
        // (row_bytes >= 32), kper = MMAs per row sweep
        uint64_t da[4], db[4];
        for (int k = 0; k < 4; ++k) {
            uint32_t ao, bo;
            if (mode == 0) { ao = (uint32_t)k * 2 * 160 * 16; bo = (uint32_t)k * 2 * N * 16; }
            else { ao = bo = (uint32_t)(k % kper) * 32; }
            da[k] = make_desc(mode, a_base + ao + (uint32_t)dshift * row_bytes_of(mode), 160);
            db[k] = make_desc(mode, b_base + bo, N);
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tmem + (uint32_t)((it & (512 / N - 1)) * N);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_i8(acc, da[k], db[k], idesc, 1);
        }
        umma_commit(f8::smem_u32(&bar));
        mbar_wait(f8::smem_u32(&bar), 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------- (2) shifted start address
// A patch: 192 rows x (row bytes) ; B: 64 rows x 32 B of K with B[n][k] = (n == k) for n < 32.
// One MMA (M=128, N=64, K=32) with A start = row d, K offset koff: D[m][n] should be A[m+d][koff+n].
__global__ void __launch_bounds__(128, 1) shift_kernel(int mode, int d, int koff, int base_off_field,
                                                       const uint8_t *a_img, int a_bytes, int32_t *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int AROWS = 192;
    uint8_t *A = smem;
    uint8_t *B = smem + 32768;
    for (int i = threadIdx.x; i < a_bytes; i += blockDim.x) A[i] = a_img[i];
    for (int i = threadIdx.x; i < 64 * 128; i += blockDim.x) B[i] = 0;
    __syncthreads();
    if (threadIdx.x < 32) B[elem_off(mode, 64, threadIdx.x, threadIdx.x)] = 1;
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t rb = row_bytes_of(mode);
        uint32_t a_addr = f8::smem_u32(A), b_addr = f8::smem_u32(B);
        if (mode == 0) a_addr += (uint32_t)d * 16 + (uint32_t)(koff >> 4) * AROWS * 16;
        else a_addr += (uint32_t)d * rb + (uint32_t)koff;
        uint64_t ad = make_desc(mode, a_addr, AROWS);
        ad |= (uint64_t)(base_off_field & 7) << 49;
        umma_i8(tmem, ad, make_desc(mode, b_addr, 64), instr_desc(false, 64), 0);
        umma_commit(f8::smem_u32(&bar));
    }
    mbar_wait(f8::smem_u32(&bar), 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < 32; c += 16) {
        int32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 32 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int N>
void run_rate(int mode, int dshift, long long *dout) {
    const int iters = 2000;
    const int kper = mode == 0 ? 4 : (int)(row_bytes_of(mode) / 32);
    const size_t smem = (160 + N) * 128 + 2048;
    CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int grid : {148}) {
        rate_kernel<N><<<grid, 160, smem>>>(mode, iters, kper, dshift, dout);
        CK(cudaDeviceSynchronize());
        rate_kernel<N><<<grid, 160, smem>>>(mode, iters, kper, dshift, dout);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), dout, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0, mn = 1ll << 60;
        for (auto v : h) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
        const double n_mma = (double)iters * 4;
        printf("rate N=%3d mode=%d dshift=%d grid=%3d: %.1f .. %.1f cycles per K=32 MMA (floor %d; smem-read floor %.0f)\n", N, mode, dshift,
               grid, mn / n_mma, mx / n_mma, N / 2, (128 + N) * 32 / 128.0);
    }
}

int main() {
    long long *dout;
    CK(cudaMalloc(&dout, 148 * sizeof(long long)));
    for (int mode = 0; mode < 4; ++mode) {
        for (int d : {0, 1, 3, 4, 8}) {
            run_rate<64>(mode, d, dout);
            run_rate<128>(mode, d, dout);
            run_rate<256>(mode, d, dout);
        }
    }
    // shift semantics
    const int AROWS = 192;
    uint8_t *a_dev;
    int32_t *o_dev;
    CK(cudaMalloc(&a_dev, 32768));
    CK(cudaMalloc(&o_dev, 128 * 32 * 4));
    CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int mode = 0; mode < 4; ++mode) {
        const int rb = mode == 0 ? 64 : (int)row_bytes_of(mode);   // K bytes stored per row
        std::vector<uint8_t> img(32768, 0), val(AROWS * rb);
        for (int r = 0; r < AROWS; ++r)
            for (int k = 0; k < rb; ++k) {
                const uint8_t v = (uint8_t)((r * 37 + k * 11 + 5) % 251);
                val[r * rb + k] = v;
                img[elem_off(mode, AROWS, r, k)] = v;
            }
        CK(cudaMemcpy(a_dev, img.data(), 32768, cudaMemcpyHostToDevice));
        for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
            int ok = 0, total = 0;
            char detail[256] = "";
            for (int d = 0; d < 20; ++d)
                for (int koff = 0; koff + 32 <= rb; koff += 32) {
                    // bo_mode 1: descriptor base_offset field = (start address >> 7) & 7
                    const int rbb = mode == 0 ? 16 : rb;
                    const int bo = bo_mode ? (((d * rbb + koff) >> 7) & 7) : 0;
                    CK(cudaMemset(o_dev, 0xff, 128 * 32 * 4));
                    shift_kernel<<<1, 128, 34 * 1024 + 8192 + 2048>>>(mode, d, koff, bo, a_dev, 32768, o_dev);
                    CK(cudaDeviceSynchronize());
                    std::vector<int32_t> h(128 * 32);
                    CK(cudaMemcpy(h.data(), o_dev, h.size() * 4, cudaMemcpyDeviceToHost));
                    int bad = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 32; ++n)
                            if (h[m * 32 + n] != (int32_t)val[(m + d) * rb + koff + n]) ++bad;
                    ++total;
                    if (!bad) ++ok;
                    else if (strlen(detail) < 200) sprintf(detail + strlen(detail), " d%d/k%d:%d", d, koff, bad);
                }
            printf("shift mode=%d base_offset=%s: %d/%d exact%s%s\n", mode, bo_mode ? "addr>>7&7" : "0", ok, total,
                   detail[0] ? " | bad:" : "", detail);
            if (mode == 0) break;
        }
    }
    return 0;
}
