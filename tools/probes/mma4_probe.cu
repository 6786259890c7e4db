// mma4_probe.cu -- how much issue-side latency between bursts of tcgen05.mma does the tensor pipe hide?
// One thread issues bursts of G MMAs (M128 N64 K32, 48 cycles each when back to back); between bursts it
// runs a dependent chain of L integer multiply-adds whose result feeds the next burst's descriptors (as a
// kernel's per-stage barrier wait + descriptor arithmetic does).  cycles per burst vs 48 * G shows how deep
// the MMA queue is, i.e. how many cycles of issue-side work a stage boundary may cost before the pipe drains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma4_probe mma4_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../f8net_b200/csrc/umma_common.cuh"

using namespace f8u;

template <int N, int G>
__global__ void __launch_bounds__(64, 1) k(int L, int iters, int mul, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t a_base = f8::smem_u32(smem), b_base = a_base + 48 * 1024;
    for (int i = threadIdx.x; i < 112 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) { mbar_init(f8::smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(false, N);
        const uint64_t da0 = smem_desc(a_base, 16, 512) | (4ull << 61);
        const uint64_t db0 = smem_desc(b_base, N * 16, 128);
        uint32_t x = (uint32_t)mul;              // dependent chain state; stays 0 or tiny so that addresses stay valid
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int l = 0; l < L; ++l) x = x * (uint32_t)mul + (uint32_t)l;        // issue-side latency (mul == 0 at run time)
            const uint64_t da = da0 + (uint64_t)(x & 3u), db = db0 + (uint64_t)(x & 1u);
#pragma unroll
            for (int j = 0; j < G; ++j)
                umma_i8(tmem + (uint32_t)((j & 3) * N), da + (uint64_t)((j % 6) * 4), db + (uint64_t)((j & 1) * 2 * N), idesc, 1);
        }
        umma_commit(f8::smem_u32(&bar));
        mbar_wait(f8::smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
        if (x == 0x12345) out[200] = x;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
static long long *dout;

template <int N, int G>
void run() {
    const int iters = 500;
    CK(cudaFuncSetAttribute(k<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    for (int L : {0, 8, 16, 32, 64, 96, 128, 192, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
            k<N, G><<<148, 64, 130 * 1024>>>(L, iters, 0, dout);
            CK(cudaDeviceSynchronize());
        }
        std::vector<long long> h(148);
        CK(cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("N=%3d burst=%2d chain L=%3d: %7.1f cycles/burst = %5.1f cycles/MMA (back-to-back %d)\n", N, G, L, (double)mx / iters,
               (double)mx / iters / G, G * (N / 2 > (128 + N) / 4 ? N / 2 : (128 + N) / 4));
    }
}

int main() {
    CK(cudaMalloc(&dout, 256 * sizeof(long long)));
    run<64, 8>();
    run<64, 24>();
    run<128, 12>();
    run<32, 24>();
    return 0;
}
