// tmem_probe.cu -- TMEM read bandwidth: W warps loop tcgen05.ld.32x32b.x16 (+ optional x32/x64) over
// the 512 columns; reports bytes per cycle per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../f8net_b200/csrc/umma_common.cuh"
using namespace f8u;
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, int inflight, long long *out, int *sink) {
    __shared__ uint32_t tslot;
    if (threadIdx.x < 32) tmem_alloc(f8::smem_u32(&tslot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    int acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (X == 16) {
            int32_t a[16], b[16], c[16];
            const uint32_t col = (uint32_t)(((it * 3 + (warp >> 2)) * 48) & 255);
            tmem_ld16(base + col, a);
            if (inflight > 1) tmem_ld16(base + col + 16, b);
            if (inflight > 2) tmem_ld16(base + col + 32, c);
            tmem_ld_wait();
            acc += a[0] + a[15];
            if (inflight > 1) acc += b[3];
            if (inflight > 2) acc += c[5];
        } else {
            int32_t a[32];
            const uint32_t col = (uint32_t)(((it + (warp >> 2)) * 32) & 255);
            tmem_ld32(base + col, a);
            tmem_ld_wait();
            acc += a[0] + a[31];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678) *sink = acc;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
int main() {
    long long *d; int *sink; cudaMalloc(&d, 8 * 148); cudaMalloc(&sink, 4);
    const int iters = 4000;
    for (int warps : {4, 8, 16}) for (int inflight : {1, 3}) {
        k<16><<<148, warps * 32>>>(iters, inflight, d, sink); cudaDeviceSynchronize();
        k<16><<<148, warps * 32>>>(iters, inflight, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * inflight * 16 * 32 * 4 * warps;
        printf("x16 warps=%2d inflight=%d: %.1f B/clk/SM (%.0f cycles/iter) %s\n", warps, inflight, bytes / h, (double)h / iters, cudaGetErrorString(e));
    }
    for (int warps : {4, 8, 16}) {
        k<32><<<148, warps * 32>>>(iters, 1, d, sink); cudaDeviceSynchronize();
        k<32><<<148, warps * 32>>>(iters, 1, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * 32 * 32 * 4 * warps;
        printf("x32 warps=%2d: %.1f B/clk/SM (%.0f cycles/iter) %s\n", warps, bytes / h, (double)h / iters, cudaGetErrorString(e));
    }
    return 0;
}
