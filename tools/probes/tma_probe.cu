// tma_probe.cu -- which tensor-map shapes does the TMA accept for the head image view?
// usage: tma_probe <variant>; prints OK / the CUDA error.  One variant per process (an illegal
// instruction poisons the context).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../f8net_b200/csrc/umma_common.cuh"
using namespace f8u;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Dummy { const void *a; int b, c; long long *d; };
struct Dummy2 { const void *p[6]; int q[12]; };
__global__ void k(const Dummy g, const Dummy2 e, const __grid_constant__ CUtensorMap tmap, uint32_t doff, int c0, int c1, int c2, int c3, uint32_t bytes, uint32_t *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t b = f8::smem_u32(&bar);
    if (threadIdx.x == 0) { mbar_init(b, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(b, bytes);
        mbar_arrive(b);
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(f8::smem_u32(smem) + doff), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b) : "memory");
    }
    mbar_wait(b, 0);
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((uint32_t *)(smem + doff))[i];
}
int main(int argc, char **argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int N = 2;
    std::vector<uint32_t> img(N * 224 * 224);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint32_t)(i + 1);
    uint32_t *d, *o;
    cudaMalloc(&d, img.size() * 4); cudaMalloc(&o, 65536);
    cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cuuint64_t gd[4] = {224, 4, 56, N};
    cuuint64_t gs[3] = {896, 3584, 224 * 896};
    cuuint32_t bx[4] = {232, 1, 6, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    int c0 = -4, c1 = 1, c2 = 3, c3 = 1;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (v == 1) bx[0] = 224, c0 = 0;
    if (v == 2) bx[0] = 224;             // negative start, box == dim
    if (v == 3) c0 = 0;                  // oversize box, start 0
    if (v == 4) bx[2] = 1;
    if (v == 5) l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (v == 6) bx[0] = 128;
    if (v == 7) { bx[0] = 232; c2 = -1; }
    CUtensorMap tm;
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: encode rc=%d box %u %u %u %u start %d %d %d %d\n", v, (int)r, bx[0], bx[1], bx[2], bx[3], c0, c1, c2, c3);
    if (r) return 0;
    const uint32_t bytes = bx[0] * bx[1] * bx[2] * bx[3] * 4;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    Dummy dg{}; Dummy2 de{}; uint32_t doff = 0; if (v == 8) doff = 128; if (v == 9) doff = 5632; if (v == 10) doff = 22528 + 3 * 5632;
    k<<<1, 128, 65536>>>(dg, de, tm, doff, c0, c1, c2, c3, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: run: %s\n", v, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint32_t> h(bytes / 4);
        cudaMemcpy(h.data(), o, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (uint32_t j = 0; j < bx[2]; ++j)
            for (uint32_t x = 0; x < bx[0]; ++x) {
                const int xx = c0 + (int)x, J = c2 + (int)j;
                uint32_t want = 0;
                if (xx >= 0 && xx < 224 && J >= 0 && J < 56) want = img[((size_t)c3 * 224 + (4 * J + c1)) * 224 + xx];
                if (h[j * bx[0] + x] != want) ++bad;
            }
        printf("variant %d: %d mismatches of %u\n", v, bad, bx[0] * bx[2]);
    }
    return 0;
}
