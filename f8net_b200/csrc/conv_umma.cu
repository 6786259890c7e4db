// conv_umma.cu -- dense int8 convolution / linear as an implicit GEMM on the Blackwell
// tensor core: tcgen05.mma.kind::i8 (u8|s8 x s8 -> s32) with the accumulator in TMEM, the
// weight tiles brought in by the bulk-copy engine (cp.async.bulk + mbarrier complete_tx) and
// the whole F8Net inter-layer epilogue fused after tcgen05.ld.
//
// Replaces, per launch: int nn.Conv2d.__call__ / nn.Linear.__call__ built by
// int_conv()/int_fc() (/root/reference/models/fix_quant_ops.py:680-714, :1165-1195) and the
// tensor-op chain around it in IntBlock.forward (/root/reference/models/fix_resnet.py:28-77):
// bias, residual shift-add-clamp, ReLU, consumer-side int_op_only_fix_quant, .float().
//
// GEMM view: M = n*hout*wout output pixels, N = cout_pad, K = kh*kw*cin_pad bytes.
// CTA tile 128 x BN (BN = 64 | 128 | 256 TMEM columns), K step 64 bytes per pipeline stage
// (two K=32 MMAs).  Warp roles:
//   warps 0-3  producers, then epilogue.  Thread t owns output pixel m0+t: it gathers the
//              pixel's K bytes from the NHWC activation with 16-byte cp.async (zero fill for
//              the padding halo) into the canonical K-major no-swizzle operand layout
//              [K/16][128 rows][16 B] (core matrix = 8 rows x 16 B contiguous, SBO = 128 B,
//              LBO = 2048 B), fences the generic->async proxy and arrives on the stage's
//              "full" mbarrier.  Thread 0 also posts the stage's weight bytes: 4 bulk copies
//              of BN*16 contiguous bytes from the chunk-major weight image.
//   warp 4     lane 0 issues the MMAs (tcgen05.mma is a single-thread instruction) and
//              commits each stage to its "empty" mbarrier; the last commit signals the
//              epilogue.  Warp 4 also owns the TMEM allocation.
// After the K loop each producer thread reads its own accumulator row (TMEM lane = tile row)
// 16 columns at a time and runs the integer epilogue of f8_common.cuh.
// Integer accumulation is associative mod 2^32: tiling and MMA order cannot change results.
#include "f8_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int PRODUCERS = 128;
constexpr int THREADS = 160;
constexpr int A_STAGE = BM * BK;          // 8192 B: [4 chunks][128 rows][16 B]
constexpr int A_CHUNK = BM * 16;          // 2048 B between K chunks (LBO of A)

struct UGeom {
    const uint8_t *in;
    const uint8_t *wpack;  // [K_pad/16][wrows][16]
    int wrows;
    int M;
    int hin, win, cin_pad;
    int hout, wout;
    int kh, kw, stride, pad;
    int ktiles;            // K_pad / 64
    int row_bytes;         // small-C mode
    int shift_px;          // small-C mode
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, bool pred) {
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 operands, int32 accumulate, no saturation
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (layout_type 0), version 1:
// start address [0,14), leading byte offset [16,30), stride byte offset [32,46), all >> 4
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor for kind::i8: c_format S32 (2) at [4,6), a_format at [7,10)
// (0 = u8, 1 = s8), b_format s8 at [10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc(bool a_signed, int n) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

template <int BN, bool A_SIGNED, bool SMALL_C>
__global__ void __launch_bounds__(THREADS)
conv_umma_kernel(const UGeom g, const f8::Epilogue ep) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int B_STAGE = BN * BK;
    constexpr int B_CHUNK = BN * 16;
    constexpr int STAGE = A_STAGE + B_STAGE;
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t bar_base = smem_base + STAGES * STAGE;       // full[S] empty[S] accum
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + STAGES * STAGE + (2 * STAGES + 1) * 8);
    auto full_bar = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (uint32_t)(STAGES + s) * 8; };
    const uint32_t accum_bar = bar_base + 2 * STAGES * 8;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), PRODUCERS);
                mbar_init(empty_bar(s), 1);
            }
            mbar_init(accum_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== producer: A gather + B bulk copies ===========
        const int row = tid;
        const int m = m0 + row;
        const bool valid = m < g.M;
        const int HW = g.hout * g.wout;
        const int mm = valid ? m : 0;
        const int img = mm / HW;
        const int rem = mm - img * HW;
        const int p = rem / g.wout, q = rem - p * g.wout;
        const int ih0 = p * g.stride - g.pad;
        const int iw0 = q * g.stride - g.pad - (SMALL_C ? g.shift_px : 0);
        const uint8_t *base = g.in + (size_t)img * g.hin * g.win * g.cin_pad;
        int k_r = 0, k_s = 0, k_c = 0;

        for (int kt = 0; kt < g.ktiles; ++kt) {
            const int slot = kt % STAGES;
            if (kt >= STAGES) mbar_wait(empty_bar(slot), ((kt / STAGES) - 1) & 1);
            const uint32_t sa = smem_base + slot * STAGE;
            if (tid == 0) {
                const uint32_t sb = sa + A_STAGE;
                mbar_expect_tx(full_bar(slot), B_STAGE);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    bulk_g2s(sb + j * B_CHUNK,
                             g.wpack + ((size_t)(kt * 4 + j) * g.wrows + n0) * 16, B_CHUNK,
                             full_bar(slot));
            }
            if constexpr (SMALL_C) {
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int ih = ih0 + k_r;
                    const int iw = iw0 + (k_c >> 2);
                    const bool ok = valid && k_r < g.kh && (unsigned)ih < (unsigned)g.hin &&
                                    (unsigned)iw < (unsigned)g.win;
                    const uint8_t *src = ok ? base + ((size_t)ih * g.win + iw) * 4 : g.in;
                    cp_async8(sa + (c8 >> 1) * A_CHUNK + row * 16 + (c8 & 1) * 8, src, ok);
                    k_c += 8;
                    if (k_c >= g.row_bytes) { k_c = 0; ++k_r; }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ih = ih0 + k_r;
                    const int iw = iw0 + k_s;
                    const bool ok = valid && k_r < g.kh && (unsigned)ih < (unsigned)g.hin &&
                                    (unsigned)iw < (unsigned)g.win;
                    const uint8_t *src =
                        ok ? base + ((size_t)ih * g.win + iw) * g.cin_pad + k_c : g.in;
                    cp_async16(sa + j * A_CHUNK + row * 16, src, ok);
                    k_c += 16;
                    if (k_c >= g.cin_pad) {
                        k_c = 0;
                        if (++k_s == g.kw) { k_s = 0; ++k_r; }
                    }
                }
            }
            cp_async_commit();
            if (kt >= STAGES - 1) {
                cp_async_wait<STAGES - 1>();      // stage kt-(STAGES-1) has landed
                fence_proxy_async();              // generic-proxy writes -> visible to the MMA
                mbar_arrive(full_bar((kt - (STAGES - 1)) % STAGES));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        {
            int first = g.ktiles - (STAGES - 1);
            if (first < 0) first = 0;
            for (int i = first; i < g.ktiles; ++i) mbar_arrive(full_bar(i % STAGES));
        }

        // =========================== epilogue ====================================
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const bool has_carry = ep.carry_in != nullptr;
        int ncols = ep.cout_pad - n0;
        if (ncols > BN) ncols = BN;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            int32_t v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid) {
                const int gc = n0 + c0;
                const size_t o = (size_t)m * ep.cout_pad + gc;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const int4 b = __ldg(reinterpret_cast<const int4 *>(ep.bias + gc + i));
                    v[i + 0] = (int32_t)((uint32_t)v[i + 0] + (uint32_t)b.x);
                    v[i + 1] = (int32_t)((uint32_t)v[i + 1] + (uint32_t)b.y);
                    v[i + 2] = (int32_t)((uint32_t)v[i + 2] + (uint32_t)b.z);
                    v[i + 3] = (int32_t)((uint32_t)v[i + 3] + (uint32_t)b.w);
                    int4 c = make_int4(0, 0, 0, 0);
                    if (has_carry) c = *reinterpret_cast<const int4 *>(ep.carry_in + o + i);
                    v[i + 0] = f8::residual_relu(v[i + 0], has_carry, c.x, ep.carry_shift, ep.relu);
                    v[i + 1] = f8::residual_relu(v[i + 1], has_carry, c.y, ep.carry_shift, ep.relu);
                    v[i + 2] = f8::residual_relu(v[i + 2], has_carry, c.z, ep.carry_shift, ep.relu);
                    v[i + 3] = f8::residual_relu(v[i + 3], has_carry, c.w, ep.carry_shift, ep.relu);
                    if (ep.carry_out)
                        *reinterpret_cast<int4 *>(ep.carry_out + o + i) =
                            make_int4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                if (ep.out0) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        w[i] = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            w[i] |= ((uint32_t)f8::requant(v[i * 4 + b], ep.shift0, ep.signed0) & 0xffu)
                                    << (8 * b);
                    }
                    *reinterpret_cast<uint4 *>(ep.out0 + o) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (ep.out1) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        w[i] = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            w[i] |= ((uint32_t)f8::requant(v[i * 4 + b], ep.shift1, ep.signed1) & 0xffu)
                                    << (8 * b);
                    }
                    *reinterpret_cast<uint4 *>(ep.out1 + o) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (ep.out_f32) {
                    float *f = ep.out_f32 + (size_t)m * ep.out_f32_ld + gc;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (gc + i < ep.cout) f[i] = (float)v[i];
                }
            }
        }
    } else if (lane == 0) {
        // =========================== MMA issuer ==================================
        constexpr uint32_t idesc = instr_desc(A_SIGNED, BN);
        for (int kt = 0; kt < g.ktiles; ++kt) {
            const int slot = kt % STAGES;
            mbar_wait(full_bar(slot), (kt / STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = smem_base + slot * STAGE;
            const uint32_t sb = sa + A_STAGE;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint64_t ad = smem_desc(sa + i * 2 * A_CHUNK, A_CHUNK, 128);
                const uint64_t bd = smem_desc(sb + i * 2 * B_CHUNK, B_CHUNK, 128);
                umma_i8(tmem_base, ad, bd, idesc, (uint32_t)((kt | i) != 0));
            }
            umma_commit(empty_bar(slot));     // frees the stage when these MMAs have read it
        }
        umma_commit(accum_bar);               // accumulator complete -> epilogue
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN, bool A_SIGNED, bool SMALL_C>
int launch_t(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s) {
    constexpr int smem_bytes = STAGES * (A_STAGE + BN * BK) + (2 * STAGES + 1) * 8 + 16;
    auto kern = conv_umma_kernel<BN, A_SIGNED, SMALL_C>;
    static bool attr_done = false;
    if (!attr_done) {
        F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_done = true;
    }
    dim3 grid((g.M + BM - 1) / BM, (ep.cout_pad + BN - 1) / BN);
    kern<<<grid, THREADS, smem_bytes, s>>>(g, ep);
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

template <int BN>
int launch_bn(const UGeom &g, const f8::Epilogue &ep, bool sgn, bool small_c, cudaStream_t s) {
    if (small_c) return sgn ? launch_t<BN, true, true>(g, ep, s) : launch_t<BN, false, true>(g, ep, s);
    return sgn ? launch_t<BN, true, false>(g, ep, s) : launch_t<BN, false, false>(g, ep, s);
}

}  // namespace

namespace f8host {

int launch_conv_umma(const f8_conv_args &a, cudaStream_t s) {
    if ((a.cin_pad != 4 && a.cin_pad % 16 != 0) || a.cout_pad % 16 != 0) return F8_ERR_UNSUPPORTED;
    const DensePack pk = dense_pack_geometry(a.cin_pad, a.cout_pad, a.kh, a.kw);
    if (pk.mode == 1 && ((a.stride & 1) || ((a.pad + pk.shift_px) & 1) || (a.win & 1) ||
                         (pk.row_bytes & 7)))
        return F8_ERR_UNSUPPORTED;
    const long long M = (long long)a.n * a.hout * a.wout;
    if (M <= 0 || M > 0x7fffffffLL) {
        set_error("conv_dense: pixel count %lld out of range", M);
        return F8_ERR_ARG;
    }
    UGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.wpack = static_cast<const uint8_t *>(a.wpack);
    g.wrows = pk.rows;
    g.M = (int)M;
    g.hin = a.hin; g.win = a.win; g.cin_pad = a.cin_pad;
    g.hout = a.hout; g.wout = a.wout;
    g.kh = a.kh; g.kw = a.kw; g.stride = a.stride; g.pad = a.pad;
    g.ktiles = pk.K_pad / BK;
    g.row_bytes = pk.row_bytes;
    g.shift_px = pk.shift_px;
    f8::Epilogue ep{};
    ep.bias = a.bias;
    ep.carry_in = a.carry_in;
    ep.carry_out = a.carry_out;
    ep.out0 = static_cast<uint8_t *>(a.out[0]);
    ep.out1 = static_cast<uint8_t *>(a.out[1]);
    ep.out_f32 = a.out_f32;
    ep.out_f32_ld = a.out_f32_ld;
    ep.carry_shift = a.carry_shift;
    ep.relu = a.relu;
    ep.shift0 = a.out_shift[0]; ep.signed0 = a.out_signed[0];
    ep.shift1 = a.out_shift[1]; ep.signed1 = a.out_signed[1];
    ep.cout = a.cout;
    ep.cout_pad = a.cout_pad;
    const bool sgn = a.in_signed != 0;
    const bool small_c = pk.mode == 1;
    if (a.cout_pad <= 64) return launch_bn<64>(g, ep, sgn, small_c, s);
    if (a.cout_pad <= 128) return launch_bn<128>(g, ep, sgn, small_c, s);
    return launch_bn<256>(g, ep, sgn, small_c, s);
}

}  // namespace f8host
