// dw_conv.cu -- depthwise 3x3 convolution (stride 1 or 2, pad 1) on CUDA cores with the
// fused F8Net epilogue.
//
// Replaces: int nn.Conv2d.__call__ with groups == in_channels built by int_conv()
// (/root/reference/models/fix_quant_ops.py:680-714), which ATen runs as C separate
// single-channel convolutions, plus the ReLU and the consumer-side
// int_op_only_fix_quant that follow it (/root/reference/models/fix_mobilenet_v1.py:27-38,
// /root/reference/models/fix_mobilenet_v2.py:22-33).
//
// Layout: NHWC 8-bit, channels padded to a multiple of 16.  One thread computes 16 channels
// (one 16-byte vector) of one output pixel; threads are laid over the flattened (pixel, channel
// group) space so every tap load of a warp is one contiguous 512-byte run and the 9-tap overlap
// between neighbouring outputs is served by L1.  Products use dp4a on tap-transposed bytes: for
// each channel the 9 taps are packed 4+4+1 into three dp4a operands, which cuts the integer
// instruction count ~2.4x against one IMAD per tap per channel.  HBM-bound by design: each input
// byte is read once from DRAM.
#include <cstdlib>

#include "f8_common.cuh"

namespace {

constexpr int THREADS = 256;

struct DwGeom {
    const uint8_t *in;
    const uint32_t *w;       // [12][cpad/4] : dp4a operand j of channel c4 (see pack)
    int n, hin, win, hout, wout, cpad, stride;
};

// transpose a 4x4 byte matrix held in 4 registers (rows = taps, cols = channels)
__device__ __forceinline__ void transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                           uint32_t (&o)[4]) {
    const uint32_t ab_lo = __byte_perm(a, b, 0x5140);  // a0 b0 a1 b1
    const uint32_t ab_hi = __byte_perm(a, b, 0x7362);  // a2 b2 a3 b3
    const uint32_t cd_lo = __byte_perm(c, d, 0x5140);
    const uint32_t cd_hi = __byte_perm(c, d, 0x7362);
    o[0] = __byte_perm(ab_lo, cd_lo, 0x5410);          // a0 b0 c0 d0
    o[1] = __byte_perm(ab_lo, cd_lo, 0x7632);          // a1 b1 c1 d1
    o[2] = __byte_perm(ab_hi, cd_hi, 0x5410);
    o[3] = __byte_perm(ab_hi, cd_hi, 0x7632);
}

template <bool A_SIGNED>
__device__ __forceinline__ int32_t dot4(uint32_t x, uint32_t w, int32_t acc) {
    // x: activations (u8 or s8), w: weights (s8)
    if constexpr (A_SIGNED) {
        asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(w));
    } else {
        asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(w));
    }
    return acc;
}

// One thread = one output pixel x 16 channels (one 16-byte vector).  Threads are laid over the
// flattened (pixel, 16-channel group) space, so a warp's load of one tap is a contiguous
// 512-byte run and the overlap between neighbouring outputs' taps is served by L1.  The grid
// stride is a multiple of the number of channel groups: a thread keeps its channel group for
// the whole launch and holds its 48 weight words and 16 biases in registers.
template <bool A_SIGNED, int STRIDE>
__global__ void __launch_bounds__(THREADS, 2)
dw3x3_kernel(const DwGeom g, const f8::Epilogue ep, const long long total, const long long stride_items) {
    const int c16n = g.cpad >> 4;                      // 16-channel groups per pixel
    long long idx = blockIdx.x * (long long)THREADS + threadIdx.x;
    if (idx >= total) return;
    const int c16 = (int)((uint32_t)idx % (uint32_t)c16n);
    const int c4n = g.cpad >> 2;
    // weights of this thread's 16 channels: word j of channel quad c4 at w[j * c4n + c4]
    uint32_t wv[4][12];
#pragma unroll
    for (int cq = 0; cq < 4; ++cq)
#pragma unroll
        for (int j = 0; j < 12; ++j) wv[cq][j] = __ldg(g.w + (size_t)j * c4n + c16 * 4 + cq);
    int32_t bias[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const int4 b = __ldg(reinterpret_cast<const int4 *>(ep.bias + c16 * 16 + i));
        bias[i] = b.x; bias[i + 1] = b.y; bias[i + 2] = b.z; bias[i + 3] = b.w;
    }
    const bool has_carry = ep.carry_in != nullptr;
    const bool plain = f8::epilogue_is_plain_u8(ep);
    const f8::EpiConst kc = f8::epi_const(ep, has_carry);
    // programmatic dependent launch: weights and bias (plan constants) are in registers before
    // the previous layer has finished; its output is only touched below
    f8::pdl_trigger();
    f8::pdl_wait();
    if (plain) {
#pragma unroll
        for (int i = 0; i < 16; ++i) bias[i] = (int32_t)((uint32_t)bias[i] + (1u << (ep.shift0 - 1)));
    }
    const uint4 *in16 = reinterpret_cast<const uint4 *>(g.in);
    for (; idx < total; idx += stride_items) {
        // 32-bit index arithmetic (the launcher guarantees total < 2^31): 64-bit divisions cost
        // ~100 instructions each
        const uint32_t pix = (uint32_t)idx / (uint32_t)c16n;      // output pixel (image-major)
        const uint32_t t = pix / (uint32_t)g.wout;
        const int q = (int)(pix - t * (uint32_t)g.wout);
        const int img = (int)(t / (uint32_t)g.hout);
        const int p = (int)(t - (uint32_t)img * (uint32_t)g.hout);
        const int ih0 = p * STRIDE - 1, iw0 = q * STRIDE - 1;
        uint4 x[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int ih = ih0 + r;
            const bool rok = (unsigned)ih < (unsigned)g.hin;
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) {
                const int iw = iw0 + s2;
                const bool ok = rok && (unsigned)iw < (unsigned)g.win;
                x[r * 3 + s2] = ok ? __ldg(in16 + (((size_t)img * g.hin + ih) * g.win + iw) * c16n + c16)
                                   : make_uint4(0, 0, 0, 0);
            }
        }
        int32_t v[16];
#pragma unroll
        for (int cq = 0; cq < 4; ++cq) {
            // channel quad cq: word cq of each tap; taps 0-3, 4-7 transposed to per-channel operands
            auto w_of = [&](const uint4 &u) { return cq == 0 ? u.x : cq == 1 ? u.y : cq == 2 ? u.z : u.w; };
            uint32_t ta[4], tb[4], tc[4];
            transpose4(w_of(x[0]), w_of(x[1]), w_of(x[2]), w_of(x[3]), ta);
            transpose4(w_of(x[4]), w_of(x[5]), w_of(x[6]), w_of(x[7]), tb);
            transpose4(w_of(x[8]), 0u, 0u, 0u, tc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int32_t a = bias[cq * 4 + c];
                a = dot4<A_SIGNED>(ta[c], wv[cq][c * 3 + 0], a);
                a = dot4<A_SIGNED>(tb[c], wv[cq][c * 3 + 1], a);
                a = dot4<A_SIGNED>(tc[c], wv[cq][c * 3 + 2], a);
                v[cq * 4 + c] = a;
            }
        }
        const size_t o = (size_t)pix * ep.cout_pad + c16 * 16;
        if (plain) {
            // bias already holds bias + 2^(n-1); unsigned clamp absorbs the ReLU (f8_common.cuh)
            const int n = ep.shift0;
            const uint32_t mask = (1u << n) - 1u;
            uint32_t w4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int32_t r4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t tt = (uint32_t)v[4 * k + e];
                    r4[e] = (int32_t)tt >> n;
                    if ((tt & mask) == 0u) r4[e] &= ~1;
                }
                uint32_t hi;
                asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r4[3]), "r"(r4[2]), "r"(0));
                asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w4[k]) : "r"(r4[1]), "r"(r4[0]), "r"(hi));
            }
            *reinterpret_cast<uint4 *>(ep.out0 + o) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        } else {
            int4 c[4];
            if (has_carry) {
                const int32_t *src = ep.carry_in + f8::carry_off((size_t)pix, c16 * 16, ep.cout_pad);
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = __ldg(reinterpret_cast<const int4 *>(src + k * 512));
            }
            // v already includes the bias: run the shared math with a zero bias vector
            const int4 z = make_int4(0, 0, 0, 0);
            int32_t zero16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            (void)z;
            f8::epilogue16_math(v, zero16, kc, c, has_carry);
            if (ep.carry_out) {
                int32_t *dst = ep.carry_out + f8::carry_off((size_t)pix, c16 * 16, ep.cout_pad);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<int4 *>(dst + k * 512) =
                        make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
            if (ep.out0) *reinterpret_cast<uint4 *>(ep.out0 + o) = f8::requant_pack16(v, ep.shift0, ep.signed0);
            if (ep.out1) *reinterpret_cast<uint4 *>(ep.out1 + o) = f8::requant_pack16(v, ep.shift1, ep.signed1);
        }
    }
}

}  // namespace

namespace f8host {

int launch_dw3x3(const f8_conv_args &a, cudaStream_t s) {
#ifdef F8_WITH_UMMA
    // the tensor-core kernel (diagonal 64 x 64 weight blocks over the TMA-staged patch); odd input
    // sizes at stride 2 stay on the CUDA cores
    static const bool cuda_core_only = getenv("F8_DW_CUDA_CORE") != nullptr;
    if (!cuda_core_only) {
        const int rc = launch_conv3x3_dw(a, s);
        if (rc != F8_ERR_UNSUPPORTED) return rc;
    }
#endif
    if (a.kh != 3 || a.kw != 3 || a.pad != 1 || (a.stride != 1 && a.stride != 2) ||
        a.cin_pad != a.cout_pad || a.cin_pad % 16 != 0) {
        set_error("conv_dw3x3: only 3x3 pad 1 stride 1|2 with cin_pad == cout_pad (multiple of 16)");
        return F8_ERR_UNSUPPORTED;
    }
    DwGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.w = static_cast<const uint32_t *>(a.wpack);
    g.n = a.n; g.hin = a.hin; g.win = a.win; g.hout = a.hout; g.wout = a.wout;
    g.cpad = a.cin_pad; g.stride = a.stride;
    f8::Epilogue ep{};
    ep.bias = a.bias;
    ep.carry_in = a.carry_in;
    ep.carry_out = a.carry_out;
    ep.out0 = static_cast<uint8_t *>(a.out[0]);
    ep.out1 = static_cast<uint8_t *>(a.out[1]);
    ep.carry_shift = a.carry_shift;
    ep.relu = a.relu;
    ep.shift0 = a.out_shift[0]; ep.signed0 = a.out_signed[0];
    ep.shift1 = a.out_shift[1]; ep.signed1 = a.out_signed[1];
    ep.cout = a.cout; ep.cout_pad = a.cout_pad;
    const int c16n = g.cpad >> 4;
    const long long total = (long long)g.n * g.hout * g.wout * c16n;
    if (total >= 0x7fffffffLL) {
        set_error("conv_dw3x3: %lld items exceed the 32-bit index range", total);
        return F8_ERR_UNSUPPORTED;
    }
    // grid: ~2 blocks per SM x 8 waves, rounded so that the thread count is a multiple of the
    // number of channel groups (every thread then keeps one channel group)
    long long blocks = (total + THREADS - 1) / THREADS;
    const long long cap = 148LL * 2 * 8;
    if (blocks > cap) blocks = cap;
    long long mult = c16n;                       // blocks * 256 % c16n == 0  <=  blocks % c16n' == 0
    for (int f = 2; f <= 256; f *= 2) if (mult % 2 == 0) mult /= 2;
    blocks = (blocks + mult - 1) / mult * mult;
    const long long stride_items = blocks * THREADS;
    const bool sgn = a.in_signed != 0;
    note_kernel("dw3x3<s%d>", a.stride);
    if (a.stride == 1) {
        if (sgn) F8_CUDA(f8host::launch_pdl(dw3x3_kernel<true, 1>, (unsigned)blocks, THREADS, 0, s, g, ep, total, stride_items));
        else F8_CUDA(f8host::launch_pdl(dw3x3_kernel<false, 1>, (unsigned)blocks, THREADS, 0, s, g, ep, total, stride_items));
    } else {
        if (sgn) F8_CUDA(f8host::launch_pdl(dw3x3_kernel<true, 2>, (unsigned)blocks, THREADS, 0, s, g, ep, total, stride_items));
        else F8_CUDA(f8host::launch_pdl(dw3x3_kernel<false, 2>, (unsigned)blocks, THREADS, 0, s, g, ep, total, stride_items));
    }
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

}  // namespace f8host
