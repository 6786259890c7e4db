// dw_conv.cu -- depthwise 3x3 convolution (stride 1 or 2, pad 1) on CUDA cores with the
// fused F8Net epilogue.
//
// Replaces: int nn.Conv2d.__call__ with groups == in_channels built by int_conv()
// (/root/reference/models/fix_quant_ops.py:680-714), which ATen runs as C separate
// single-channel convolutions, plus the ReLU and the consumer-side
// int_op_only_fix_quant that follow it (/root/reference/models/fix_mobilenet_v1.py:27-38,
// /root/reference/models/fix_mobilenet_v2.py:22-33).
//
// Layout: NHWC 8-bit, channels padded to a multiple of 32.  One thread owns 4 channels
// (one 32-bit word) of TW consecutive output columns of one output row, so the three
// input rows it touches are read once per thread with word loads that are contiguous
// across the threads of a warp (channel-fastest).  Products use dp4a on tap-transposed
// bytes: for each channel the 9 taps are packed 4+4+1 into three dp4a operands, which
// cuts the integer instruction count ~2.4x against one IMAD per tap per channel.
// HBM-bound by design: each input byte is read once from DRAM (neighbouring rows hit L1/L2).
#include "f8_common.cuh"

namespace {

constexpr int TW = 4;          // output columns per thread
constexpr int THREADS = 256;

struct DwGeom {
    const uint8_t *in;
    const uint32_t *w;       // [12][cpad/4] : dp4a operand j of channel c4 (see pack)
    int n, hin, win, hout, wout, cpad, stride;
    int wtiles;              // ceil(wout / TW)
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

template <bool A_SIGNED, int STRIDE>
__global__ void __launch_bounds__(THREADS)
dw3x3_kernel(const DwGeom g, const f8::Epilogue ep) {
    const int c4n = g.cpad >> 2;                       // channel words per pixel
    const long long total = (long long)g.n * g.hout * g.wtiles * c4n;
    for (long long idx = blockIdx.x * (long long)THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * THREADS) {
        const int c4 = (int)(idx % c4n);
        long long t = idx / c4n;
        const int wt = (int)(t % g.wtiles);
        t /= g.wtiles;
        const int p = (int)(t % g.hout);
        const int img = (int)(t / g.hout);
        const int q0 = wt * TW;

        // weights of this channel word: 3 dp4a operands per channel, 4 channels
        uint32_t wv[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) wv[j] = __ldg(g.w + (size_t)j * c4n + c4);

        constexpr int IW = (TW - 1) * STRIDE + 3;      // input columns needed
        uint32_t x[3][IW];
        const int ih0 = p * STRIDE - 1, iw0 = q0 * STRIDE - 1;
        const uint32_t *inw = reinterpret_cast<const uint32_t *>(g.in) +
                              (size_t)img * g.hin * g.win * c4n + c4;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int ih = ih0 + r;
            const bool rok = (unsigned)ih < (unsigned)g.hin;
#pragma unroll
            for (int j = 0; j < IW; ++j) {
                const int iw = iw0 + j;
                const bool ok = rok && (unsigned)iw < (unsigned)g.win;
                x[r][j] = ok ? __ldg(inw + ((size_t)ih * g.win + iw) * c4n) : 0u;
            }
        }
        const int4 bias = *reinterpret_cast<const int4 *>(ep.bias + c4 * 4);
        const int bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
        for (int j = 0; j < TW; ++j) {
            const int q = q0 + j;
            if (q >= g.wout) break;
            const int jj = j * STRIDE;
            // taps 0-3, 4-7 transposed to per-channel dp4a operands; tap 8 separately
            uint32_t ta[4], tb[4], tc[4];
            transpose4(x[0][jj], x[0][jj + 1], x[0][jj + 2], x[1][jj], ta);
            transpose4(x[1][jj + 1], x[1][jj + 2], x[2][jj], x[2][jj + 1], tb);
            transpose4(x[2][jj + 2], 0u, 0u, 0u, tc);
            int32_t v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int32_t a = bb[c];
                a = dot4<A_SIGNED>(ta[c], wv[c * 3 + 0], a);
                a = dot4<A_SIGNED>(tb[c], wv[c * 3 + 1], a);
                a = dot4<A_SIGNED>(tc[c], wv[c * 3 + 2], a);
                v[c] = a;
            }
            const size_t m = ((size_t)img * g.hout + p) * g.wout + q;
            const size_t o = m * ep.cout_pad + c4 * 4;
            const bool has_carry = ep.carry_in != nullptr;
            int4 cin = make_int4(0, 0, 0, 0);
            if (has_carry) cin = *reinterpret_cast<const int4 *>(ep.carry_in + f8::carry_off(m, c4 * 4, ep.cout_pad));
            v[0] = f8::residual_relu(v[0], has_carry, cin.x, ep.carry_shift, ep.relu);
            v[1] = f8::residual_relu(v[1], has_carry, cin.y, ep.carry_shift, ep.relu);
            v[2] = f8::residual_relu(v[2], has_carry, cin.z, ep.carry_shift, ep.relu);
            v[3] = f8::residual_relu(v[3], has_carry, cin.w, ep.carry_shift, ep.relu);
            if (ep.carry_out)
                *reinterpret_cast<int4 *>(ep.carry_out + f8::carry_off(m, c4 * 4, ep.cout_pad)) =
                    make_int4(v[0], v[1], v[2], v[3]);
            if (ep.out0) {
                uint32_t pk = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    pk |= ((uint32_t)f8::requant(v[c], ep.shift0, ep.signed0) & 0xffu) << (8 * c);
                *reinterpret_cast<uint32_t *>(ep.out0 + o) = pk;
            }
            if (ep.out1) {
                uint32_t pk = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    pk |= ((uint32_t)f8::requant(v[c], ep.shift1, ep.signed1) & 0xffu) << (8 * c);
                *reinterpret_cast<uint32_t *>(ep.out1 + o) = pk;
            }
        }
    }
}

}  // namespace

namespace f8host {

int launch_dw3x3(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 3 || a.kw != 3 || a.pad != 1 || (a.stride != 1 && a.stride != 2) ||
        a.cin_pad != a.cout_pad || a.cin_pad % 4 != 0) {
        set_error("conv_dw3x3: only 3x3 pad 1 stride 1|2 with cin_pad == cout_pad (mult of 4)");
        return F8_ERR_UNSUPPORTED;
    }
    DwGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.w = static_cast<const uint32_t *>(a.wpack);
    g.n = a.n; g.hin = a.hin; g.win = a.win; g.hout = a.hout; g.wout = a.wout;
    g.cpad = a.cin_pad; g.stride = a.stride;
    g.wtiles = (a.wout + TW - 1) / TW;
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
    const long long total = (long long)g.n * g.hout * g.wtiles * (g.cpad >> 2);
    long long blocks = (total + THREADS - 1) / THREADS;
    const long long cap = 148LL * 8 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool sgn = a.in_signed != 0;
    if (a.stride == 1) {
        if (sgn) dw3x3_kernel<true, 1><<<(unsigned)blocks, THREADS, 0, s>>>(g, ep);
        else dw3x3_kernel<false, 1><<<(unsigned)blocks, THREADS, 0, s>>>(g, ep);
    } else {
        if (sgn) dw3x3_kernel<true, 2><<<(unsigned)blocks, THREADS, 0, s>>>(g, ep);
        else dw3x3_kernel<false, 2><<<(unsigned)blocks, THREADS, 0, s>>>(g, ep);
    }
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

}  // namespace f8host
