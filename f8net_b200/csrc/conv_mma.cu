// conv_mma.cu -- dense int8 convolution / linear as an implicit GEMM on the legacy
// warp-level tensor-core path (mma.sync.m16n8k32 u8|s8 x s8 -> s32), with the whole
// F8Net inter-layer epilogue fused (bias, residual shift-add-clamp, ReLU, int32 carry,
// up to two requantised 8-bit outputs, float logits).
//
// Replaces, per launch: int nn.Conv2d.__call__ / nn.Linear.__call__ built by
// int_conv()/int_fc() (/root/reference/models/fix_quant_ops.py:680-714, :1165-1195) and the
// tensor-op chain around it in IntBlock.forward (/root/reference/models/fix_resnet.py:28-77).
//
// This is the always-available backend (backend 0): it handles every shape on the path,
// including the Cin=3 heads (small-C row-window gather) and the classifier.  The tcgen05
// backend (conv_umma.cu) takes over the shapes it supports.
//
// GEMM view:  M = n*hout*wout output pixels, N = cout_pad, K = kh*kw*cin_pad.
// CTA tile 128 x BN x 64, 8 warps (4 along M x 2 along N), 4-stage cp.async pipeline,
// A gathered on the fly from the NHWC activation (zero-fill for padding), B from the
// packed weight image (16-byte K chunks, chunk-major: [K_pad/16][rows][16], the layout the
// tcgen05 kernel's bulk copies want; both backends share one image).  Integer accumulation is associative mod 2^32, so the
// tiling order cannot change the result.
#include "f8_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, bool pred) {
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3,
                                            uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

template <bool A_SIGNED>
__device__ __forceinline__ void mma_i8(int32_t (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                       uint32_t b1) {
    if constexpr (A_SIGNED) {
        asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};\n"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};\n"
            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}

struct ConvGeom {
    const uint8_t *in;
    const uint8_t *wpack;  // [K_pad/16][wrows][16]  (chunk-major, see dense_pack_geometry)
    int wrows;
    int M;                 // n*hout*wout
    int hin, win, cin_pad;
    int hout, wout;
    int kh, kw, stride, pad;
    int K_pad;
    int ktiles;            // K_pad / BK
    int row_bytes;         // small-C mode
    int shift_px;          // small-C mode
};

// physical byte offset of 16B chunk `chunk` of row `row` inside a [rows][64B] tile
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return (uint32_t)(row * BK + ((chunk ^ ((row >> 1) & 3)) << 4));
}

template <int BN, bool A_SIGNED, bool SMALL_C>
__global__ void __launch_bounds__(THREADS, (BN == 64) ? 2 : 1)
conv_mma_kernel(const ConvGeom g, const f8::Epilogue ep) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int A_STAGE = BM * BK;
    constexpr int B_STAGE = BN * BK;
    constexpr int STAGE = A_STAGE + B_STAGE;
    const uint32_t smem_base = f8::smem_u32(smem);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2;     // 4 x 2 warps
    constexpr int WN = BN / 2;                   // warp tile 32 x WN
    constexpr int NT = WN / 8;                   // n8 tiles per warp

    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = g.hout * g.wout;

    // ---------------- A-gather bookkeeping ----------------
    constexpr int A_ROWS = SMALL_C ? 4 : 2;      // rows handled per thread per stage
    const int a_chunk = SMALL_C ? (tid & 7) : (tid & 3);   // 8B or 16B chunk inside the 64B row
    const int a_row0 = SMALL_C ? (tid >> 3) : (tid >> 2);
    constexpr int A_ROW_STEP = SMALL_C ? 32 : 64;
    int a_ih0[A_ROWS], a_iw0[A_ROWS];
    const uint8_t *a_base[A_ROWS];
    bool a_valid[A_ROWS];
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
        const int m = m0 + a_row0 + i * A_ROW_STEP;
        a_valid[i] = m < g.M;
        const int mm = a_valid[i] ? m : 0;
        const int img = mm / HW;
        const int rem = mm - img * HW;
        const int p = rem / g.wout, q = rem - p * g.wout;
        a_ih0[i] = p * g.stride - g.pad;
        a_iw0[i] = q * g.stride - g.pad - (SMALL_C ? g.shift_px : 0);
        a_base[i] = g.in + (size_t)img * g.hin * g.win * g.cin_pad;
    }
    // running decomposition of this thread's K byte offset
    int k_r = 0, k_s = 0, k_c = 0;   // generic: tap row, tap col, channel; small-C: r, -, byte offset in row
    {
        const int kb = SMALL_C ? a_chunk * 8 : a_chunk * 16;
        if constexpr (SMALL_C) {
            k_r = kb / g.row_bytes;
            k_c = kb - k_r * g.row_bytes;
        } else {
            const int tap = kb / g.cin_pad;
            k_c = kb - tap * g.cin_pad;
            k_r = tap / g.kw;
            k_s = tap - k_r * g.kw;
        }
    }
    // B rows handled by this thread
    constexpr int B_ROWS = BN / 64;
    const int b_chunk = tid >> 6;
    const int b_row0 = tid & 63;

    auto load_stage = [&](int stage, int kt) {
        const uint32_t sa = smem_base + stage * STAGE;
        const uint32_t sb = sa + A_STAGE;
        // ---- A ----
#pragma unroll
        for (int i = 0; i < A_ROWS; ++i) {
            const int row = a_row0 + i * A_ROW_STEP;
            const int ih = a_ih0[i] + k_r;
            bool ok;
            const uint8_t *src;
            if constexpr (SMALL_C) {
                const int iw = a_iw0[i] + (k_c >> 2);
                ok = a_valid[i] && k_r < g.kh && (unsigned)ih < (unsigned)g.hin &&
                     (unsigned)iw < (unsigned)g.win;
                src = ok ? a_base[i] + ((size_t)ih * g.win + iw) * 4 : g.in;
                const uint32_t dst = sa + tile_off(row, a_chunk >> 1) + ((a_chunk & 1) << 3);
                cp_async8(dst, src, ok);
            } else {
                const int iw = a_iw0[i] + k_s;
                ok = a_valid[i] && k_r < g.kh && (unsigned)ih < (unsigned)g.hin &&
                     (unsigned)iw < (unsigned)g.win;
                src = ok ? a_base[i] + ((size_t)ih * g.win + iw) * g.cin_pad + k_c : g.in;
                cp_async16(sa + tile_off(row, a_chunk), src, ok);
            }
        }
        // ---- B ----
#pragma unroll
        for (int i = 0; i < B_ROWS; ++i) {
            const int row = b_row0 + i * 64;
            const uint8_t *src = g.wpack + ((size_t)(kt * 4 + b_chunk) * g.wrows + n0 + row) * 16;
            cp_async16(sb + tile_off(row, b_chunk), src, true);
        }
        // ---- advance K decomposition by one tile (64 bytes) ----
        if constexpr (SMALL_C) {
            k_c += BK;
            while (k_c >= g.row_bytes) { k_c -= g.row_bytes; ++k_r; }
        } else {
            k_c += BK;
            while (k_c >= g.cin_pad) {
                k_c -= g.cin_pad;
                if (++k_s == g.kw) { k_s = 0; ++k_r; }
            }
        }
    };

    int32_t acc[2][NT][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < g.ktiles) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < g.ktiles; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // prefetch tile kt + STAGES - 1 into the slot freed by iteration kt-1
        {
            const int nk = kt + STAGES - 1;
            if (nk < g.ktiles) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const uint32_t sa = smem_base + (kt % STAGES) * STAGE;
        const uint32_t sb = sa + A_STAGE;
#pragma unroll
        for (int ks = 0; ks < BK / 32; ++ks) {
            uint32_t af[2][4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int row = wm * 32 + mi * 16 + (lane & 15);
                const int chunk = ks * 2 + (lane >> 4);
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], sa + tile_off(row, chunk));
            }
#pragma unroll
            for (int nj = 0; nj < NT / 2; ++nj) {
                uint32_t b0, b1, b2, b3;
                const int row = wn * WN + nj * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(b0, b1, b2, b3, sb + tile_off(row, chunk));
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    mma_i8<A_SIGNED>(acc[mi][nj * 2 + 0], af[mi], b0, b1);
                    mma_i8<A_SIGNED>(acc[mi][nj * 2 + 1], af[mi], b2, b3);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // pipeline smem is free from here on: reuse as the output staging tile

    // ---------------- fused epilogue ----------------
    constexpr int SPITCH = BN + 16;            // bytes per staged row (16B aligned, skewed)
    uint8_t *stage0 = smem;
    uint8_t *stage1 = smem + BM * SPITCH;
    const int g4 = lane >> 2, t4 = lane & 3;
    const bool has_carry = ep.carry_in != nullptr;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = wm * 32 + mi * 16 + g4 + half * 8;
            const int m = m0 + row;
            const bool row_ok = m < g.M;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = wn * WN + nt * 8 + t4 * 2;
                const int gc = n0 + col;
                const bool ok = row_ok && gc < ep.cout_pad;
                int32_t v0 = acc[mi][nt][half * 2 + 0];
                int32_t v1 = acc[mi][nt][half * 2 + 1];
                if (ok) {
                    const int2 b = *reinterpret_cast<const int2 *>(ep.bias + gc);
                    v0 = (int32_t)((uint32_t)v0 + (uint32_t)b.x);
                    v1 = (int32_t)((uint32_t)v1 + (uint32_t)b.y);
                    int2 c = make_int2(0, 0);
                    if (has_carry)
                        c = *reinterpret_cast<const int2 *>(ep.carry_in + f8::carry_off((size_t)m, gc, ep.cout_pad));
                    v0 = f8::residual_relu(v0, has_carry, c.x, ep.carry_shift, ep.relu);
                    v1 = f8::residual_relu(v1, has_carry, c.y, ep.carry_shift, ep.relu);
                    if (ep.carry_out)
                        *reinterpret_cast<int2 *>(ep.carry_out + f8::carry_off((size_t)m, gc, ep.cout_pad)) =
                            make_int2(v0, v1);
                    if (ep.out_f32) {
                        float *o = ep.out_f32 + (size_t)m * ep.out_f32_ld + gc;
                        if (gc < ep.cout) o[0] = (float)v0;
                        if (gc + 1 < ep.cout) o[1] = (float)v1;
                    }
                }
                if (ep.out0) {
                    const uint32_t q0 = (uint32_t)f8::requant(v0, ep.shift0, ep.signed0) & 0xffu;
                    const uint32_t q1 = (uint32_t)f8::requant(v1, ep.shift0, ep.signed0) & 0xffu;
                    *reinterpret_cast<uint16_t *>(stage0 + row * SPITCH + col) =
                        (uint16_t)(q0 | (q1 << 8));
                }
                if (ep.out1) {
                    const uint32_t q0 = (uint32_t)f8::requant(v0, ep.shift1, ep.signed1) & 0xffu;
                    const uint32_t q1 = (uint32_t)f8::requant(v1, ep.shift1, ep.signed1) & 0xffu;
                    *reinterpret_cast<uint16_t *>(stage1 + row * SPITCH + col) =
                        (uint16_t)(q0 | (q1 << 8));
                }
            }
        }
    }
    if (ep.out0 || ep.out1) {
        __syncthreads();
        constexpr int CHUNKS = BN / 16;
        for (int idx = tid; idx < BM * CHUNKS; idx += THREADS) {
            const int row = idx / CHUNKS, ch = idx - row * CHUNKS;
            const int m = m0 + row, gc = n0 + ch * 16;
            if (m < g.M && gc < ep.cout_pad) {
                if (ep.out0)
                    *reinterpret_cast<uint4 *>(ep.out0 + (size_t)m * ep.cout_pad + gc) =
                        *reinterpret_cast<const uint4 *>(stage0 + row * SPITCH + ch * 16);
                if (ep.out1)
                    *reinterpret_cast<uint4 *>(ep.out1 + (size_t)m * ep.cout_pad + gc) =
                        *reinterpret_cast<const uint4 *>(stage1 + row * SPITCH + ch * 16);
            }
        }
    }
}

template <int BN, bool A_SIGNED, bool SMALL_C>
int launch_t(const ConvGeom &g, const f8::Epilogue &ep, int ntiles_n, cudaStream_t s) {
    constexpr int smem_pipe = STAGES * (BM * BK + BN * BK);
    constexpr int smem_epi = 2 * BM * (BN + 16);
    constexpr int smem_bytes = smem_pipe > smem_epi ? smem_pipe : smem_epi;
    auto kern = conv_mma_kernel<BN, A_SIGNED, SMALL_C>;
    static f8host::DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = f8host::device_once(once, &num_sms, [&]() -> int {
            F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            return F8_OK;
        });
        if (rc) return rc;
    }
    dim3 grid((g.M + BM - 1) / BM, ntiles_n);
    f8host::note_kernel("conv_mma<BN=%d%s>", BN, SMALL_C ? ",small_c" : "");
    kern<<<grid, THREADS, smem_bytes, s>>>(g, ep);
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

}  // namespace

namespace f8host {

DensePack dense_pack_geometry(int cin_pad, int cout_pad, int kh, int kw) {
    DensePack p{};
    if (cin_pad == 4) {
        p.mode = 1;
        p.shift_px = 1;                                   // keeps the window start even (8B aligned)
        const int px = (kw + p.shift_px + 1) / 2 * 2;     // 7x7 -> 8 pixels, 3x3 -> 4 pixels
        p.row_bytes = px * 4;
        p.K = kh * p.row_bytes;
    } else {
        p.mode = 0;
        p.K = kh * kw * cin_pad;
    }
    p.K_pad = (p.K + 63) / 64 * 64;
    p.rows = (cout_pad + 255) / 256 * 256;
    return p;
}

int launch_conv_mma(const f8_conv_args &a, cudaStream_t s) {
    if (a.cin_pad != 4 && a.cin_pad % 16 != 0) {
        set_error("conv_dense: cin_pad %d must be 4 or a multiple of 16", a.cin_pad);
        return F8_ERR_ARG;
    }
    if (a.cout_pad % 16 != 0) {
        set_error("conv_dense: cout_pad %d must be a multiple of 16", a.cout_pad);
        return F8_ERR_ARG;
    }
    const DensePack pk = dense_pack_geometry(a.cin_pad, a.cout_pad, a.kh, a.kw);
    if (pk.mode == 1 && ((a.stride & 1) || ((a.pad + pk.shift_px) & 1) || (a.win & 1))) {
        set_error("conv_dense: small-C mode needs even stride, odd pad and even width");
        return F8_ERR_UNSUPPORTED;
    }
    const long long M = (long long)a.n * a.hout * a.wout;
    if (M <= 0 || M > 0x7fffffffLL) {
        set_error("conv_dense: pixel count %lld out of range", M);
        return F8_ERR_ARG;
    }
    ConvGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.wpack = static_cast<const uint8_t *>(a.wpack);
    g.M = (int)M;
    g.hin = a.hin; g.win = a.win; g.cin_pad = a.cin_pad;
    g.hout = a.hout; g.wout = a.wout;
    g.kh = a.kh; g.kw = a.kw; g.stride = a.stride; g.pad = a.pad;
    g.K_pad = pk.K_pad;
    g.wrows = pk.rows;
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
    // tile width: the one that wastes fewer padded columns; ties go to 128
    const int t64 = (a.cout_pad + 63) / 64, t128 = (a.cout_pad + 127) / 128;
    const bool use64 = t64 * 64 < t128 * 128;
    const bool sgn = a.in_signed != 0;
    if (pk.mode == 1) {
        if (use64) return sgn ? launch_t<64, true, true>(g, ep, t64, s) : launch_t<64, false, true>(g, ep, t64, s);
        return sgn ? launch_t<128, true, true>(g, ep, t128, s) : launch_t<128, false, true>(g, ep, t128, s);
    }
    if (use64) return sgn ? launch_t<64, true, false>(g, ep, t64, s) : launch_t<64, false, false>(g, ep, t64, s);
    return sgn ? launch_t<128, true, false>(g, ep, t128, s) : launch_t<128, false, false>(g, ep, t128, s);
}

}  // namespace f8host
