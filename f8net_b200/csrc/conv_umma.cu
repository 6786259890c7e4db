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
// Tile 128 x BN (BN = 64 | 128 | 256), K step 64 bytes per pipeline stage (two K=32 MMAs).
// PERSISTENT kernel, one or two CTAs per SM, each looping over its tiles with one operand
// ring and two TMEM accumulators, so the loads of tile i+1 and the epilogue of tile i-1
// overlap the MMAs of tile i.  Warp roles:
//   4 warps    producers.  Thread r owns tile row r (one output pixel): it gathers the
//              pixel's K bytes from the NHWC activation with 16-byte cp.async (zero fill for
//              the padding halo) into the canonical K-major no-swizzle operand layout
//              [K/16][128 rows][16 B] (core matrix = 8 rows x 16 B contiguous, SBO = 128 B,
//              LBO = 2048 B), fences the generic->async proxy and arrives on the stage's
//              "full" mbarrier.  Its thread 0 also posts the stage's weight bytes: 4 bulk
//              copies of BN*16 contiguous bytes from the chunk-major weight image.
//   1 warp     one elected lane issues the MMAs (tcgen05.mma is a single-thread instruction), commits
//              each stage to its "empty" mbarrier and each finished accumulator to
//              "acc_full".  It also owns the TMEM allocation (2 x BN columns).
//   epilogue   8 (BN <= 128, two CTAs per SM) or 16 (BN = 256) warps: warp w reads TMEM lane
//              group w % 4 (lane = tile row) and a column slice, 16 columns at a time, runs the
//              integer epilogue of f8_common.cuh (residual carries prefetched two steps ahead)
//              and releases the accumulator ("acc_empty").
// Integer accumulation is associative mod 2^32: tiling and MMA order cannot change results.
#include <cstdlib>
#include <cstring>

#include "tma_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// epilogue warps: warp w reads TMEM lane group w % 4, column slice w / 4.  BN = 256 runs one CTA
// per SM with 16 of them, BN <= 128 two CTAs per SM with 8 each: 16 epilogue warps per SM.
__host__ __device__ constexpr int epi_warps_for(int bn) { return bn == 256 ? 16 : 8; }
constexpr int PRODUCERS = 128;
// gather path: 4 producer warps + the MMA warp; TMA path: one producer warp + the MMA warp (fewer
// threads = a higher register cap for the epilogue: 96 instead of 72 / 80)
__host__ __device__ constexpr int threads_for(int bn, bool tma_a) { return (epi_warps_for(bn) + (tma_a ? 2 : 5)) * 32; }
// ring depth per tile width: ~96-120 KB of operand bytes in flight per CTA
__host__ __device__ constexpr int stages_for(int bn) { return bn <= 64 ? 8 : (bn <= 128 ? 6 : 5); }
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

using namespace f8u;

// Persistent, warp-specialised kernel.  Static tile schedule: CTA b runs tiles b, b+grid, ...
// with the N tile fastest, so CTAs that are co-resident read the same activation rows.
// TMA_A (1x1 stride 1 convolutions and nn.Linear): the A tile of a stage is one
// TMA box {64 channels, 128 pixels} of the activation seen as a (C, M) matrix, landing in the
// 64-byte-swizzled K-major layout; rows past M are the out-of-bounds zero fill.
template <int BN, bool A_SIGNED, bool SMALL_C, bool TMA_A>
__global__ void __launch_bounds__(threads_for(BN, TMA_A), (BN <= 128) ? 2 : 1)
conv_umma_kernel(const UGeom g, const f8::Epilogue ep, const int mtiles, const int ntiles_n,
                 const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (f8::smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int EPI_WARPS = epi_warps_for(BN);
    constexpr int EPI_THREADS = EPI_WARPS * 32;
    constexpr int PRODUCER_WARP0 = EPI_WARPS;     // 4 warps: A gather (+ its thread 0: B bulk copies)
    constexpr int MMA_WARP = EPI_WARPS + (TMA_A ? 1 : 4);   // TMEM alloc, one elected lane issues tcgen05.mma
    constexpr int THREADS = threads_for(BN, TMA_A);
    constexpr int S = stages_for(BN);
    constexpr int B_STAGE = BN * BK;
    constexpr int B_CHUNK = BN * 16;
    constexpr int STAGE = A_STAGE + B_STAGE;
    const uint32_t smem_base = f8::smem_u32(smem);
    // after the ring: full[S] empty[S] acc_full[2] acc_empty[2] | tmem slot | bias[2][BN]
    const uint32_t bar_base = smem_base + S * STAGE;
    auto full_bar = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (uint32_t)(S + s) * 8; };
    auto acc_full_bar = [&](int b) { return bar_base + (uint32_t)(2 * S + b) * 8; };
    auto acc_empty_bar = [&](int b) { return bar_base + (uint32_t)(2 * S + 2 + b) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S * STAGE + (2 * S + 4) * 8);
    int32_t *sbias = reinterpret_cast<int32_t *>(smem + S * STAGE + (2 * S + 4) * 8 + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int total_tiles = mtiles * ntiles_n;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                mbar_init(full_bar(s), TMA_A ? 1 : PRODUCERS);
                mbar_init(empty_bar(s), 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(acc_full_bar(b), 1);
                mbar_init(acc_empty_bar(b), EPI_THREADS);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 2 * BN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: the prologue above overlaps the previous layer's tail; the
    // trigger comes after this CTA holds its TMEM columns (a dependent CTA must never take them first)
    f8::pdl_trigger();
    f8::pdl_wait();

    if (TMA_A && warp >= PRODUCER_WARP0 && warp < MMA_WARP) {
        // =========================== producer: one thread, A by TMA + B bulk copies =====
        if (tid == PRODUCER_WARP0 * 32) {
            tma_prefetch_desc(&tmap);
            int slot = 0, phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int mt = t / ntiles_n;
                const int n0 = (t - mt * ntiles_n) * BN;
                for (int kt = 0; kt < g.ktiles; ++kt) {
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    const uint32_t sa = smem_base + slot * STAGE;
                    mbar_expect_tx(full_bar(slot), A_STAGE + B_STAGE);
                    mbar_arrive(full_bar(slot));
                    tma_load_4d(sa, &tmap, kt * BK, mt * BM, 0, 0, full_bar(slot));
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        bulk_g2s(sa + A_STAGE + j * B_CHUNK, g.wpack + ((size_t)(kt * 4 + j) * g.wrows + n0) * 16,
                                 B_CHUNK, full_bar(slot));
                    if (++slot == S) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= PRODUCER_WARP0 && warp < MMA_WARP) {
        // =========================== producers: A gather + B bulk copies ==========
        const int row = tid - PRODUCER_WARP0 * 32;
        const int HW = g.hout * g.wout;
        int slot = 0, phase = 0;       // ring position of the stage being issued
        int aslot = 0;                 // ring position of the next stage to signal
        int issued = 0;                // stages issued and not yet signalled
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int mt = t / ntiles_n;
            const int n0 = (t - mt * ntiles_n) * BN;
            const int m = mt * BM + row;
            const bool valid = m < g.M;
            const int mm = valid ? m : 0;
            const int img = mm / HW;
            const int rem = mm - img * HW;
            const int p = rem / g.wout, q = rem - p * g.wout;
            const int ih0 = p * g.stride - g.pad;
            const int iw0 = q * g.stride - g.pad - (SMALL_C ? g.shift_px : 0);
            const uint8_t *base = g.in + (size_t)img * g.hin * g.win * g.cin_pad;
            int k_r = 0, k_s = 0, k_c = 0;
            for (int kt = 0; kt < g.ktiles; ++kt) {
                mbar_wait(empty_bar(slot), phase ^ 1);
                const uint32_t sa = smem_base + slot * STAGE;
                if (row == 0) {
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
                if (++slot == S) { slot = 0; phase ^= 1; }
                // signal with a lag of S/2 stages: deep enough to cover the load latency, and
                // never waiting for a slot the MMA of the immediately preceding stage holds
                if (++issued > S / 2) {
                    cp_async_wait<S / 2>();       // the oldest unsignalled stage has landed
                    fence_proxy_async();          // generic-proxy writes -> visible to the MMA
                    mbar_arrive(full_bar(aslot));
                    if (++aslot == S) aslot = 0;
                    --issued;
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (; issued > 0; --issued) {
            mbar_arrive(full_bar(aslot));
            if (++aslot == S) aslot = 0;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ==================================
        // whole warp in the loop (uniform control flow), one elected lane issues; descriptor
        // high words are constants, low words advance by 32-bit adds
        constexpr uint32_t idesc = instr_desc(A_SIGNED, BN);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);          // SBO = 128 B, version 1
        // TMA_A: SWIZZLE_64B (layout type 4), SBO = 8 rows x 64 B, second K half at +32 B
        constexpr uint32_t desc_hi_a = TMA_A ? ((512u >> 4) | (1u << 14) | (4u << 29)) : desc_hi;
        constexpr uint32_t a_lbo_field = TMA_A ? (1u << 16) : (((uint32_t)A_CHUNK >> 4) << 16);
        constexpr uint32_t a_khalf = TMA_A ? 2u : (uint32_t)((2 * A_CHUNK) >> 4);
        constexpr uint32_t b_lbo_field = ((uint32_t)B_CHUNK >> 4) << 16;
        int slot = 0, phase = 0;
        int buf = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            mbar_wait(acc_empty_bar(buf), acc_phase ^ 1);   // epilogue drained this buffer
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
            for (int kt = 0; kt < g.ktiles; ++kt) {
                mbar_wait(full_bar(slot), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + slot * STAGE;
                const uint32_t a_lo = ((sa & 0x3ffffu) >> 4) | a_lbo_field;
                const uint32_t b_lo = (((sa + A_STAGE) & 0x3ffffu) >> 4) | b_lbo_field;
                if (elect_one()) {
                    umma_i8_lohi(tacc, a_lo, desc_hi_a, b_lo, desc_hi, idesc, (uint32_t)(kt != 0));
                    umma_i8_lohi(tacc, a_lo + a_khalf, desc_hi_a, b_lo + ((2 * B_CHUNK) >> 4),
                                 desc_hi, idesc, 1u);
                    umma_commit(empty_bar(slot));   // frees the stage once these MMAs have read it
                }
                __syncwarp();
                if (++slot == S) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full_bar(buf));     // accumulator complete -> epilogue
            __syncwarp();
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
    } else {
        // =========================== epilogue ====================================
        // Warp (lg, cs): TMEM lane group lg = rows [32*lg, 32*lg+32) of the tile, column slice cs of
        // CW columns, walked in steps of 16 columns.  Residual carries (pixel-interleaved layout,
        // f8_common.cuh: contiguous 512-byte runs per warp access) are requested one or two steps ahead
        // into registers -- across tile boundaries -- so that their DRAM latency is hidden.
        const int lg = warp & 3;
        const int cs = warp >> 2;
        constexpr int CW = BN / (EPI_WARPS / 4);
        constexpr int NS = CW / 16;                    // steps per tile for this warp
        const int row = lg * 32 + lane;                // TMEM lane == tile row
        const bool has_carry = ep.carry_in != nullptr;
        const bool plain = f8::epilogue_is_plain_u8(ep);
        const f8::EpiConst kc = f8::epi_const(ep, has_carry);
        // prefetch cursor: (tile, step) of the carry request two steps ahead of the consumer
        int pf_t = blockIdx.x, pf_s = 0;
        constexpr int PF = (BN == 256) ? 2 : 1;        // prefetch depth in steps (2 CTAs per SM need less)
        int4 cq[PF][4] = {};
        auto request = [&](int4 (&dst)[4]) {
            if (has_carry && pf_t < total_tiles) {
                const int mt = pf_t / ntiles_n;
                const int col = (pf_t - mt * ntiles_n) * BN + cs * CW + 16 * pf_s;
                const int m = mt * BM + row;
                if (m < g.M && col < ep.cout_pad) {
                    const int32_t *src = ep.carry_in + f8::carry_off((size_t)m, col, ep.cout_pad);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k] = __ldg(reinterpret_cast<const int4 *>(src + k * 512));
                }
            }
            if (++pf_s == NS) { pf_s = 0; pf_t += gridDim.x; }
        };
#pragma unroll
        for (int d = 0; d < PF; ++d) request(cq[d]);
        int buf = 0, acc_phase = 0;
        int bias_n0[2] = {-1, -1};                     // N tile whose bias each buffer's shared copy holds
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int mt = t / ntiles_n;
            const int n0 = (t - mt * ntiles_n) * BN;
            const int m = mt * BM + row;
            const bool valid = m < g.M;
            int ncols = ep.cout_pad - n0;
            if (ncols > BN) ncols = BN;
            int32_t *bias_s = sbias + buf * BN;
            if (bias_n0[buf] != n0) {            // warp-uniform; a single N tile loads its bias twice per launch
                // every warp is done with the tile that last read this copy before it is rewritten
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                for (int i = tid; i < ncols; i += EPI_THREADS) {
                    int32_t b = __ldg(ep.bias + n0 + i);
                    if (plain) b = (int32_t)((uint32_t)b + (1u << (ep.shift0 - 1)));   // bias + half
                    bias_s[i] = b;
                }
                bias_n0[buf] = n0;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            mbar_wait(acc_full_bar(buf), acc_phase);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll
            for (int sidx = 0; sidx < NS; ++sidx) {
                const int c0 = cs * CW + 16 * sidx;
                int4 c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    c[k] = cq[0][k];
                    if (PF == 2) cq[0][k] = cq[PF - 1][k];
                }
                request(cq[PF - 1]);
                if (c0 < ncols) {                               // warp-uniform
                    int32_t v[16];
                    tmem_ld16(trow + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (valid) {
                        const int col = n0 + c0;
                        const size_t o = (size_t)m * ep.cout_pad + col;
                        if (plain) {
                            f8::epilogue16_plain_u8(v, bias_s + c0, ep.out0 + o, ep.shift0);
                        } else {
                            f8::epilogue16_math(v, bias_s + c0, kc, c, has_carry);
                            if (ep.carry_out) {
                                int32_t *dst = ep.carry_out + f8::carry_off((size_t)m, col, ep.cout_pad);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    *reinterpret_cast<int4 *>(dst + k * 512) =
                                        make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                            }
                            if (ep.out0)
                                *reinterpret_cast<uint4 *>(ep.out0 + o) = f8::requant_pack16(v, ep.shift0, ep.signed0);
                            if (ep.out1)
                                *reinterpret_cast<uint4 *>(ep.out1 + o) = f8::requant_pack16(v, ep.shift1, ep.signed1);
                            if (ep.out_f32) {
                                float *f = ep.out_f32 + (size_t)m * ep.out_f32_ld + col;
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (col + i < ep.cout) f[i] = (float)v[i];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty_bar(buf));       // this thread's columns are drained
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

template <int BN>
constexpr int smem_bytes_for() {
    return stages_for(BN) * (A_STAGE + BN * BK) + (2 * stages_for(BN) + 4) * 8 + 16 + 2 * BN * 4 + 1024;
}

template <int BN, bool A_SIGNED, bool SMALL_C, bool TMA_A>
int launch_t(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s) {
    constexpr int smem_bytes = smem_bytes_for<BN>();
    auto kern = conv_umma_kernel<BN, A_SIGNED, SMALL_C, TMA_A>;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (TMA_A) {
        const uint64_t dims[4] = {(uint64_t)g.cin_pad, (uint64_t)g.M, 1u, 1u};
        const uint64_t row = (uint64_t)g.cin_pad, all = (uint64_t)g.cin_pad * (uint64_t)g.M;
        const uint64_t strides[3] = {row, all, all};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)BM, 1u, 1u};
        const int rc = f8host::encode_tmap_u8_4d(&tmap, g.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc != F8_OK) return rc;
    }
    static bool attr_done = false;
    static int num_sms = 0;
    if (!attr_done) {
        F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        int dev = 0;
        F8_CUDA(cudaGetDevice(&dev));
        F8_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    const int mtiles = (g.M + BM - 1) / BM;
    const int ntn = (ep.cout_pad + BN - 1) / BN;
    const long long total = (long long)mtiles * ntn;
    // 2 x BN TMEM columns and the smem ring per CTA: two co-resident CTAs for BN <= 128
    const int per_sm = (BN <= 128) ? 2 : 1;
    long long grid = (long long)num_sms * per_sm;
    if (grid > total) grid = total;
    F8_CUDA(f8host::launch_pdl(kern, (unsigned)grid, threads_for(BN, TMA_A), smem_bytes, s, g, ep, mtiles, ntn, tmap));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

template <int BN>
int launch_bn(const UGeom &g, const f8::Epilogue &ep, bool sgn, bool small_c, bool tma_a, cudaStream_t s) {
    if (small_c) return sgn ? launch_t<BN, true, true, false>(g, ep, s) : launch_t<BN, false, true, false>(g, ep, s);
    if (tma_a) return sgn ? launch_t<BN, true, false, true>(g, ep, s) : launch_t<BN, false, false, true>(g, ep, s);
    return sgn ? launch_t<BN, true, false, false>(g, ep, s) : launch_t<BN, false, false, false>(g, ep, s);
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
    // 1x1 stride 1 (and nn.Linear): the A operand is a plain (C, M) matrix -> TMA
    static const bool no_tma = getenv("F8_GATHER_NO_TMA") != nullptr;
    const bool tma_a = !no_tma && !small_c && a.kh == 1 && a.kw == 1 && a.stride == 1 && a.pad == 0 &&
                       a.cin_pad % 16 == 0 && (g.M * (long long)a.cin_pad) < (1LL << 40);   // channels past cin_pad: zero fill
    if (a.cout_pad <= 64) return launch_bn<64>(g, ep, sgn, small_c, tma_a, s);
    if (a.cout_pad <= 128) return launch_bn<128>(g, ep, sgn, small_c, tma_a, s);
    return launch_bn<256>(g, ep, sgn, small_c, tma_a, s);
}

}  // namespace f8host
