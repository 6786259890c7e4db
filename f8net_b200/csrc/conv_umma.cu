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
#include <cstdio>
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
// TMA path: a ring stage is G K-steps of 64 bytes (G TMA boxes + 4 G weight chunks behind ONE mbarrier pair),
// so that the MMA warp and the producer warp pay their stage boundary -- several hundred cycles of
// shared-memory-path latency each, profiles/r02_mma_loop_experiments.md -- once per 2 G MMAs instead of once per 2
__host__ __device__ constexpr int ksteps_for(int bn, bool tma_a) { return !tma_a ? 1 : (bn == 64 ? 4 : 2); }
__host__ __device__ constexpr int ring_for(int bn, bool tma_a) { return !tma_a ? stages_for(bn) : (bn == 64 ? 2 : 3); }
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
    int probe;             // debug build: timing probes (WRONG results)
};

using namespace f8u;

// Persistent, warp-specialised kernel.  Static tile schedule: CTA b runs tiles b, b+grid, ...
// with the N tile fastest, so CTAs that are co-resident read the same activation rows.
// TMA_A (1x1 stride 1 convolutions and nn.Linear): the A tile of a stage is one
// TMA box {64 channels, 128 pixels} of the activation seen as a (C, M) matrix, landing in the
// 64-byte-swizzled K-major layout; rows past M are the out-of-bounds zero fill.
// PLAIN (compile time, TMA path): bias + ReLU + one unsigned right-shift consumer and nothing else -- the epilogue of
// most point-wise layers.  ncu showed the generic epilogue latency bound on tcgen05.ld (IPC 1.1, long-scoreboard
// stalls, 11 instructions per output value against the 4.5 the arithmetic needs): the specialised one drops the
// carry cursor and software-pipelines the TMEM loads one 16-column step ahead.
template <int BN, bool A_SIGNED, bool SMALL_C, bool TMA_A, bool PLAIN = false>
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
    constexpr int S = ring_for(BN, TMA_A);
    constexpr int G = ksteps_for(BN, TMA_A);          // K = 64 steps per ring stage
    constexpr int B_STAGE = BN * BK;
    constexpr int B_CHUNK = BN * 16;
    constexpr int KSTEP = A_STAGE + B_STAGE;          // one K = 64 step: A tile then B tile
    constexpr int STAGE = G * KSTEP;
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
                mbar_init(acc_empty_bar(b), EPI_WARPS);        // one arrival per epilogue warp
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
        // =========================== producer warp: A by TMA + B bulk copies ===========
        // Under the tensor core's operand traffic every shared-memory-path instruction (mbarrier op, TMA /
        // bulk-copy issue) takes the issuing warp 130-190 cycles (tools/probes/mma5_probe.cu), and a stage is
        // only 128-256 cycles of MMAs: the stage is announced with ONE arrive.expect_tx, lane 0 issues the
        // TMA box and lanes 1-4 the four weight chunks in one warp instruction.
        if (warp == PRODUCER_WARP0) {
            if (lane == 0) tma_prefetch_desc(&tmap);
            int slot = 0, phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int mt = t / ntiles_n;
                const int n0 = (t - mt * ntiles_n) * BN;
                for (int kt0 = 0; kt0 < g.ktiles; kt0 += G) {
                    const int kcnt = g.ktiles - kt0 < G ? g.ktiles - kt0 : G;      // K steps of this stage
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    const uint32_t sa = smem_base + slot * STAGE;
                    const bool skip_b = F8_DBG && (g.probe & 1) && t != (int)blockIdx.x;     // probes: stale operands (WRONG results)
                    const bool skip_a = F8_DBG && (g.probe & 2) && t != (int)blockIdx.x;
                    if (lane == 0) mbar_arrive_expect_tx(full_bar(slot), (uint32_t)(kcnt * ((skip_a ? 0 : A_STAGE) + (skip_b ? 0 : B_STAGE))));
                    __syncwarp();
                    // lanes 0..G-1: the TMA box of K step `lane`; lanes 8..8+4G-1: its four weight chunks
                    if (lane < kcnt && !skip_a) tma_load_4d(sa + lane * KSTEP, &tmap, (kt0 + lane) * BK, mt * BM, 0, 0, full_bar(slot));
                    if (lane >= 8 && lane < 8 + 4 * kcnt && !skip_b) {
                        const int ks = (lane - 8) >> 2, j = (lane - 8) & 3;
                        bulk_g2s(sa + ks * KSTEP + A_STAGE + j * B_CHUNK,
                                 g.wpack + ((size_t)((kt0 + ks) * 4 + j) * g.wrows + n0) * 16, B_CHUNK, full_bar(slot));
                    }
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
            for (int kt0 = 0; kt0 < g.ktiles; kt0 += G) {
                const int kcnt = g.ktiles - kt0 < G ? g.ktiles - kt0 : G;
                mbar_wait(full_bar(slot), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + slot * STAGE;
                const uint32_t a_lo = ((sa & 0x3ffffu) >> 4) | a_lbo_field;
                const uint32_t b_lo = (((sa + A_STAGE) & 0x3ffffu) >> 4) | b_lbo_field;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (j < kcnt) {
                            const uint32_t o = (uint32_t)(j * (KSTEP >> 4));
                            umma_i8_lohi(tacc, a_lo + o, desc_hi_a, b_lo + o, desc_hi, idesc, (uint32_t)((kt0 + j) != 0));
                            umma_i8_lohi(tacc, a_lo + o + a_khalf, desc_hi_a, b_lo + o + ((2 * B_CHUNK) >> 4),
                                         desc_hi, idesc, 1u);
                        }
                    }
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
        const bool has_carry = !PLAIN && ep.carry_in != nullptr;
        const bool plain = PLAIN || f8::epilogue_is_plain_u8(ep);
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
            if constexpr (PLAIN) {
                // two register sets: the load of step s + 1 is in flight while step s is requantised and stored
                int32_t va[16], vb[16];
                if (cs * CW < ncols) tmem_ld16(trow + (uint32_t)(cs * CW), va);
#pragma unroll
                for (int sidx = 0; sidx < NS; ++sidx) {
                    const int c0 = cs * CW + 16 * sidx;
                    if (c0 < ncols) {                           // warp-uniform
                        tmem_ld_wait();
                        if (sidx + 1 < NS && c0 + 16 < ncols) tmem_ld16(trow + (uint32_t)(c0 + 16), (sidx & 1) ? va : vb);
                        if (valid)
                            f8::epilogue16_plain_u8((sidx & 1) ? vb : va, bias_s + c0,
                                                    ep.out0 + (size_t)m * ep.cout_pad + n0 + c0, ep.shift0);
                    }
                }
            } else
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
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(buf));       // this warp's columns are drained
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


// ------------------------------------------------------------------------------------------
// conv1x1_res_kernel -- 1x1 stride-1 convolutions with FEW channels and MANY pixels (the early
// MobileNet point-wise layers, the 56x56 bottleneck 1x1s of ResNet50): per 128-pixel tile the
// generic kernel above spends more on pipeline hand-overs (a weight re-fetch of 4 bulk copies, two
// mbarrier round trips, one accumulator hand-over) than on moving its ~6 KB.  Here
//   * the layer's whole weight image (K_pad x BN bytes, one N tile) is RESIDENT in shared memory,
//     fetched once per CTA -- before griddepcontrol.wait, i.e. under the previous layer's tail;
//   * a tile is MSEG = 256 / BN segments of 128 pixels: one ring stage = MSEG TMA boxes of one
//     64-channel K step, one accumulator set = MSEG x BN TMEM columns (two sets = 512 columns),
//     so every mbarrier round trip is amortised over 512 (BN = 64) or 256 (BN = 128) pixels;
//   * 16 epilogue warps: warp (lane group, unit) owns 32 pixels x 64 columns of one segment.
// One CTA per SM, persistent.  Same operand layouts, descriptors and epilogue as the TMA_A path.
#define R_TIMED(acc, stmt)                       \
    do {                                         \
        if (F8_DBG && stats) {                   \
            const long long _t0 = clock64();     \
            stmt;                                \
            acc += clock64() - _t0;              \
        } else {                                 \
            stmt;                                \
        }                                        \
    } while (0)
constexpr int R_EPI_WARPS = 16;
constexpr int R_THREADS = (R_EPI_WARPS + 2) * 32;
constexpr int R_SMAX = 6;

template <int BN, bool A_SIGNED, bool PLAIN>
__global__ void __launch_bounds__(R_THREADS, 1)
conv1x1_res_kernel(const UGeom g, const f8::Epilogue ep, const int mtiles, const int S,
                   const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap,
                   long long *stats, const int probe, const int cpa) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (f8::smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int MSEG = 256 / BN;
    constexpr int TM = MSEG * BM;
    constexpr int EPI_THREADS = R_EPI_WARPS * 32;
    constexpr int PRODUCER_WARP = R_EPI_WARPS;
    constexpr int MMA_WARP = R_EPI_WARPS + 1;
    constexpr int STAGE = MSEG * A_STAGE;
    constexpr int B_CHUNK = BN * 16;
    const int w_bytes = g.ktiles * 4 * B_CHUNK;
    // 8-bit output staging: one 32-pixel x 64-byte box per epilogue warp (64-byte swizzle), written
    // to global memory by TMA -- a warp's direct stores would be 32 16-byte pieces at a cout_pad
    // pitch (one L1 wavefront each)
    constexpr int OSTAGE = 32 * 64;
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t o_base = smem_base + S * STAGE;
    const uint32_t w_base = o_base + R_EPI_WARPS * OSTAGE;
    const uint32_t bar_base = w_base + w_bytes;
    auto full_bar = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (uint32_t)(R_SMAX + s) * 8; };
    auto acc_full_bar = [&](int b) { return bar_base + (uint32_t)(2 * R_SMAX + b) * 8; };
    auto acc_empty_bar = [&](int b) { return bar_base + (uint32_t)(2 * R_SMAX + 2 + b) * 8; };
    const uint32_t w_full = bar_base + (uint32_t)(2 * R_SMAX + 4) * 8;
    uint8_t *after = smem + S * STAGE + R_EPI_WARPS * OSTAGE + w_bytes + (2 * R_SMAX + 5) * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(after);
    int32_t *sbias = reinterpret_cast<int32_t *>(after + 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            // acc_empty: one arrival per epilogue WARP (512 per-thread arrivals on one mbarrier are 512
            // serialised shared-memory atomics per tile)
            for (int b = 0; b < 2; ++b) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), R_EPI_WARPS); }
            mbar_init(w_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 512);
    }
    constexpr bool plain = PLAIN;       // f8::epilogue_is_plain_u8(ep), decided by the launcher
    if (tid < BN) {                     // one N tile: the bias copy is loaded once
        int32_t b = tid < ep.cout_pad ? __ldg(ep.bias + tid) : 0;
        if (plain) b = (int32_t)((uint32_t)b + (1u << (ep.shift0 - 1)));   // bias + half
        sbias[tid] = b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == PRODUCER_WARP && cpa) {
        // ---- inputs of 16 / 32 channels: a TMA box over such narrow rows costs ~5 cycles per row, so the
        // warp copies the pixels with 16-byte cp.async straight into the canonical no-swizzle K-major
        // layout [chunk][TM rows][16 B] (cpa = chunks per pixel); the K = 64 stage becomes ONE K = 32 MMA
        if (lane == 0) {
            tma_prefetch_desc(&omap);
            mbar_expect_tx(w_full, (uint32_t)w_bytes);
            mbar_arrive(w_full);
            for (int kc = 0; kc < g.ktiles * 4; ++kc)
                bulk_g2s(w_base + kc * B_CHUNK, g.wpack + (size_t)kc * g.wrows * 16, B_CHUNK, w_full);
        }
        f8::pdl_wait();                         // the activation is the previous launch's output
        constexpr int LAG = 2;                  // a stage is signalled once LAG younger ones are issued
        int slot = 0, phase = 0, aslot = 0, issued = 0;
        for (int t = blockIdx.x; t < mtiles; t += gridDim.x) {
            mbar_wait(empty_bar(slot), phase ^ 1);
            const uint32_t sa = smem_base + slot * STAGE;
            for (int idx = lane; idx < TM * cpa; idx += 32) {
                const int r = cpa == 2 ? idx >> 1 : idx, ch = cpa == 2 ? idx & 1 : 0;
                const int m = t * TM + r;
                const bool ok = m < g.M;
                cp_async16(sa + ch * (TM * 16) + r * 16, g.in + (size_t)(ok ? m : 0) * g.cin_pad + ch * 16, ok);
            }
            cp_async_commit();
            if (++slot == S) { slot = 0; phase ^= 1; }
            if (++issued > LAG) {
                cp_async_wait<LAG>();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(aslot));
                if (++aslot == S) aslot = 0;
                --issued;
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        for (; issued > 0; --issued) {
            if (lane == 0) mbar_arrive(full_bar(aslot));
            if (++aslot == S) aslot = 0;
        }
    } else if (warp == PRODUCER_WARP) {
        {
            if (lane == 0) {
                tma_prefetch_desc(&tmap);
                tma_prefetch_desc(&omap);
                // resident weights: chunk-major image rows [0, BN) of every 16-byte K chunk
                mbar_arrive_expect_tx(w_full, (uint32_t)w_bytes);
            }
            __syncwarp();
            for (int kc = lane; kc < g.ktiles * 4; kc += 32)        // one chunk per lane per warp instruction
                bulk_g2s(w_base + kc * B_CHUNK, g.wpack + (size_t)kc * g.wrows * 16, B_CHUNK, w_full);
            f8::pdl_wait();                     // the activation is the previous launch's output
            int slot = 0, phase = 0;
            long long st_a = 0;
            const long long st_t0 = clock64();
            for (int t = blockIdx.x; t < mtiles; t += gridDim.x) {
                int nseg = (g.M - t * TM + BM - 1) / BM;       // segments with at least one real pixel
                if (nseg > MSEG) nseg = MSEG;
                for (int kt = 0; kt < g.ktiles; ++kt) {
                    R_TIMED(st_a, mbar_wait(empty_bar(slot), phase ^ 1));
                    const uint32_t sa = smem_base + slot * STAGE;
                    if (lane == 0) mbar_arrive_expect_tx(full_bar(slot), (uint32_t)(nseg * A_STAGE));
                    __syncwarp();
                    // lane sg issues the TMA box of segment sg: one warp instruction for the whole stage
                    if (lane < nseg)
                        tma_load_4d(sa + lane * A_STAGE, &tmap, kt * BK, (t * MSEG + lane) * BM, 0, 0, full_bar(slot));
                    if (++slot == S) { slot = 0; phase ^= 1; }
                }
            }
            if (F8_DBG && stats && lane == 0) { stats[blockIdx.x * 16 + 0] = clock64() - st_t0; stats[blockIdx.x * 16 + 1] = st_a; }
        }
    } else if (warp == MMA_WARP) {
        constexpr uint32_t idesc = instr_desc(A_SIGNED, BN);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                         // B: SBO = 128 B
        constexpr uint32_t desc_hi_a = (512u >> 4) | (1u << 14) | (4u << 29);          // A: SWIZZLE_64B
        constexpr uint32_t a_lbo_field = 1u << 16;
        constexpr uint32_t b_lbo_field = ((uint32_t)B_CHUNK >> 4) << 16;
        int slot = 0, phase = 0, buf = 0, acc_phase = 0;
        mbar_wait(w_full, 0);
        long long st_acc = 0, st_full = 0;
        const long long st_t0 = clock64();
        for (int t = blockIdx.x; t < mtiles; t += gridDim.x) {
            R_TIMED(st_acc, mbar_wait(acc_empty_bar(buf), acc_phase ^ 1));
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * 256);
            for (int kt = 0; kt < g.ktiles; ++kt) {
                R_TIMED(st_full, mbar_wait(full_bar(slot), phase));
                tc_fence_after();
                const uint32_t sa = smem_base + slot * STAGE;
                const uint32_t a_lo = ((sa & 0x3ffffu) >> 4) | a_lbo_field;
                const uint32_t b_lo = (((w_base + kt * 4 * B_CHUNK) & 0x3ffffu) >> 4) | b_lbo_field;
                if (cpa) {
                    // no-swizzle operand: K chunk 1 sits TM rows behind chunk 0 (16-channel inputs: its weights are zero)
                    const uint32_t a_lo_c = ((sa & 0x3ffffu) >> 4) | ((uint32_t)((TM * 16) >> 4) << 16);
                    if (elect_one()) {
#pragma unroll
                        for (int sg = 0; sg < MSEG; ++sg)
                            umma_i8_lohi(tacc + (uint32_t)(sg * BN), a_lo_c + (uint32_t)(sg * ((BM * 16) >> 4)), desc_hi,
                                         b_lo, desc_hi, idesc, 0u);
                        umma_commit(empty_bar(slot));
                    }
                } else if (elect_one()) {
#pragma unroll
                    for (int sg = 0; sg < MSEG; ++sg) {
                        umma_i8_lohi(tacc + (uint32_t)(sg * BN), a_lo + (uint32_t)(sg * (A_STAGE >> 4)), desc_hi_a,
                                     b_lo, desc_hi, idesc, (uint32_t)(kt != 0));
                        umma_i8_lohi(tacc + (uint32_t)(sg * BN), a_lo + (uint32_t)(sg * (A_STAGE >> 4)) + 2u, desc_hi_a,
                                     b_lo + ((2 * B_CHUNK) >> 4), desc_hi, idesc, 1u);
                    }
                    umma_commit(empty_bar(slot));
                }
                __syncwarp();
                if (++slot == S) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full_bar(buf));
            __syncwarp();
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
        if (F8_DBG && stats && lane == 0) {
            stats[blockIdx.x * 16 + 2] = clock64() - st_t0; stats[blockIdx.x * 16 + 3] = st_acc; stats[blockIdx.x * 16 + 4] = st_full;
        }
    } else {
        // =========================== epilogue (16 warps) =========================
        f8::pdl_wait();                                  // residual carry = an earlier launch's output
        constexpr int UPS = BN / 64;                     // 64-column units per segment
        const int lg = warp & 3;
        const int u = warp >> 2;
        const int seg = u / UPS;
        const int cbase = (u - seg * UPS) * 64;
        const int row = seg * BM + lg * 32 + lane;       // row inside the tile
        const bool has_carry = !PLAIN && ep.carry_in != nullptr;
        const f8::EpiConst kc = f8::epi_const(ep, has_carry);
        int pf_t = blockIdx.x, pf_s = 0;                 // carry prefetch cursor: one step ahead
        int4 cq[4] = {};
        auto request = [&](int4 (&dst)[4]) {
            if (PLAIN) return;
            if (has_carry && pf_t < mtiles) {
                const int col = cbase + 16 * pf_s;
                const int m = pf_t * TM + row;
                if (m < g.M && col < ep.cout_pad) {
                    const int32_t *src = ep.carry_in + f8::carry_off((size_t)m, col, ep.cout_pad);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k] = __ldg(reinterpret_cast<const int4 *>(src + k * 512));
                }
            }
            if (++pf_s == 4) { pf_s = 0; pf_t += gridDim.x; }
        };
        request(cq);
        // this thread's row of the warp's staging box; 16-byte chunk j of row r sits at chunk
        // j ^ ((r >> 1) & 3) (SWIZZLE_64B: conflict-free for the 16-byte row-per-lane writes)
        uint8_t *ostage = smem + S * STAGE + warp * OSTAGE + lane * 64;
        const int oswz = (lane >> 1) & 3;
        const bool out_unit = ep.out0 != nullptr && cbase < ep.cout_pad;      // warp-uniform
        int buf = 0, acc_phase = 0;
        long long st_w = 0, st_st = 0, st_p[4] = {0, 0, 0, 0};
        const long long st_t0 = clock64();
        for (int t = blockIdx.x; t < mtiles; t += gridDim.x) {
            const int m = t * TM + row;
            const bool valid = m < g.M;
            R_TIMED(st_w, mbar_wait(acc_full_bar(buf), acc_phase));
            tc_fence_after();
            long long ph0 = (F8_DBG && stats) ? clock64() : 0;
            if (out_unit) {            // the previous tile's store has finished reading the staging box
                if (lane == 0) R_TIMED(st_st, tma_store_wait_read());
                __syncwarp();
            }
            if (F8_DBG && stats) { const long long n = clock64(); st_p[0] += n - ph0; ph0 = n; }
            const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 256 + seg * BN);
            // the TMEM load of step s + 1 is in flight while step s is computed
            // (plain path only: the generic path has no registers to spare for a second buffer)
            int32_t va[16], vb[16];
            if (PLAIN && cbase < ep.cout_pad) tmem_ld16(trow + (uint32_t)cbase, va);
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const int c0 = cbase + 16 * sidx;
                int4 c[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = cq[k];
                request(cq);
                if (c0 < ep.cout_pad) {                         // warp-uniform
                    int32_t (&v)[16] = (PLAIN && (sidx & 1)) ? vb : va;
                    if (!PLAIN) tmem_ld16(trow + (uint32_t)c0, va);
                    tmem_ld_wait();
                    if (PLAIN && sidx < 3 && c0 + 16 < ep.cout_pad)
                        tmem_ld16(trow + (uint32_t)(c0 + 16), (sidx & 1) ? va : vb);
                    {
                        const size_t o = (size_t)m * ep.cout_pad + c0;
                        uint8_t *so = ostage + ((sidx ^ oswz) << 4);
                        if (plain) {
                            if (F8_DBG && (probe & 1)) *reinterpret_cast<uint4 *>(so) = make_uint4(v[0], v[5], v[10], v[15]);   // timing probe: WRONG results
                            else f8::epilogue16_plain_u8(v, sbias + c0, so, ep.shift0);
                        } else if (valid) {
                            f8::epilogue16_math(v, sbias + c0, kc, c, has_carry);
                            if (ep.carry_out) {
                                int32_t *dst = ep.carry_out + f8::carry_off((size_t)m, c0, ep.cout_pad);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    *reinterpret_cast<int4 *>(dst + k * 512) =
                                        make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                            }
                            if (ep.out0)
                                *reinterpret_cast<uint4 *>(so) = f8::requant_pack16(v, ep.shift0, ep.signed0);
                            if (ep.out1)
                                *reinterpret_cast<uint4 *>(ep.out1 + o) = f8::requant_pack16(v, ep.shift1, ep.signed1);
                        }
                    }
                }
            }
            if (F8_DBG && stats) { const long long n = clock64(); st_p[1] += n - ph0; ph0 = n; }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(buf));
            if (F8_DBG && stats) { const long long n = clock64(); st_p[2] += n - ph0; ph0 = n; }
            if (out_unit) {
                // rows past M and columns past cout_pad of the box are clipped by the tensor map
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(F8_DBG && (probe & 2))) {
                    tma_store_4d(&omap, cbase, t * TM + seg * BM + lg * 32, 0, 0,
                                 o_base + (uint32_t)(warp * OSTAGE));
                    tma_store_commit();
                }
            }
            if (F8_DBG && stats) { const long long n = clock64(); st_p[3] += n - ph0; ph0 = n; }
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
        if (out_unit && lane == 0) tma_store_wait_all();
        if (F8_DBG && stats && tid == 0)
            for (int k = 0; k < 4; ++k) stats[blockIdx.x * 16 + 8 + k] = st_p[k];
        if (F8_DBG && stats && tid == 0) {
            stats[blockIdx.x * 16 + 5] = clock64() - st_t0; stats[blockIdx.x * 16 + 6] = st_w; stats[blockIdx.x * 16 + 7] = st_st;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Returns F8_ERR_UNSUPPORTED when the layer is not one this kernel is for (the caller then
// takes the generic path).
template <int BN, bool A_SIGNED, bool PLAIN>
int launch_res_p(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s);
template <int BN, bool A_SIGNED>
int launch_res_t(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s) {
    return f8::epilogue_is_plain_u8(ep) ? launch_res_p<BN, A_SIGNED, true>(g, ep, s)
                                        : launch_res_p<BN, A_SIGNED, false>(g, ep, s);
}
template <int BN, bool A_SIGNED, bool PLAIN>
int launch_res_p(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s) {
    constexpr int MSEG = 256 / BN;
    constexpr int TM = MSEG * BM;
    constexpr int STAGE = MSEG * A_STAGE;
    const int w_bytes = g.ktiles * BK * BN;
    auto kern = conv1x1_res_kernel<BN, A_SIGNED, PLAIN>;
    static f8host::DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = f8host::device_once(once, &num_sms, [&]() -> int {
            F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            return F8_OK;
        });
        if (rc) return rc;
    }
    const int mtiles = (g.M + TM - 1) / TM;
    if (w_bytes > 64 * 1024 || mtiles < 2 * num_sms) return F8_ERR_UNSUPPORTED;
    const size_t fixed = (size_t)w_bytes + R_EPI_WARPS * 32 * 64 + (2 * R_SMAX + 5) * 8 + 8 + BN * 4 + 1024;
    int S = (int)((200 * 1024 - fixed) / STAGE);
    if (S > R_SMAX) S = R_SMAX;
    if (S < 2) return F8_ERR_UNSUPPORTED;
    size_t smem_bytes = fixed + (size_t)S * STAGE;
    if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;       // all 512 TMEM columns: one CTA per SM
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const uint64_t dims[4] = {(uint64_t)g.cin_pad, (uint64_t)g.M, 1u, 1u};
    const uint64_t row = (uint64_t)g.cin_pad, all = (uint64_t)g.cin_pad * (uint64_t)g.M;
    const uint64_t strides[3] = {row, all, all};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)BM, 1u, 1u};
    const int rc = f8host::encode_tmap_u8_4d(&tmap, g.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != F8_OK) return rc;
    CUtensorMap omap;
    memset(&omap, 0, sizeof(omap));
    if (ep.out0) {
        const uint64_t odims[4] = {(uint64_t)ep.cout_pad, (uint64_t)g.M, 1u, 1u};
        const uint64_t orow = (uint64_t)ep.cout_pad, oall = (uint64_t)ep.cout_pad * (uint64_t)g.M;
        const uint64_t ostrides[3] = {orow, oall, oall};
        const uint32_t obox[4] = {64u, 32u, 1u, 1u};
        const int orc = f8host::encode_tmap_u8_4d(&omap, ep.out0, odims, ostrides, obox, CU_TENSOR_MAP_SWIZZLE_64B);
        if (orc != F8_OK) return orc;
    }
    const int grid = mtiles < num_sms ? mtiles : num_sms;
    // 16 / 32 input channels: cp.async operand loader (chunks per pixel), see the kernel
    // (measured neutral: these layers are bound by the TMEM read rate of the epilogue, not by the loader;
    // kept behind F8_CPA=1)
    static const bool use_cpa = getenv("F8_CPA") != nullptr && atoi(getenv("F8_CPA")) != 0;
    const int cpa = (use_cpa && S >= 4 && g.ktiles == 1 && (g.cin_pad == 16 || g.cin_pad == 32)) ? g.cin_pad / 16 : 0;
    static const bool want_stats = f8host::debug_env("F8_STATS") != nullptr;
    static const int rprobe = f8host::debug_env("F8_RPROBE") ? atoi(f8host::debug_env("F8_RPROBE")) : 0;   // timing probes: WRONG results
    static long long *stats_dev = nullptr;
    if (want_stats) {
        if (!stats_dev) F8_CUDA(cudaMalloc(&stats_dev, 16 * 1024 * sizeof(long long)));
        F8_CUDA(cudaMemsetAsync(stats_dev, 0, 16 * 1024 * sizeof(long long), s));
    }
    f8host::note_kernel("conv1x1_res<BN=%d,%s>", BN, PLAIN ? "plain" : "generic");
    F8_CUDA(f8host::launch_pdl(kern, (unsigned)grid, (unsigned)R_THREADS, smem_bytes, s, g, ep, mtiles, S, tmap, omap,
                               want_stats ? stats_dev : (long long *)nullptr, rprobe, cpa));
    F8_CUDA(cudaGetLastError());
    if (want_stats) {
        static long long host[16 * 1024];
        F8_CUDA(cudaStreamSynchronize(s));
        F8_CUDA(cudaMemcpy(host, stats_dev, sizeof(host), cudaMemcpyDeviceToHost));
        double acc[16] = {0};
        for (int b = 0; b < grid; ++b)
            for (int k = 0; k < 16; ++k) acc[k] += (double)host[b * 16 + k] / grid;
        fprintf(stderr, "[f8 stats] conv1x1_res BN=%d cin=%d cout=%d M=%d tiles/cta=%.1f S=%d | producer total %.0f wait_empty %.0f | "
                "mma total %.0f wait_acc %.0f wait_full %.0f | epi(warp 0) total %.0f wait_full %.0f wait_store %.0f (cycles, mean per CTA)\n",
                BN, g.cin_pad, ep.cout_pad, g.M, (double)mtiles / grid, S, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]);
        fprintf(stderr, "[f8 stats]   epilogue warp 0 phases: store-wait+sync %.0f | tmem loads + math + staging %.0f | fence+sync+arrive %.0f | "
                "proxy fence+sync+store issue %.0f\n", acc[8], acc[9], acc[10], acc[11]);
    }
    return F8_OK;
}

template <int BN, bool TMA_A>
constexpr int smem_bytes_for() {
    return ring_for(BN, TMA_A) * ksteps_for(BN, TMA_A) * (A_STAGE + BN * BK) + (2 * ring_for(BN, TMA_A) + 4) * 8 + 16 +
           2 * BN * 4 + 1024;
}

template <int BN, bool A_SIGNED, bool SMALL_C, bool TMA_A, bool PLAIN = false>
int launch_t(const UGeom &g, const f8::Epilogue &ep, cudaStream_t s) {
    constexpr int smem_bytes = smem_bytes_for<BN, TMA_A>();
    auto kern = conv_umma_kernel<BN, A_SIGNED, SMALL_C, TMA_A, PLAIN>;
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
    static f8host::DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = f8host::device_once(once, &num_sms, [&]() -> int {
            F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            return F8_OK;
        });
        if (rc) return rc;
    }
    const int mtiles = (g.M + BM - 1) / BM;
    const int ntn = (ep.cout_pad + BN - 1) / BN;
    const long long total = (long long)mtiles * ntn;
    // 2 x BN TMEM columns and the smem ring per CTA: two co-resident CTAs for BN <= 128
    const int per_sm = (BN <= 128) ? 2 : 1;
    long long grid = (long long)num_sms * per_sm;
    if (grid > total) grid = total;
    f8host::note_kernel("conv_umma<BN=%d,%s%s,k%ds%d>", BN, SMALL_C ? "small_c" : (TMA_A ? "tma_a" : "gather"), PLAIN ? ",plain" : "",
                        g.kh, g.stride);
    F8_CUDA(f8host::launch_pdl(kern, (unsigned)grid, threads_for(BN, TMA_A), smem_bytes, s, g, ep, mtiles, ntn, tmap));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

template <int BN>
int launch_bn(const UGeom &g, const f8::Epilogue &ep, bool sgn, bool small_c, bool tma_a, cudaStream_t s) {
    if (small_c) return sgn ? launch_t<BN, true, true, false>(g, ep, s) : launch_t<BN, false, true, false>(g, ep, s);
    if (tma_a && f8::epilogue_is_plain_u8(ep))
        return sgn ? launch_t<BN, true, false, true, true>(g, ep, s) : launch_t<BN, false, false, true, true>(g, ep, s);
    if (tma_a) return sgn ? launch_t<BN, true, false, true>(g, ep, s) : launch_t<BN, false, false, true>(g, ep, s);
    return sgn ? launch_t<BN, true, false, false>(g, ep, s) : launch_t<BN, false, false, false>(g, ep, s);
}

}  // namespace

namespace f8host {

int launch_conv_umma(const f8_conv_args &a, cudaStream_t s) {
    if ((a.cin_pad != 4 && a.cin_pad % 16 != 0) || a.cout_pad % 16 != 0) return F8_ERR_UNSUPPORTED;
    if (a.cin_pad == 4 && a.kh == 3) {          // the MobileNet head: space-to-depth kernel (head3x3_umma.cu)
        const int rc = launch_head3x3s2(a, s);
        if (rc != F8_ERR_UNSUPPORTED) return rc;
    }
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
    if (const char *e = f8host::debug_env("F8_UPROBE")) g.probe = atoi(e);
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
    // few channels, many pixels: resident weights + multi-segment tiles (conv1x1_res_kernel)
    static const bool no_res = getenv("F8_NO_CONV1X1_RES") != nullptr;
    if (tma_a && !no_res && a.cout_pad <= 256 && !ep.out_f32) {
        int rc;
        if (a.cout_pad <= 64) rc = sgn ? launch_res_t<64, true>(g, ep, s) : launch_res_t<64, false>(g, ep, s);
        else if (a.cout_pad <= 128) rc = sgn ? launch_res_t<128, true>(g, ep, s) : launch_res_t<128, false>(g, ep, s);
        else rc = sgn ? launch_res_t<256, true>(g, ep, s) : launch_res_t<256, false>(g, ep, s);
        if (rc != F8_ERR_UNSUPPORTED) return rc;
    }
    if (a.cout_pad <= 64) return launch_bn<64>(g, ep, sgn, small_c, tma_a, s);
    if (a.cout_pad <= 128) return launch_bn<128>(g, ep, sgn, small_c, tma_a, s);
    return launch_bn<256>(g, ep, sgn, small_c, tma_a, s);
}

}  // namespace f8host
