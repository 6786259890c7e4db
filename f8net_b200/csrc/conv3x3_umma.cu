// conv3x3_umma.cu -- dense 3x3 / pad 1 int8 convolution (stride 1, and stride 2 through four
// input parity planes) on tcgen05 with the input
// patch RESIDENT in shared memory: every input byte is fetched once per output tile and the
// nine filter taps are nine shifted views of the same patch, expressed purely through the
// start address of the tcgen05 shared-memory descriptor.
//
// Replaces, per launch: the 3x3 int nn.Conv2d of BasicBlock / Bottleneck bodies built by
// int_conv() (/root/reference/models/fix_quant_ops.py:680-714) and the tensor-op chain around
// it in IntBlock.forward (/root/reference/models/fix_resnet.py:28-77).
//
// Padded linear pixel space.  Stack the images of the batch into one tall image with ONE zero
// row above every image and ONE zero column before every row; with pitch PW = W+1 the zero
// column also serves as the right halo of the previous row, the zero row as the bottom halo of
// the previous image:
//     input  slot   pi(img, y, x) = (img*(H+1) + y + 1)*PW + x + 1
//     output index  m (img, y, x) = (img*(H+1) + y    )*PW + x
// so the input needed by output m for tap (r, s) is slot  m + r*PW + s  -- one constant offset
// per tap for ALL pixels.  A tile is any 512 consecutive output indices (four M=128 MMA
// segments); its patch is the slot range [m0, m0 + 512 + 2*PW + 2).  Output indices that fall
// on the zero column / zero row are computed and dropped (W/(W+1) * H/(H+1) of the MMA rows are
// real pixels: 96 % at 56x56, 77 % at 7x7).
//
// Shared-memory operand layout: the patch of one 64-channel group arrives by TMA as [slot][64 B] in
// the 64-byte-swizzled K-major layout (SBO = 8 slots x 64 B); the swizzle XOR acts on absolute
// address bits, so a tap shift of d slots is still "start address += 64*d" of the descriptor.
// K order: for each 64-channel group, for each tap: two K=32 MMAs per M segment.  The weight ring
// stage is one filter row (3 taps x 64 channels x BN columns, 12 bulk copies), shared by every
// M segment of the tile.
//
// Persistent CTA, 1 per SM, 512 TMEM columns = 2 accumulator sets x MB segments x BN columns
// (BN = 64: 4 segments, BN = 128: 2):
//   warps 0-15 epilogue (TMEM read rate and the exact integer requantisation bound it: 16 warps)
//   | warp 16 patch loader (TMA, one lane per box) | warps 17, 19: MMA issuers (one elected lane each),
//   alternating weight stages | warp 18 lane 0 weight-stage loader (cp.async.bulk).
// Two MMA warps because the tensor pipe does not run ahead of the issuing thread (it queues ~4 MMAs,
// tools/probes/mma4_probe.cu) and a stage boundary costs the issuing warp 500-700 cycles of plain
// instruction latency (commit, ring counters, barrier wait, descriptors, elect / reconvergence;
// profiles/r02_mma_loop_experiments.md): with one issuer the pipe idles for a third of every stage.
// The two warps take the weight stages in turn and pass the turn through a named barrier, so each
// one's boundary runs under the other's burst.
// Variants: STRIDE = 2 (four parity-plane tensor maps), DW (depthwise: diagonal 64 x 64 weight
// blocks, two N = 32 MMAs per tap).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tma_common.cuh"

namespace {

using namespace f8u;

// Tile = MB segments of 128 output indices x BN output channels, MB * BN = 256 so that two
// accumulator sets fill the 512 TMEM columns.  BN = 128 (MB = 2) balances the tensor pipe
// against operand fetch from shared memory (per K=32 MMA: 4 KB of A + 4 KB of B in 64 tensor
// cycles); BN = 64 (MB = 4) serves the 64-channel layers.
__host__ __device__ constexpr int mb_for(int bn) { return 256 / bn; }
// Ring depths.
// weight ring: one stage = one filter row (3 taps x 64 channels x BN columns), so the MMA warp
// synchronises three times per 64-channel group instead of nine
__host__ __device__ constexpr int sb_for(int bn, bool plain) { (void)plain; return bn == 64 ? 4 : 3; }
__host__ __device__ constexpr int sa_for(int stride) { return stride == 2 ? 2 : 3; }   // patch ring (a stride-2 patch is 4 parity planes)
// a patch stage is signalled once A_LAG younger stages are issued (never the whole ring)
__host__ __device__ constexpr int a_lag_for(int stride) { return stride == 2 ? 0 : 1; }
// epilogue warps: warp w reads TMEM lane group w % 4 and owns one 64-column unit (w / 4): the
// exact integer requantisation is instruction-bound, hence 16 warps
__host__ __device__ constexpr int epi_warps_for(bool plain) { (void)plain; return 16; }
constexpr int LOADERS = 128;

// stride 1: m[0] = the NHWC activation; stride 2: m[p] = its parity plane p = (row parity, column
// parity) as a strided view (every second pixel of every second row)
struct TMaps {
    CUtensorMap m[4];
};

struct PGeom {
    const uint8_t *in;
    const uint8_t *wpack;   // [K_pad/16][wrows][16], k = (r*3+s)*C + c
    const uint8_t *wstage;  // dense: stage-major image [N tile][group][filter row][column][4 chunks][BN][16], or nullptr
    int wrows;
    int N, H, W, C;         // H x W = OUTPUT size per image, C = cin_pad (multiple of 64)
    int Hin, Win;           // input size per image (= H, W for stride 1; 2H, 2W for stride 2)
    int plane_slots;        // slots of one parity plane (stride 1: the only plane)
    int PW;                 // W + 1
    uint32_t mPW, mHP, mBY; // floor(2^32 / d) + 1 for d = PW, H + 1, BY: n / d == __umulhi(n, m) while n * d < 2^32
    int tm;                 // 128 * MB
    int slots;              // planes * plane_slots
    int slots_pad;          // slots rounded up to 8
    int n_super;            // tiles along the padded linear space
    int ntiles_n;           // cout_pad / BN (rounded up)
    long long *stats;       // debug (F8_STATS=1): per-CTA wait-cycle counters, else nullptr
    int probe;              // debug (F8_PROBE): timing probes, WRONG results
    int lx;                 // log2 of the 8-slot items per padded row (PW <= 8 << lx)
    int BY;                 // TMA path: padded rows per box, a divisor of H + 1 (boxes never straddle images)
    int sa;                 // patch ring depth (<= sa_for(STRIDE))
    int sb;                 // weight ring depth (<= sb_for(BN))
    int pps;                // slots of one plane's box-aligned patch region in a stage (stride 2)
    int dw;                 // depthwise mode: output tile n reads only input channel group n, through a
                            // block-diagonal 64 x 64 weight image per group (wpack = [group][36][64][16])
    int box_slots;          // TMA path: BY * PW
};

#define F8_TIMED_WAIT(acc, stmt)                 \
    do {                                         \
        if (F8_DBG && g.stats) {                 \
            const long long _t0 = clock64();     \
            stmt;                                \
            acc += clock64() - _t0;              \
        } else {                                 \
            stmt;                                \
        }                                        \
    } while (0)

// MC = 2 | 4 (dense stride 1, BN = 128): the kernel runs as clusters of MC CTAs that work on MC different M
// tiles of the SAME N tile in lock step and share one weight stream: each CTA fetches 1/MC of every weight
// stage and multicasts it into all MC shared memories (cp.async.bulk ... .multicast::cluster), each issues its
// own cta_group::1 MMAs, and every stage release is committed to all CTAs' "empty" barriers.  The L2 -> SM
// weight traffic -- which bounds these layers (590 KB per tile at 512 channels) -- drops to 1/MC.
template <int BN, bool A_SIGNED, bool PLAIN_U8, int STRIDE, bool DW, int MC>
__global__ void __launch_bounds__((epi_warps_for(PLAIN_U8) + 4) * 32, 1)
conv3x3_umma_kernel(const PGeom g, const f8::Epilogue ep, const __grid_constant__ TMaps tmaps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // stride 1: the patch arrives by TMA in the 64-byte-swizzled K-major layout [slot][64 B]
    // (stage bases 1024-byte aligned); stride 2: cp.async into [16-byte chunk][slot][16 B]
    constexpr bool TMA = true;
    constexpr int PLANES = STRIDE == 2 ? 4 : 1;
    uint8_t *smem = smem_raw + ((1024u - (f8::smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int EPI_WARPS = epi_warps_for(PLAIN_U8);
    constexpr int EPI_THREADS = EPI_WARPS * 32;
    constexpr int LOADER_WARP0 = EPI_WARPS;
    constexpr int MMA_WARP = EPI_WARPS + 1;       // one TMA warp, one MMA warp, one weight loader:
    constexpr int WLOAD_WARP = EPI_WARPS + 2;
    constexpr int MMA_WARP2 = EPI_WARPS + 3;      // 20 warps: 96 registers each at launch (generic epilogue: setmaxnreg below)
    constexpr int MB = mb_for(BN);
    constexpr int TM = 128 * MB;
    constexpr int SB_MAX = sb_for(BN, PLAIN_U8);
    const int SB = g.sb;
    constexpr int SA_MAX = sa_for(STRIDE);
    const int SA = g.sa;
    constexpr int A_LAG = a_lag_for(STRIDE);
    constexpr int BROWS = DW ? 64 : BN;                   // weight rows of a tile (depthwise: one 64-channel group)
    constexpr int B_TILE = BROWS * 64;                    // one tap
    constexpr int B_STAGE = 3 * B_TILE;                   // one filter row
    constexpr int CW = BN / (EPI_WARPS / 4);               // columns per epilogue warp slice (plain path)
    const int a_stage = g.slots_pad * 64;                 // bytes of one patch stage
    const uint32_t lbo_a = (uint32_t)g.slots_pad * 16;
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t sb_base = smem_base + SA * a_stage;
    const uint32_t bar_base = sb_base + SB * B_STAGE;
    // a_full[SA] a_empty[SA] b_full[SB] b_empty[SB] acc_full[2] acc_empty[2]
    auto a_full = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto a_empty = [&](int s) { return bar_base + (uint32_t)(SA_MAX + s) * 8; };
    auto b_full = [&](int s) { return bar_base + (uint32_t)(2 * SA_MAX + s) * 8; };
    auto b_empty = [&](int s) { return bar_base + (uint32_t)(2 * SA_MAX + SB_MAX + s) * 8; };
    auto acc_full = [&](int b) { return bar_base + (uint32_t)(2 * SA_MAX + 2 * SB_MAX + b) * 8; };
    auto acc_empty = [&](int b) { return bar_base + (uint32_t)(2 * SA_MAX + 2 * SB_MAX + 2 + b) * 8; };
    constexpr int NBARS = 2 * SA_MAX + 2 * SB_MAX + 4;
    uint8_t *after = smem + SA * a_stage + SB * B_STAGE + ((NBARS * 8 + 15) & ~15);   // bias copy: 16-byte aligned
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(after);
    int32_t *sbias = reinterpret_cast<int32_t *>(after + 16);

    const long long t_entry = clock64();
    if (F8_DBG && g.stats && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.stats[blockIdx.x * 16 + 2] = (long long)gt;
    }
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // MC: the cluster is the scheduling unit; cluster item `it` = (group of CLS super-tiles, N tile), CTA rank r takes
    // super-tile CLS * group + r (a super-tile past the batch is all padding: zero-filled by the TMA, dropped by the epilogue)
    constexpr int CLS = MC ? MC : 1;
    const int total_items = ((g.n_super + CLS - 1) / CLS) * g.ntiles_n;
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    constexpr uint16_t CTA_MASK = (uint16_t)((1u << CLS) - 1u);
    const int bid = (int)blockIdx.x / CLS;
    const int nb = (int)gridDim.x / CLS;
    const int n_loc = g.N;
    constexpr int pix_off = 0;
    auto super_of = [&](int it_) { const int sp = it_ / g.ntiles_n; return MC ? sp * CLS + (int)rank : sp; };
    // K loop: dense = every 64-channel group of the input; depthwise = the BN / 64 channel groups
    // of the output tile itself (each through its own diagonal weight block, on its own columns)
    constexpr int GPT = BN / 64;
    const int ngroups = (ep.cout_pad + 63) >> 6;
    auto tile_group0 = [&](int it) { return DW ? (it - (it / g.ntiles_n) * g.ntiles_n) * GPT : 0; };
    auto tile_ncg = [&](int it) {
        if (!DW) return g.C >> 6;
        const int left = ngroups - tile_group0(it);
        return left < GPT ? left : GPT;
    };
    const int HP = g.H + 1;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            // a patch / an accumulator set is released (handed on) by BOTH MMA warps: tcgen05.commit covers the
            // committing thread's own MMAs only, and each warp issues at least one stage of every channel group
            for (int s = 0; s < SA; ++s) { mbar_init(a_full(s), TMA ? 1 : LOADERS); mbar_init(a_empty(s), 2); }
            for (int p = 0; p < PLANES; ++p) tma_prefetch_desc(&tmaps.m[p]);
            // MC: a weight stage is overwritten in BOTH CTAs, so both CTAs' MMAs must have released it
            for (int s = 0; s < SB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), CLS); }
            // one arrival per epilogue warp (not 512 serialised atomics)
            for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 2); mbar_init(acc_empty(b), EPI_WARPS); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();         // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // The next layer's launch may begin its own prologue as this grid's CTAs retire; everything
    // that touches the previous layer's output (patch TMA, residual carry) waits for that layer
    // here.  The weight loader does not: weights and biases are plan constants, so it fills
    // its ring while the previous launch drains.
    f8::pdl_trigger();
    if (warp != WLOAD_WARP) f8::pdl_wait();

    // Register re-balancing (generic epilogue only): 20 warps x 96 registers is the launch allocation and the pool
    // setmaxnreg works in (it is per CTA: an increase can only take what a decrease has released).  The four
    // producer / issuer warps (one warpgroup, one setmaxnreg for all four) drop to 64 and release 128 x 32 registers,
    // the sixteen epilogue warps (four warpgroups) take 512 x 8 = the same amount and run with 104 -- room for a
    // residual-carry ring two 16-column steps deep without spilling.  Each setmaxnreg dominates the code that runs
    // under it (ptxas budgets a region by the value that reaches it).
    if (warp >= EPI_WARPS) {
    if (!PLAIN_U8) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (TMA && warp >= LOADER_WARP0 && warp < MMA_WARP) {
        // =========================== patch loader (TMA) ===========================
        // One warp; lane l issues the tensor load of box l of the stage: a box is one padded row
        // (or one whole padded image when that is at most 64 slots) of 64 channels, start
        // coordinate x = -1 / y = -1: the zero column, the zero row and everything outside the
        // batch are the TMA's out-of-bounds zero fill.
        if (warp == LOADER_WARP0) {
            int slot = 0, phase = 0;
            long long w_empty = 0;
            const long long t_begin = clock64();
            const int PW = g.PW, BY = g.BY, BS = g.box_slots;
            for (int it = bid; it < total_items; it += nb) {
                const int st = super_of(it);
                const int cg0 = tile_group0(it), ncg = tile_ncg(it);
                const int pi0 = st * TM;
                const int Y0 = (int)__umulhi((uint32_t)pi0, g.mPW);
                const int Ylast = (int)__umulhi((uint32_t)(pi0 + g.plane_slots - 1), g.mPW);
                const int Yb0 = BY == 1 ? Y0 : (int)__umulhi((uint32_t)Y0, g.mBY) * BY;
                const int nbox = BY == 1 ? Ylast - Yb0 + 1 : (int)__umulhi((uint32_t)(Ylast - Yb0), g.mBY) + 1;
                const int Yb = Yb0 + lane * BY;
                const int img = (int)__umulhi((uint32_t)Yb, g.mHP);
                const int yy = Yb - img * HP;
                for (int cg = 0; cg < ncg; ++cg) {
                    F8_TIMED_WAIT(w_empty, mbar_wait(a_empty(slot), phase ^ 1));
                    const bool skip_tma = F8_DBG && (g.probe & 256) && it != bid;      // probe: stale patch (WRONG results)
                    if (lane == 0) mbar_arrive_expect_tx(a_full(slot), skip_tma ? 0u : (uint32_t)(PLANES * nbox * BS * 64));
                    __syncwarp();
                    if (skip_tma) {
                    } else if (PLANES == 1) {
                        if (lane < nbox)
                            tma_load_4d(smem_base + slot * a_stage + lane * BS * 64, &tmaps.m[0], (cg0 + cg) * 64, -1,
                                        yy - 1, img, a_full(slot));
                    } else {
                        // box b of plane p: the same padded rows of every parity plane
                        for (int b = lane; b < PLANES * nbox; b += 32) {
                            int p = 0, bb = b;
                            while (bb >= nbox) { bb -= nbox; ++p; }
                            const int Ybb = Yb0 + bb * BY;
                            const int im = (int)__umulhi((uint32_t)Ybb, g.mHP);
                            const int y0 = Ybb - im * HP;
                            tma_load_4d(smem_base + slot * a_stage + (p * g.pps + bb * BS) * 64, &tmaps.m[p],
                                        (cg0 + cg) * 64, -1, y0 - 1, im, a_full(slot));
                        }
                    }
                    if (++slot == SA) { slot = 0; phase ^= 1; }
                }
            }
            if (F8_DBG && g.stats && lane == 0) {
                g.stats[blockIdx.x * 16 + 0] = clock64() - t_begin;
                g.stats[blockIdx.x * 16 + 1] = w_empty;
            }
        }
    } else if (warp == WLOAD_WARP) {
        // =========================== weight-tile loader ===========================
        // The whole warp runs the loop; lane l issues bulk copy l of the stage (12 copies: 3 taps x 4
        // sixteen-byte K chunks, each BROWS rows).  One lane issuing all twelve was THE bound of this kernel:
        // under the tensor core's operand traffic every cp.async.bulk issue takes ~130 cycles, so a stage
        // took the loader ~1600 cycles -- longer than its 1152 cycles of MMAs (profiles/r02_mma_loop_experiments.md).
        {
            int slot = 0, phase = 0;
            long long w_bempty = 0;
            const long long t_begin = clock64();
            const int fs = lane >> 2, j = lane & 3;             // this lane's copy: tap column fs, K chunk j
            for (int it = bid; it < total_items; it += nb) {
                const int st = it / g.ntiles_n;
                const int n0 = DW ? 0 : (it - st * g.ntiles_n) * BN;
                const int Ck = DW ? 64 : g.C;
                const int ncg = tile_ncg(it);
                // depthwise: 64 weight rows per group image, dense: BN rows of the shared image
                constexpr uint32_t ROWS_B = (uint32_t)BROWS;
                for (int cg = 0; cg < ncg; ++cg) {
                    const uint8_t *wsrc = g.wpack + (DW ? (size_t)(tile_group0(it) + cg) * (36 * 64 * 16) : 0);
                    const int kcg = DW ? 0 : cg;
                    for (int fr = 0; fr < 3; ++fr) {
                        F8_TIMED_WAIT(w_bempty, mbar_wait(b_empty(slot), phase ^ 1));
                        const uint32_t sb = sb_base + slot * B_STAGE;
                        if (lane == 0) mbar_arrive_expect_tx(b_full(slot), 12u * ROWS_B * 16u);
                        __syncwarp();
                        if (DW) {
                            // depthwise: the group's diagonal weight image is chunk-major with 64 rows, so the twelve
                            // chunks of a filter row are 12 KB of CONTIGUOUS memory in the order shared memory wants:
                            // one bulk copy (the copy unit retires ~one copy per 100 cycles whatever lane issues it)
                            if (lane == 0) bulk_g2s(sb, wsrc + (size_t)(fr * 3) * 4 * 64 * 16, 12u * ROWS_B * 16u, b_full(slot));
                        } else if (g.wstage) {
                            // stage-major image: the stage is 12 * BN * 16 contiguous bytes in shared-memory order:
                            // ONE bulk copy (pairs: each CTA copies its half into both)
                            const uint8_t *ssrc = g.wstage + ((size_t)((n0 / BN) * (g.C >> 6) + cg) * 3 + fr) * (size_t)(12 * BN * 16);
                            if (lane == 0) {
                                if (!MC) bulk_g2s(sb, ssrc, 12u * ROWS_B * 16u, b_full(slot));
                                else bulk_g2s_mc(sb + rank * (12u / CLS) * ROWS_B * 16u, ssrc + rank * (12u / CLS) * ROWS_B * 16u,
                                                 (12u / CLS) * ROWS_B * 16u, b_full(slot), CTA_MASK);
                            }
                        } else if (lane < 12) {
                            const size_t kc = (size_t)((fr * 3 + fs) * Ck + kcg * 64) >> 4;
                            if (!MC)
                                bulk_g2s(sb + fs * B_TILE + j * (BROWS * 16), wsrc + ((kc + j) * g.wrows + n0) * 16, ROWS_B * 16u,
                                         b_full(slot));
                            else if ((uint32_t)(lane % CLS) == rank)    // my share of the stage, into every CTA of the cluster
                                bulk_g2s_mc(sb + fs * B_TILE + j * (BROWS * 16), wsrc + ((kc + j) * g.wrows + n0) * 16,
                                            ROWS_B * 16u, b_full(slot), CTA_MASK);
                        }
                        if (++slot == SB) { slot = 0; phase ^= 1; }
                    }
                }
            }
            if (F8_DBG && g.stats && lane == 0) {
                g.stats[blockIdx.x * 16 + 3] = clock64() - t_begin;
                g.stats[blockIdx.x * 16 + 4] = w_bempty;
            }
        }
    } else if (warp == MMA_WARP || warp == MMA_WARP2) {
        // =========================== MMA issuers ==================================
        // Two warps, each running the whole loop (uniform control flow, one elected lane issues); warp p owns
        // the weight stages with (running stage number) % 2 == p.  A stage may only be issued after the
        // previous one HAS BEEN ISSUED (accumulation order inside a tile): the turn passes through two named
        // barriers (bar.arrive after a warp's stage, bar.sync before the other's) -- no shared-memory access.
        // Everything else of a stage boundary (commit, ring counters, the mbarrier waits for the stage after
        // next, descriptors) then runs while the OTHER warp's burst keeps the tensor pipe busy.
        // Descriptor high words are loop constants, low words advance by 32-bit adds.
        const int me = warp == MMA_WARP ? 0 : 1;
        constexpr uint32_t idesc = instr_desc(A_SIGNED, BN);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);          // SBO = 128 B, version 1
        // TMA patch: SWIZZLE_64B (layout type 4), SBO = 8 rows x 64 B, LBO unused
        constexpr uint32_t desc_hi_a = (512u >> 4) | (1u << 14) | (4u << 29);
        constexpr uint32_t SLOT16 = 4u;                                 // one slot in descriptor units of 16 B
        constexpr uint32_t a_lbo_field = 1u << 16;
        constexpr uint32_t b_lbo_field = ((uint32_t)(BROWS * 16) >> 4) << 16;
        int aslot = 0, aphase = 0, bslot = 0, bphase = 0, buf = 0, acc_phase = 0;
        // start-address offsets of the taps in descriptor units (16 B = one slot)
        uint32_t tap_row[3], tap_col[2];
        if (STRIDE == 2) {
            // row fr: plane bit (fr != 1) * 2, plus one padded row for fr > 0; column fs: plane bit
            // (fs != 1), plus one padded column for fs > 0
            tap_row[0] = SLOT16 * 2u * g.pps; tap_row[1] = SLOT16 * (uint32_t)g.PW; tap_row[2] = SLOT16 * (2u * g.pps + g.PW);
            tap_col[0] = SLOT16 * (uint32_t)g.pps; tap_col[1] = SLOT16 * ((uint32_t)g.pps + 1u);
        } else {
            tap_row[0] = 0u; tap_row[1] = SLOT16 * (uint32_t)g.PW; tap_row[2] = SLOT16 * 2u * g.PW;
            tap_col[0] = tap_col[1] = 0u;
        }
        long long w_acc = 0, w_a = 0, w_b = 0;
        const long long t_begin = clock64();
        long long t_first_a = 0;
        int turn = 0;                     // whose stage is next (running stage number mod 2)
        bool started = false;             // this warp has issued a stage before (the other warp's turn signal exists)
        for (int it = bid; it < total_items; it += nb) {
            bool acc_ready = false;
            const uint32_t tacc = tmem_base + (uint32_t)(buf * MB * BN);
            int tile_soff = 0;                           // first slot of the tile inside its box-aligned patch
            const int ncg = tile_ncg(it);
            {
                const int pi0 = super_of(it) * TM;
                const int Y0 = (int)__umulhi((uint32_t)pi0, g.mPW);
                const int Yb0 = g.BY == 1 ? Y0 : (int)__umulhi((uint32_t)Y0, g.mBY) * g.BY;
                tile_soff = pi0 - Yb0 * g.PW;
            }
            for (int cg = 0; cg < ncg; ++cg) {
                bool a_ready = false;
                const uint32_t sa = smem_base + aslot * a_stage + (uint32_t)tile_soff * 64u;
                const uint32_t first = (uint32_t)(cg != 0);
                // depthwise: does this channel group have channels 32..63?
                const int dw_halves = (DW && ep.cout_pad - (tile_group0(it) + cg) * 64 <= 32) ? 1 : 2;
                // this warp's last stage of the channel group / of the tile (stages alternate, three per group)
                const int my_last_fr = ((turn + 2) & 1) == me ? 2 : 1;
#pragma unroll
                for (int fr = 0; fr < 3; ++fr) {
                    if (turn == me) {
                        // waits of MY stage: they ran ahead while the other warp was issuing
                        if (!acc_ready) { F8_TIMED_WAIT(w_acc, mbar_wait(acc_empty(buf), acc_phase ^ 1)); acc_ready = true; }
                        if (!a_ready) { F8_TIMED_WAIT(w_a, mbar_wait(a_full(aslot), aphase)); a_ready = true; }
                        F8_TIMED_WAIT(w_b, mbar_wait(b_full(bslot), bphase));
                        if (F8_DBG && g.stats && t_first_a == 0) t_first_a = clock64();
                        const uint32_t sb = sb_base + bslot * B_STAGE;
                        const uint32_t a_row = (((sa & 0x3ffffu) >> 4) | a_lbo_field) + tap_row[fr];
                        const uint32_t b_row = ((sb & 0x3ffffu) >> 4) | b_lbo_field;
                        // my turn: the other warp has issued the previous stage (none before the very first one)
                        if (started || me == 1) turn_wait(me ^ 1);
                        started = true;
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int fs = 0; fs < 3; ++fs) {
                                // slot offset of tap (fr, fs): stride 1: fr*PW + fs; stride 2: parity plane
                                // ((fr != 1), (fs != 1)) and a one-row / one-column step for fr > 0 / fs > 0
                                const uint32_t a_lo0 = a_row + (STRIDE == 2 ? (fs == 0 ? tap_col[0] : (fs == 1 ? SLOT16 : tap_col[1]))
                                                                            : SLOT16 * (uint32_t)fs);
                                const uint32_t b_lo0 = b_row + (uint32_t)(fs * (B_TILE >> 4));
                                if constexpr (DW) {
                                    // diagonal 64 x 64 block: K half h (channels 32h..32h+31) only reaches output
                                    // columns 32h..32h+31 -> two N = 32 MMAs on disjoint columns (one if the
                                    // group has no upper half)
                                    constexpr uint32_t idesc32 = instr_desc(A_SIGNED, 32);
#pragma unroll
                                    for (int i = 0; i < MB; ++i) {
#pragma unroll
                                        for (int h = 0; h < 2; ++h)
                                            if (h < dw_halves)
                                                umma_i8_lohi(tacc + (uint32_t)(i * BN + 64 * cg + 32 * h),
                                                             a_lo0 + (uint32_t)(i * 512 + h * 2), desc_hi_a,
                                                             b_lo0 + (uint32_t)h * (((2 * BROWS * 16) >> 4) + 32u), desc_hi, idesc32,
                                                             (fs | fr) ? 1u : 0u);
                                    }
                                } else {
#pragma unroll
                                    for (int i = 0; i < MB; ++i) {
#pragma unroll
                                        for (int h = 0; h < 2; ++h)
                                            umma_i8_lohi(tacc + (uint32_t)(i * BN), a_lo0 + (uint32_t)(i * 512 + h * 2), desc_hi_a,
                                                         b_lo0 + (uint32_t)h * ((2 * BROWS * 16) >> 4), desc_hi, idesc,
                                                         (h | fs | fr) ? 1u : first);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        // the next stage (the other warp's) may be issued; none follows the CTA's very last stage
                        if (!(it + nb >= total_items && cg == ncg - 1 && fr == 2)) turn_pass(me);
                        if (elect_one()) {
                            if (MC) umma_commit_mc(b_empty(bslot), CTA_MASK);     // every CTA's loader writes this slot next
                            else umma_commit(b_empty(bslot));
                            if (fr == my_last_fr) {
                                umma_commit(a_empty(aslot));                     // my MMAs on this patch
                                if (cg == ncg - 1) umma_commit(acc_full(buf));   // my MMAs on this tile
                            }
                        }
                        __syncwarp();
                    }
                    turn ^= 1;
                    if (++bslot == SB) { bslot = 0; bphase ^= 1; }
                }
                if (++aslot == SA) { aslot = 0; aphase ^= 1; }
            }
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
        if (F8_DBG && g.stats && lane == 0 && me == 0) {
            g.stats[blockIdx.x * 16 + 5] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 6] = w_acc;
            g.stats[blockIdx.x * 16 + 7] = w_a;
            g.stats[blockIdx.x * 16 + 8] = w_b;
            g.stats[blockIdx.x * 16 + 15] = ((t_begin - t_entry) << 32) | ((t_first_a - t_entry) & 0xffffffffll);
        }
    }
    } else {
        if (!PLAIN_U8) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // =========================== epilogue (warps 0-15) ========================
        const int lg = warp & 3;                       // TMEM lane group of this warp
        const int cw0 = (warp >> 2) * CW;              // this warp's column slice of the tile
        const int row = lg * 32 + lane;
        int buf = 0, acc_phase = 0;
        int bias_n0[2] = {-1, -1};                     // N tile whose bias each buffer's shared copy holds
        bool epi_primed = false;
        int4 cr0[4] = {}, cr1[4] = {};   // residual carry of the even / odd 16-column steps, requested two steps ahead
        long long w_full = 0, t_issue = 0, t_wait = 0, t_math = 0, t_store = 0;
        const long long t_begin = clock64();
        for (int it = bid; it < total_items; it += nb) {
            const int st = super_of(it);
            const int n0 = (it - (it / g.ntiles_n) * g.ntiles_n) * BN;
            int ncols = ep.cout_pad - n0;
            if (ncols > BN) ncols = BN;
            int32_t *bias_s = sbias + buf * BN;
            if (bias_n0[buf] != n0) {            // warp-uniform: reload only when this buffer's N tile changes
                // every warp is done with the tile that last read this copy before it is rewritten
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (tid < ncols) {
                    int32_t b = __ldg(ep.bias + n0 + tid);
                    if (PLAIN_U8) b = (int32_t)((uint32_t)b + (1u << (ep.shift0 - 1)));   // bias + half
                    bias_s[tid] = b;
                }
                bias_n0[buf] = n0;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }
            if (PLAIN_U8) {
                F8_TIMED_WAIT(w_full, mbar_wait(acc_full(buf), acc_phase));
                tc_fence_after();
                if (cw0 < ncols && !(F8_DBG && (g.probe & 16))) {
#pragma unroll
                    for (int i = 0; i < MB; ++i) {
                        const int m = st * TM + i * 128 + row;
                        const int Yo = (int)__umulhi((uint32_t)m, g.mPW);
                        const int xo = m - Yo * g.PW;
                        const int img = (int)__umulhi((uint32_t)Yo, g.mHP);
                        const int y = Yo - img * HP;
                        const bool valid = xo < g.W && y < g.H && img < n_loc;
                        const size_t opix = ((size_t)(img * g.H + y) * g.W + xo) + (size_t)pix_off;
                        const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) +
                                              (uint32_t)((buf * MB + i) * BN);
#pragma unroll
                        for (int c0 = cw0; c0 < cw0 + CW; c0 += 16) {
                            if (c0 < ncols) {
                                int32_t v[16];
                                tmem_ld16(trow + (uint32_t)c0, v);
                                tmem_ld_wait();
                                if (valid)
                                    f8::epilogue16_plain_u8(v, bias_s + c0,
                                                            ep.out0 + opix * ep.cout_pad + n0 + c0, ep.shift0);
                            }
                        }
                    }
                }
            } else {
                // ---- generic epilogue: residual carries, int32 carry out, dual / signed outputs ----
                // Warp (lg, u) owns rows [32*lg, 32*lg+32) of unit u (one M segment x 64 columns) and
                // walks it in four steps of 16 columns.  int32 carries live in the pixel-interleaved
                // layout of f8_common.cuh: the warp's 32 (nearly) consecutive pixels make every
                // 16-byte carry access a contiguous 512-byte run.  The carries of steps q+1 and q+2 (running
                // over into the next tile's first steps) are in flight in registers while step q is computed.
                constexpr int UPS = BN / 64;                 // units per segment
                const bool has_carry = ep.carry_in != nullptr;
                const f8::EpiConst kc = f8::epi_const(ep, has_carry);
                const int u = warp >> 2;
                const int seg = u / UPS;
                const int cbase = (u - seg * UPS) * 64;      // first column of the unit inside the tile
                auto unit_pixel = [&](int st_) -> int {      // output pixel of this thread's row, -1 = dropped
                    const int m = st_ * TM + seg * 128 + row;
                    const int Yo = (int)__umulhi((uint32_t)m, g.mPW);
                    const int xo = m - Yo * g.PW;
                    const int img = (int)__umulhi((uint32_t)Yo, g.mHP);
                    const int y = Yo - img * HP;
                    return (xo < g.W && y < g.H && img < n_loc) ? (img * g.H + y) * g.W + xo + pix_off : -1;
                };
                auto load_carry = [&](int pix_, int col, int4 (&c)[4]) {
                    if (has_carry && pix_ >= 0 && col < ep.cout_pad) {
                        const int32_t *src = ep.carry_in + f8::carry_off((size_t)pix_, col, ep.cout_pad);
#pragma unroll
                        for (int k = 0; k < 4; ++k) c[k] = __ldg(reinterpret_cast<const int4 *>(src + k * 512));
                    }
                };
                const int pix = unit_pixel(st);
                if (!epi_primed) {                // very first steps of this CTA
                    load_carry(pix, n0 + cbase, cr0);
                    load_carry(pix, n0 + cbase + 16, cr1);
                    epi_primed = true;
                }
                // the next tile of this CTA: its first two steps are requested during this tile's last two
                const int it2 = it + nb;
                const int pix2 = it2 < total_items ? unit_pixel(super_of(it2)) : -1;
                const int col2 = (it2 - (it2 / g.ntiles_n) * g.ntiles_n) * BN + cbase;
                F8_TIMED_WAIT(w_full, mbar_wait(acc_full(buf), acc_phase));
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // step q's carry was requested two steps ago (a step is ~2 us under load, a DRAM round trip under
                    // load not much less: one step of lead left a quarter of the epilogue's time waiting for it)
                    int4 c[4];
                    int4 (&ring)[4] = (q & 1) ? cr1 : cr0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) c[k] = ring[k];
                    if (q < 2) load_carry(pix, n0 + cbase + 16 * (q + 2), ring);
                    else load_carry(pix2, col2 + 16 * (q - 2), ring);
                    const int col = n0 + cbase + 16 * q;
                    if (col < ep.cout_pad) {                               // warp-uniform
                        int32_t v[16];
                        tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) +
                                      (uint32_t)(buf * MB * BN + seg * BN + cbase + 16 * q), v);
                        tmem_ld_wait();
                        if (pix >= 0) {
                            f8::epilogue16_math(v, bias_s + cbase + 16 * q, kc, c, has_carry);
                            if (ep.carry_out) {
                                int32_t *dst = ep.carry_out + f8::carry_off((size_t)pix, col, ep.cout_pad);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    *reinterpret_cast<int4 *>(dst + k * 512) =
                                        make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                            }
                            const size_t o = (size_t)pix * ep.cout_pad + col;
                            if (ep.out0)
                                *reinterpret_cast<uint4 *>(ep.out0 + o) = f8::requant_pack16(v, ep.shift0, ep.signed0);
                            if (ep.out1)
                                *reinterpret_cast<uint4 *>(ep.out1 + o) = f8::requant_pack16(v, ep.shift1, ep.signed1);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf));      // this warp's accumulator columns are drained
            }
            if (PLAIN_U8) {
                // every accumulator column this warp owns is in registers (or consumed)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf));
            }
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
        if (F8_DBG && g.stats && tid == 0) {
            g.stats[blockIdx.x * 16 + 9] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 10] = w_full;
            g.stats[blockIdx.x * 16 + 11] = t_issue;
            g.stats[blockIdx.x * 16 + 12] = t_wait;
            g.stats[blockIdx.x * 16 + 13] = t_math;
            g.stats[blockIdx.x * 16 + 14] = t_store;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();         // neither CTA leaves while the peer may still multicast into it / arrive on its barriers
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (F8_DBG && g.stats && tid == 0) {
        g.stats[blockIdx.x * 16 + 11] = clock64() - t_entry;
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.stats[blockIdx.x * 16 + 13] = (long long)gt;
    }
}

}  // namespace

namespace {

template <int BN, int STRIDE, bool DW = false, int MC = 0>
int launch_bn(const f8_conv_args &a, cudaStream_t s) {
    constexpr bool dw = DW;
    constexpr int MB = mb_for(BN);
    constexpr int TM = 128 * MB;
    constexpr int B_TILE = (DW ? 64 : BN) * 64;
    constexpr bool TMA = true;
    constexpr int PLANES = STRIDE == 2 ? 4 : 1;
    // an even pitch keeps every box (one padded row of PW slots x 64 B) 128-byte aligned
    const int PW = (a.wout + 2) & ~1;
    const int plane_slots = TM + (STRIDE == 2 ? PW : 2 * PW) + 2;
    const int slots = (STRIDE == 2 ? 4 : 1) * plane_slots;
    const int HPh = a.hout + 1;
    // rows per TMA box (a divisor of H + 1, so that boxes never straddle images).  Stride 1: a whole
    // padded image when that is at most 64 slots, else one row.  Otherwise (too many boxes for the
    // TMA warp, or four parity planes): the smallest divisor that keeps a stage within the box budget,
    // which is also the one that wastes the least shared memory on box alignment.
    const int box_budget = PLANES == 1 ? 32 : 64;
    auto boxes_for = [&](int d) { return PLANES * ((plane_slots + d * PW - 1) / (d * PW) + 1); };
    int BY = (PLANES == 1 && HPh * PW <= 64) ? HPh : 1;
    if (boxes_for(BY) > box_budget)
        for (int d = 1; d <= HPh; ++d)
            if (HPh % d == 0 && d * PW <= 256) {
                BY = d;
                if (boxes_for(d) <= box_budget) break;
            }
    const int box_slots = BY * PW;
    // a stage holds the boxes covering any tile's slot range: at most (range / box) + 2 boxes
    const int max_boxes = (plane_slots + box_slots - 1) / box_slots + 1;
    if (PLANES * max_boxes > box_budget) return F8_ERR_UNSUPPORTED;
    const int pps = (max_boxes * box_slots + 7) / 8 * 8;                 // slots of one plane's region
    const int slots_pad = (PLANES * pps + 15) / 16 * 16;
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
    ep.cout = a.cout;
    ep.cout_pad = a.cout_pad;
    int probe_bits = 0;
    if (const char *probe = f8host::debug_env("F8_PROBE")) {   // timing probes only: WRONG results
        const int pv = atoi(probe);
        if (pv & 1) ep.carry_in = nullptr;
        if (pv & 2) ep.carry_out = nullptr;
        if (pv & 4) ep.out0 = nullptr;
        probe_bits = pv;
    }
    const bool plain = f8::epilogue_is_plain_u8(ep);
    constexpr int SA_MAX = sa_for(STRIDE);
    const int SB_MAX = sb_for(BN, plain);
    int SA = SA_MAX, SB = SB_MAX;
    auto smem_for = [&](int sa, int sb) {
        return (size_t)sa * slots_pad * 64 + (size_t)sb * 3 * B_TILE + (3 * SA_MAX + 3 * SB_MAX + 4) * 8 + 16 +
               2 * BN * 4 + 1024;                                                      // + base alignment slack
    };
    // wide images / four parity planes: shallower rings
    while (SA > 2 && smem_for(SA, SB) > 227 * 1024) --SA;
    while (SB > 2 && smem_for(SA, SB) > 227 * 1024) --SB;
    const size_t smem_bytes = smem_for(SA, SB);
    if (smem_bytes > 227 * 1024) return F8_ERR_UNSUPPORTED;
    // the kernel owns all 512 TMEM columns: keep a second CTA off the SM
    const size_t smem_launch = smem_bytes < 120 * 1024 ? 120 * 1024 : smem_bytes;
    const long long lin = (long long)a.n * (a.hout + 1) * PW;     // padded linear output space
    // (the magic-number divisions need (lin + TM) * PW < 2^32)
    if ((lin + TM) * (long long)(PW > a.hout + 1 ? PW : a.hout + 1) >= 0xffffffffLL) return F8_ERR_UNSUPPORTED;
    const f8host::DensePack pk = f8host::dense_pack_geometry(a.cin_pad, a.cout_pad, 3, 3);
    PGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.wpack = static_cast<const uint8_t *>(a.wpack);
    g.wrows = pk.rows;
    // the stage-major copy of the weights, when its tile width is this kernel's
    g.wstage = (!dw && a.wpack_stage && ((a.cout_pad > 64 ? 128 : 64) == BN)) ? static_cast<const uint8_t *>(a.wpack_stage) : nullptr;
    if (dw) {       // block-diagonal group images behind the dp4a words of the depthwise pack
        g.wpack += f8host::dw_dense_offset(a.cin_pad);
        g.wrows = 64;
    }
    g.N = a.n; g.H = a.hout; g.W = a.wout; g.C = a.cin_pad;
    g.Hin = a.hin; g.Win = a.win;
    g.plane_slots = plane_slots;
    g.PW = PW;
    g.probe = probe_bits;
    g.lx = PW <= 8 ? 0 : (PW <= 16 ? 1 : (PW <= 32 ? 2 : 3));
    g.BY = BY;
    g.box_slots = box_slots;
    g.sa = SA;
    g.sb = SB;
    g.pps = pps;
    g.dw = dw ? 1 : 0;
    g.mPW = (uint32_t)(0x100000000ULL / (uint32_t)PW) + 1u;
    g.mHP = (uint32_t)(0x100000000ULL / (uint32_t)(a.hout + 1)) + 1u;
    g.mBY = (uint32_t)(0x100000000ULL / (uint32_t)BY) + 1u;
    g.tm = TM;
    g.slots = slots;
    g.slots_pad = slots_pad;
    g.n_super = (int)((lin + TM - 1) / TM);
    g.ntiles_n = (a.cout_pad + BN - 1) / BN;
    static f8host::DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = f8host::device_once(once, &num_sms, []() -> int {
            F8_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel<BN, false, false, STRIDE, DW, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            F8_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel<BN, true, false, STRIDE, DW, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            F8_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel<BN, false, true, STRIDE, DW, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            F8_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel<BN, true, true, STRIDE, DW, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            return F8_OK;
        });
        if (rc) return rc;
    }
    constexpr int CLS = MC ? MC : 1;
    long long grid = (long long)CLS * ((g.n_super + CLS - 1) / CLS) * g.ntiles_n;
    if (grid > num_sms) grid = num_sms;
    grid -= grid % CLS;                                   // whole clusters
    static const bool want_stats = f8host::debug_env("F8_STATS") != nullptr;
    static long long *stats_dev = nullptr;
    if (want_stats) {
        if (!stats_dev) F8_CUDA(cudaMalloc(&stats_dev, 16 * 1024 * sizeof(long long)));
        F8_CUDA(cudaMemsetAsync(stats_dev, 0, 16 * 1024 * sizeof(long long), s));
        g.stats = stats_dev;
    }
    const unsigned gr = (unsigned)grid;
    TMaps tmaps;
    memset(&tmaps, 0, sizeof(tmaps));
    for (int p = 0; p < PLANES; ++p) {
        // stride 1: the NHWC activation as (C, W, H, N).  stride 2: parity plane p = (row parity pr,
        // column parity pc): every second pixel of every second row, an (C, W/2, H/2, N) view whose
        // base is pixel (pr, pc).  Box = 64 channels x PW pixels x BY rows of one image.
        const int pr = p >> 1, pc = p & 1;
        const uint8_t *base = static_cast<const uint8_t *>(a.in) + ((size_t)pr * a.win + pc) * a.cin_pad;
        const uint64_t dims[4] = {(uint64_t)a.cin_pad, (uint64_t)a.wout, (uint64_t)a.hout, (uint64_t)a.n};
        const uint64_t strides[3] = {(uint64_t)STRIDE * a.cin_pad, (uint64_t)STRIDE * a.win * a.cin_pad,
                                     (uint64_t)a.hin * a.win * a.cin_pad};
        const uint32_t box[4] = {64u, (uint32_t)PW, (uint32_t)BY, 1u};
        const int rc = f8host::encode_tmap_u8_4d(&tmaps.m[p], base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc != F8_OK) return rc;
    }
    const unsigned th = (epi_warps_for(plain) + 4) * 32;
    constexpr int CL = CLS;
    f8host::note_kernel("conv3x3_umma<BN=%d,s%d,%s%s%s>", BN, STRIDE, DW ? "dw" : "dense", plain ? ",plain" : ",generic",
                        MC == 4 ? ",mc4" : (MC == 2 ? ",mc2" : ""));
    if (a.in_signed) {
        if (plain) F8_CUDA(f8host::launch_pdl_cluster(conv3x3_umma_kernel<BN, true, true, STRIDE, DW, MC>, gr, th, smem_launch, s, CL, g, ep, tmaps));
        else F8_CUDA(f8host::launch_pdl_cluster(conv3x3_umma_kernel<BN, true, false, STRIDE, DW, MC>, gr, th, smem_launch, s, CL, g, ep, tmaps));
    } else {
        if (plain) F8_CUDA(f8host::launch_pdl_cluster(conv3x3_umma_kernel<BN, false, true, STRIDE, DW, MC>, gr, th, smem_launch, s, CL, g, ep, tmaps));
        else F8_CUDA(f8host::launch_pdl_cluster(conv3x3_umma_kernel<BN, false, false, STRIDE, DW, MC>, gr, th, smem_launch, s, CL, g, ep, tmaps));
    }
    F8_CUDA(cudaGetLastError());
    if (want_stats) {
        static long long host[16 * 1024];
        F8_CUDA(cudaStreamSynchronize(s));
        F8_CUDA(cudaMemcpy(host, stats_dev, sizeof(host), cudaMemcpyDeviceToHost));
        double acc[16] = {0};
        for (long long b = 0; b < grid; ++b)
            for (int k = 0; k < 16; ++k) acc[k] += (double)host[b * 16 + k] / (double)grid;
        const long long items = (long long)g.n_super * g.ntiles_n;
        fprintf(stderr,
                "[f8 stats] conv3x3 BN=%d C=%d cout=%d HxW=%dx%d items=%lld/cta=%.1f | loader total %.0f wait_empty %.0f "
                "wait_cp %.0f | wload total %.0f wait_bempty %.0f | mma total %.0f wait_acc %.0f wait_a %.0f "
                "wait_b %.0f | epi total %.0f wait_full %.0f issue %.0f cpwait %.0f math %.0f store %.0f (cycles, mean per CTA)\n",
                BN, g.C, a.cout, g.H, g.W, items, (double)items / (double)grid, acc[0], acc[1], acc[2], acc[3],
                acc[4], acc[5], acc[6], acc[7], acc[8], acc[9], acc[10], acc[11], acc[12], acc[13], acc[14]);
        {
            double pro = 0, fa = 0;
            for (long long b = 0; b < grid; ++b) {
                pro += (double)(host[b * 16 + 15] >> 32) / (double)grid;
                fa += (double)(host[b * 16 + 15] & 0xffffffffll) / (double)grid;
            }
            long long e0 = host[2], e1 = host[2], x0 = host[13], x1 = host[13];
            for (long long b = 0; b < grid; ++b) {
                e0 = std::min(e0, host[b * 16 + 2]); e1 = std::max(e1, host[b * 16 + 2]);
                x0 = std::min(x0, host[b * 16 + 13]); x1 = std::max(x1, host[b * 16 + 13]);
            }
            fprintf(stderr, "[f8 stats]   globaltimer: CTA entries spread %lld ns, first entry -> first exit %lld ns, -> last exit %lld ns; mean CTA life %.0f cycles\n",
                    e1 - e0, x0 - e0, x1 - e0, acc[11]);
            fprintf(stderr, "[f8 stats]   timeline: prologue done @%.0f, first patch landed @%.0f, kernel end @%.0f cycles (acc[11])\n", pro, fa, acc[11]);
        }
    }
    return F8_OK;
}

}  // namespace

namespace f8host {

// depthwise 3x3 (stride 1, stride 2 for even input sizes) on the tensor core: every 64-channel group is a dense 64 -> 64 conv
// with a diagonal weight matrix (F8_ERR_UNSUPPORTED => the CUDA-core kernel of dw_conv.cu)
int launch_conv3x3_dw(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 3 || a.kw != 3 || a.pad != 1 || a.cin_pad != a.cout_pad || a.cin_pad % 16 != 0 || a.out_f32 != nullptr)
        return F8_ERR_UNSUPPORTED;
    if (a.stride == 1 && a.hin == a.hout && a.win == a.wout) return launch_bn<64, 1, true>(a, s);
    // stride 2: four parity planes per stage -> 256-pixel tiles, two channel groups per tile
    // (a partial last channel group costs a full one here: measured slower than the CUDA-core kernel
    // for MobileNetV2's 96- and 144-channel stride-2 layers, so those stay there)
    if (a.stride == 2 && a.hin == 2 * a.hout && a.win == 2 * a.wout && a.cout_pad % 64 == 0) {
        const int rc = launch_bn<128, 2, true>(a, s);
        // wide images: 128-pixel tiles, four channel groups per tile
        return rc == F8_ERR_UNSUPPORTED ? launch_bn<256, 2, true>(a, s) : rc;
    }
    return F8_ERR_UNSUPPORTED;
}

// F8_ERR_UNSUPPORTED => the caller falls back to the gather kernel (conv_umma.cu)
int launch_conv3x3_umma(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 3 || a.kw != 3 || a.pad != 1 || a.cin_pad % 64 != 0 || a.cout_pad % 16 != 0 ||
        a.out_f32 != nullptr)
        return F8_ERR_UNSUPPORTED;
    if (a.stride == 1 && a.hin == a.hout && a.win == a.wout) {
        if (a.cout_pad > 64) {
            // clusters of CTAs sharing the weight stream (F8_MC = 0: single CTAs, 2: pairs, 4: quads)
            static const int mc_all = [] { const char *e = getenv("F8_MC"); return e ? atoi(e) : 2; }();
            // launches with the generic epilogue (residual layers) may be given their own cluster size
            static const int mc_gen = [] { const char *e = getenv("F8_MC_GENERIC"); return e ? atoi(e) : -1; }();
            const bool plain = a.carry_in == nullptr && a.carry_out == nullptr && a.out[1] == nullptr && a.out[0] != nullptr &&
                               a.out_shift[0] > 0 && !a.out_signed[0];
            const int mc = (!plain && mc_gen >= 0) ? mc_gen : mc_all;
            if (mc == 4) {
                const int rc = launch_bn<128, 1, false, 4>(a, s);
                if (rc != F8_ERR_UNSUPPORTED) return rc;
            }
            if (mc >= 2) {
                const int rc = launch_bn<128, 1, false, 2>(a, s);
                if (rc != F8_ERR_UNSUPPORTED) return rc;
            }
            return launch_bn<128, 1>(a, s);
        }
        return launch_bn<64, 1>(a, s);
    }
    // stride 2: four parity planes of the input; even input sizes only (every F8Net stage)
    if (a.stride == 2 && a.hin == 2 * a.hout && a.win == 2 * a.wout) return launch_bn<128, 2>(a, s);
    return F8_ERR_UNSUPPORTED;
}

}  // namespace f8host
