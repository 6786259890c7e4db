// head_pool2_umma.cu -- the ResNet head in ONE kernel, second generation: 7x7 / stride 2 / pad 3
// convolution of the 3-channel image on tcgen05 (kind::i8) with NO im2col and a THREAD-LOCAL
// 3x3 / stride 2 / pad 1 max-pool, + bias, ReLU, the float32 round trip, the consumer-side
// requantisation(s) and the int32 carry.
//
// Replaces: IntModel.forward head,  x = self.head[:-1](x); x = self.head[-1](x.float()).int()
// (/root/reference/models/fix_resnet.py:355-362; head = [int Conv2d 7x7 s2 p3, ReLU,
// MaxPool2d(3, 2, 1)], fix_resnet.py:434-440) and the int_op_only_fix_quant of the first
// block's convolutions (fix_resnet.py:28-33, :57-58).
//
// Geometry.  TMEM lane = POOLED pixel.  Pooled pixel (p, q) takes the maximum of the nine conv
// outputs (2p+dy-1, 2q+dx-1), dy, dx in {0,1,2}.  The six with dx = 1, 2 are computed for that lane,
// in its own accumulator columns; the three with dx = 0 are conv column 2q-1 = the dx = 2 column of
// pooled pixel q-1, i.e. of the lane to the left, and arrive by one warp shuffle of that lane's
// maximum over dy (1.5x the minimal MACs, no shared-memory staging, no second pass).  Lanes run over
// the padded pooled index L = 58 p + q of one image (q = 56, 57 are dropped) in GROUPS OF EIGHT THAT
// OVERLAP BY ONE: lane 8g + j of a tile is L = L0 + 7g + j - 1, so that lane j = 0 of every group is
// a helper (the left neighbour of j = 1, its own result dropped) and no value ever crosses a warp:
// 112 pooled positions per 128-lane tile.  In the operand descriptor this is SBO = 112 B.
//
// The A operand is the raw NHWC4 image, never an im2col copy.  Conv output (2p+dy-1, 2q+dx-1),
// filter row r reads image row i = 4p + t - 5 (t = 2 dy + r) and the 8 pixels 4q + 2dx - 6 ..
// 4q + 2dx + 1 (the first one has weight 0): 32 contiguous bytes whose address is LINEAR in L when
// the image rows i = 4J + k are stored as four planes k with one row J per 58 lanes (928 B):
//     address(L) = plane k base + (a - amin_k) * 928 + 16 L' + e          (u = t + 3 = 4a + k)
// i.e. a K-major SWIZZLE_NONE operand whose "core matrices" overlap: LBO = 16 B (next 16 bytes of
// the window), SBO = 128 B (8 lanes further).  The three dx need 8-byte granular starts, so the
// planes are staged twice: copy A with pixel -4 at byte 0 (dx = 1), copy B with pixel -6 at byte 0
// (dx = 2 at +16 B).  Copy A arrives by TMA: four boxes per tile (4 planes x 6 rows x
// 928 B) over the image viewed as (x, k, J, n); every zero of the padding is the TMA
// out-of-bounds fill.  TMA starts are 16-byte granular, so copy B = copy A moved up by 8 bytes is
// made by two "shifter" warps, shared memory to shared memory, while the dx = 1 MMAs run.
//
// One MMA per image row t serves every dy that uses it (N = 64 |{dy}| columns, weights of filter
// rows t - 2 dy side by side): 11 MMAs per dx, 22 per tile, into 192 accumulator columns
// [dy0 | dy1 | dy2]; two such column sets alternate between the MMA warp and the epilogue.
// Epilogue (16 warps: lane group x 16-channel group): maximum over dy of the two dx phases in
// registers, the left neighbour's dx = 2 maximum by shuffle, then bias, ReLU, float round trip (x86 cvttss2si semantics), carry / 8-bit images.
// max-then-bias equals the reference's bias-then-max unless acc + bias can wrap; the kernel checks
// the bias range and otherwise applies the bias before the maximum (exact in every case).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tma_common.cuh"

namespace {

using namespace f8u;

constexpr int IMG = 224, POOLED = 56, COUT = 64;
constexpr int LP = 58;                        // lanes per pooled row (2 dropped)
constexpr int LANES_IMG = POOLED * LP;        // 3248
constexpr int TILE_L = 112;                   // pooled positions per tile: 16 groups of 8 lanes overlapping by one
constexpr int TILES_IMG = (LANES_IMG + TILE_L - 1) / TILE_L;   // 29
constexpr int ROW_BYTES = LP * 16;            // 928: one staged image row (232 pixels)
constexpr int PLANE_ROWS = 6;
constexpr int PLANE_BYTES = 5632;             // 6 * 928 = 5568 rounded up to 128
constexpr int COPY_BYTES = 4 * PLANE_BYTES;
constexpr int STAGE_BYTES = 2 * COPY_BYTES;   // 45056
constexpr int BOX_BYTES = PLANE_ROWS * ROW_BYTES;
constexpr int NSTAGE = 3;
constexpr int EPI_WARPS = 16, EPI_THREADS = 512;
constexpr int LOAD_WARP = 16, MMA_WARP = 17, SHIFT_WARP0 = 18, SHIFT_WARPS = 2;
constexpr int THREADS = 20 * 32;
constexpr int ACC_COLS = 192;

// image row t = 2 dy + r: the dy that use it, as (first dy, count)
__host__ __device__ constexpr int t_dy0(int t) { return t <= 6 ? 0 : (t <= 8 ? 1 : 2); }
__host__ __device__ constexpr int t_dy1(int t) { return t <= 1 ? 0 : (t <= 3 ? 1 : 2); }   // last dy
__host__ __device__ constexpr int t_ndy(int t) { return t_dy1(t) - t_dy0(t) + 1; }
// the last dy with r = t - 2 dy >= 0 is min(2, t / 2); the first with r <= 6 is max(0, ceil((t - 6) / 2))
static_assert(t_dy0(7) == 1 && t_dy1(7) == 2 && t_dy0(4) == 0 && t_dy1(4) == 2 && t_ndy(10) == 1, "dy ranges");
__host__ __device__ constexpr int w_off(int t) {          // byte offset of row t's weight image
    int o = 0;
    for (int i = 0; i < t; ++i) o += t_ndy(i) * COUT * 32;
    return o;
}
constexpr int W_BYTES = w_off(11);            // 43008

constexpr int OFF_STAGE = 128;                // the helper lane of a tile's first group reads 16 bytes before its row
constexpr int OFF_W = OFF_STAGE + NSTAGE * STAGE_BYTES;
constexpr int OFF_BAR = OFF_W + W_BYTES;
constexpr int NBARS = 3 * NSTAGE + 4 + 1;     // stage_full, stage_empty, acc_full[2], acc_empty[2], w_full, stageb_full
constexpr int OFF_MISC = OFF_BAR + (NBARS * 8 + 15) / 16 * 16;    // tmem slot, bias-safe flag, bias[64]
constexpr int SMEM_BYTES = OFF_MISC + 32 + COUT * 4 + 128;         // + base alignment slack

#define H2_TIMED(acc, stmt)                      \
    do {                                         \
        if (F8_DBG && g.stats) {                 \
            const long long _t0 = clock64();     \
            stmt;                                \
            acc += clock64() - _t0;              \
        } else {                                 \
            stmt;                                \
        }                                        \
    } while (0)

struct H2Geom {
    const uint8_t *wpack;   // [16][wrows][16]: chunk 2r + half of filter row r
    int wrows;
    int N;
    long long *stats;
};

template <bool A_SIGNED>
__global__ void __launch_bounds__(THREADS, 1)
head_pool2_kernel(const H2Geom g, const f8::Epilogue ep, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (f8::smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t bar_base = smem_base + OFF_BAR;
    auto stage_full = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto stage_empty = [&](int s) { return bar_base + (uint32_t)(NSTAGE + s) * 8; };
    auto acc_full = [&](int b) { return bar_base + (uint32_t)(2 * NSTAGE + b) * 8; };
    auto acc_empty = [&](int b) { return bar_base + (uint32_t)(2 * NSTAGE + 2 + b) * 8; };
    const uint32_t w_full = bar_base + (2 * NSTAGE + 4) * 8;
    auto stageb_full = [&](int s) { return bar_base + (uint32_t)(2 * NSTAGE + 5 + s) * 8; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_MISC);
    int *bias_safe = reinterpret_cast<int *>(smem + OFF_MISC + 16);
    int32_t *sbias = reinterpret_cast<int32_t *>(smem + OFF_MISC + 32);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int total_tiles = g.N * TILES_IMG;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; ++s) {
                mbar_init(stage_full(s), 1); mbar_init(stage_empty(s), 1); mbar_init(stageb_full(s), SHIFT_WARPS);
            }
            for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), EPI_WARPS); }   // one arrival per epilogue warp
            mbar_init(w_full, 1);
            fence_barrier_init();
            tma_prefetch_desc(&tmap);
            // resident weights: for image row t, the filter rows r = t - 2 dy side by side
            mbar_expect_tx(w_full, W_BYTES);
            mbar_arrive(w_full);
            for (int t = 0; t < 11; ++t) {
                const int nt = t_ndy(t) * COUT;
                for (int d = 0; d < t_ndy(t); ++d) {
                    const int r = t - 2 * (t_dy0(t) + d);
                    for (int h = 0; h < 2; ++h)
                        bulk_g2s(smem_base + OFF_W + w_off(t) + h * nt * 16 + d * COUT * 16,
                                 g.wpack + (size_t)(2 * r + h) * g.wrows * 16, COUT * 16, w_full);
                }
            }
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 512);
    }
    if (warp == 0) {
        // acc + bias cannot wrap when |bias| <= 2^31 - 1 - 255 * 127 * 147
        const int32_t b0 = __ldg(ep.bias + lane), b1 = __ldg(ep.bias + 32 + lane);
        sbias[lane] = b0;
        sbias[32 + lane] = b1;
        const int32_t lim = 2147483647 - 255 * 127 * 147;
        const bool ok = b0 <= lim && b0 >= -lim && b1 <= lim && b1 >= -lim;
        const bool all_ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) *bias_safe = all_ok ? 1 : 0;
    }
    if (tid < NSTAGE * 4)      // bytes 0..7 of every copy-B plane (pixels -6, -5): never written again
        *reinterpret_cast<uint2 *>(smem + OFF_STAGE + (tid >> 2) * STAGE_BYTES + COPY_BYTES + (tid & 3) * PLANE_BYTES) =
            make_uint2(0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: weights, bias and TMEM are in place before the input
    // conversion launch (or the previous pass) has finished
    f8::pdl_trigger();
    f8::pdl_wait();

    if (warp == LOAD_WARP) {
        // =========================== image loader (TMA) ===========================
        int slot = 0, phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int img = t / TILES_IMG;
            const int L0 = (t - img * TILES_IMG) * TILE_L;
            const int p0 = L0 / LP;
            mbar_wait(stage_empty(slot), phase ^ 1);
            if (lane == 0) {
                mbar_expect_tx(stage_full(slot), 4 * BOX_BYTES);
                mbar_arrive(stage_full(slot));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // plane k starts at row J = p0 + amin_k - 2, amin = {1, 1, 1, 0}
                    const int J0 = p0 - (k == 3 ? 2 : 1);
                    tma_load_4d(smem_base + OFF_STAGE + slot * STAGE_BYTES + k * PLANE_BYTES, &tmap, -4, k, J0, img,
                                stage_full(slot));
                }
            }
            __syncwarp();
            if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
        }
    } else if (warp >= SHIFT_WARP0) {
        // =========================== shifters: copy B = copy A moved up by 8 bytes ===========
        // plane-wide: B[8 .. 5568) = A[0 .. 5560); the 8 bytes that cross a row boundary are the
        // zero pixels 226, 227 of one row becoming the zero pixels -6, -5 of the next
        const int sw = warp - SHIFT_WARP0;
        int slot = 0, phase = 0;
        long long w_sa = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            H2_TIMED(w_sa, mbar_wait(stage_full(slot), phase));
#pragma unroll
            for (int kk = 0; kk < 4 / SHIFT_WARPS; ++kk) {
                const int k = sw * (4 / SHIFT_WARPS) + kk;
                const uint8_t *srcp = smem + OFF_STAGE + slot * STAGE_BYTES + k * PLANE_BYTES;
                uint8_t *dstp = smem + OFF_STAGE + slot * STAGE_BYTES + COPY_BYTES + k * PLANE_BYTES + 8;
#pragma unroll 4
                for (int i = lane; i < BOX_BYTES / 16; i += 32) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(srcp + i * 16);
                    *reinterpret_cast<uint2 *>(dstp + i * 16) = make_uint2(v.x, v.y);
                    if (i * 16 + 16 < BOX_BYTES) *reinterpret_cast<uint2 *>(dstp + i * 16 + 8) = make_uint2(v.z, v.w);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(stageb_full(slot));     // one arrival per shifter warp
            if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
        }
        if (F8_DBG && g.stats && tid == SHIFT_WARP0 * 32) {
            g.stats[blockIdx.x * 16 + 6] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 7] = w_sa;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===================================
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);            // weights: SBO = 128 B, version 1
        constexpr uint32_t desc_hi_a = (112u >> 4) | (1u << 14);          // image: SBO = 112 B: groups of 8 lanes overlap by one
        constexpr uint32_t a_lbo = (16u >> 4) << 16;                      // LBO = 16 B: overlapping windows
        mbar_wait(w_full, 0);
        const uint32_t w_lo0 = ((smem_base + OFF_W) & 0x3ffffu) >> 4;
        int slot = 0, phase = 0;
        uint32_t ph = 0;                                                  // dx phase counter
        long long w_stage = 0, w_stageb = 0, w_acc = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int img = t / TILES_IMG;
            const int L0 = (t - img * TILES_IMG) * TILE_L;
            const int p0 = L0 / LP;
            H2_TIMED(w_stage, mbar_wait(stage_full(slot), phase));
            tc_fence_after();
            // descriptor start (16-byte units) of lane 0 (= position L0 - 1) in copy A, plane 0, row 0
            const uint32_t a_lo0 =
                (((smem_base + OFF_STAGE + slot * STAGE_BYTES) & 0x3ffffu) >> 4) + (uint32_t)(L0 - 1 - p0 * LP);
#pragma unroll
            for (int dxi = 0; dxi < 2; ++dxi) {
                constexpr int kDx[2] = {1, 2};           // copy A (dx = 1) first: copy B is still being made
                const int dx = kDx[dxi];
                const int buf = ph & 1;
                if (dxi == 1) { H2_TIMED(w_stageb, mbar_wait(stageb_full(slot), phase)); tc_fence_after(); }
                H2_TIMED(w_acc, mbar_wait(acc_empty(buf), ((ph >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
                if (elect_one()) {
#pragma unroll
                    for (int ti = 0; ti < 11; ++ti) {
                        // the three-dy rows first (t = 4 overwrites all 192 columns), then rows whose
                        // column ranges are already initialised
                        constexpr int kOrder[11] = {4, 5, 6, 2, 3, 7, 8, 0, 1, 9, 10};
                        const int tt = kOrder[ti];
                        const int u = tt + 3, a = u >> 2, k = u & 3;
                        const int amin = k == 3 ? 0 : 1;
                        const int nt = t_ndy(tt) * COUT;
                        const uint32_t a_off = (uint32_t)(((dx == 1 ? 0 : COPY_BYTES) + k * PLANE_BYTES +
                                                           (a - amin) * ROW_BYTES + (dx == 2 ? 16 : 0)) >> 4);
                        umma_i8_lohi(tacc + (uint32_t)(t_dy0(tt) * COUT), (a_lo0 + a_off) | a_lbo, desc_hi_a,
                                     (w_lo0 + (uint32_t)(w_off(tt) >> 4)) | ((uint32_t)((nt * 16) >> 4) << 16), desc_hi,
                                     instr_desc(A_SIGNED, nt), ti ? 1u : 0u);
                    }
                    umma_commit(acc_full(buf));
                    if (dxi == 1) umma_commit(stage_empty(slot));
                }
                __syncwarp();
                ++ph;
            }
            if (++slot == NSTAGE) { slot = 0; phase ^= 1; }
        }
        if (F8_DBG && g.stats && lane == 0) {
            g.stats[blockIdx.x * 16 + 0] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 1] = w_stage;
            g.stats[blockIdx.x * 16 + 2] = w_stageb;
            g.stats[blockIdx.x * 16 + 3] = w_acc;
        }
    } else {
        // =========================== epilogue (warps 0-15) ========================
        const int lg = warp & 3;                 // TMEM lane group
        const int cgp = warp >> 2;               // 16-channel group
        const bool safe = *bias_safe != 0;
        const int32_t *b16 = sbias + cgp * 16;
        uint32_t ph = 0;
        long long w_full = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int img = t / TILES_IMG;
            // lane 8g + j of the tile is pooled position L0 + 7g + j - 1; j = 0 is the group's helper lane
            const int li = lg * 32 + lane;
            const int L = (t - img * TILES_IMG) * TILE_L + 7 * (li >> 3) + (li & 7) - 1;
            const int p = L < 0 ? 0 : L / LP, q = L < 0 ? LP : L - p * LP;
            const bool valid = (li & 7) != 0 && p < POOLED && q < POOLED;
            // conv row / column -1 is padding of the max-pool, not a conv output
            const bool no_dy0 = p == 0, no_dx0 = q == 0;
            const int32_t ident = safe ? (int32_t)0x80000000 : 0;
            int32_t m[16];      // maximum over dy of the dx = 1 column, then of everything
            int32_t m2[16];     // maximum over dy of the dx = 2 column (the right neighbour's dx = 0 column)
#pragma unroll
            for (int dxi = 0; dxi < 2; ++dxi) {
                const int buf = ph & 1;
                H2_TIMED(w_full, mbar_wait(acc_full(buf), (ph >> 1) & 1));
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ACC_COLS + cgp * 16);
                int32_t v0[16], v1[16], v2[16];
                tmem_ld16(trow, v0);
                tmem_ld16(trow + COUT, v1);
                tmem_ld16(trow + 2 * COUT, v2);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(buf));     // this warp's columns are in registers
                // the pool's padding row: replace the excluded values by the identity of max
                // (rare: only lanes of the first pooled row, so a branch, not 16 selects)
                if (no_dy0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v0[i] = ident;
                }
                int32_t (&dst)[16] = dxi == 0 ? m : m2;
                if (safe) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dst[i] = max(v0[i], max(v1[i], v2[i]));
                } else {
                    // bias first (wrapping, like the reference's conv), ReLU through the identity 0; an excluded
                    // value must stay at the identity, so the bias is skipped for it
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t b = (uint32_t)b16[i];
                        const int32_t a0 = no_dy0 ? 0 : (int32_t)((uint32_t)v0[i] + b);
                        const int32_t a1 = (int32_t)((uint32_t)v1[i] + b);
                        const int32_t a2 = (int32_t)((uint32_t)v2[i] + b);
                        dst[i] = max(max(0, a0), max(a1, a2));
                    }
                }
                ++ph;
            }
            // dx = 0: conv column 2q - 1 is the dx = 2 column of the lane to the left (same pooled row: q >= 1);
            // the shuffle never crosses a group of eight, whose first lane is the helper
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int32_t left = __shfl_up_sync(0xffffffffu, m2[i], 1);
                m[i] = max(m[i], max(m2[i], no_dx0 ? ident : left));
            }
            if (valid) {
                // .float() max-pool .int() (fix_resnet.py:358-359): the round trip is the identity for
                // 0 <= x < 2^24, which one OR over the 16 (non-negative) values proves; the conversions
                // (quarter-rate pipe) run only for a thread that holds a larger value
                int32_t r[16];
                uint32_t any = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    r[i] = safe ? max((int32_t)((uint32_t)m[i] + (uint32_t)b16[i]), 0) : m[i];
                    any |= (uint32_t)r[i];
                }
                if (!ep.int_pool && any >= (1u << 24)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = f8::f2i_x86((float)r[i]);
                }
                const size_t opix = ((size_t)img * POOLED + p) * POOLED + q;
                const int ch0 = cgp * 16;
                if (ep.carry_out) {
                    int32_t *dst = ep.carry_out + f8::carry_off(opix, ch0, COUT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<int4 *>(dst + k * 512) = make_int4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
                }
                const size_t o = opix * COUT + ch0;
                if (ep.out0) *reinterpret_cast<uint4 *>(ep.out0 + o) = f8::requant_pack16(r, ep.shift0, ep.signed0);
                if (ep.out1) *reinterpret_cast<uint4 *>(ep.out1 + o) = f8::requant_pack16(r, ep.shift1, ep.signed1);
            }
        }
        if (F8_DBG && g.stats && tid == 0) {
            g.stats[blockIdx.x * 16 + 4] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 5] = w_full;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

namespace f8host {

// a = the head convolution's arguments with hout/wout = the POOLED size (56) and the epilogue
// of the pooled tensor.  F8_ERR_UNSUPPORTED => the caller runs conv + maxpool separately.
int launch_head_pool(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 7 || a.kw != 7 || a.stride != 2 || a.pad != 3 || a.cin_pad != 4 || a.cout != COUT ||
        a.cout_pad != COUT || a.hin != IMG || a.win != IMG || a.hout != POOLED || a.wout != POOLED ||
        a.carry_in != nullptr || a.out_f32 != nullptr)
        return F8_ERR_UNSUPPORTED;
    const DensePack pk = dense_pack_geometry(4, COUT, 7, 7);
    if (pk.mode != 1 || pk.row_bytes != 32 || pk.shift_px != 1) return F8_ERR_UNSUPPORTED;
    H2Geom g{};
    g.wpack = static_cast<const uint8_t *>(a.wpack);
    g.wrows = pk.rows;
    g.N = a.n;
    f8::Epilogue ep{};
    ep.bias = a.bias;
    ep.carry_out = a.carry_out;
    ep.out0 = static_cast<uint8_t *>(a.out[0]);
    ep.out1 = static_cast<uint8_t *>(a.out[1]);
    ep.shift0 = a.out_shift[0]; ep.signed0 = a.out_signed[0];
    ep.shift1 = a.out_shift[1]; ep.signed1 = a.out_signed[1];
    ep.cout = a.cout;
    ep.cout_pad = a.cout_pad;
    ep.int_pool = (a.flags & F8_OPF_INT_MAXPOOL) != 0;
    static DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = device_once(once, &num_sms, []() -> int {
            F8_CUDA(cudaFuncSetAttribute(head_pool2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            F8_CUDA(cudaFuncSetAttribute(head_pool2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            return F8_OK;
        });
        if (rc) return rc;
    }
    // the image as one uint32 per pixel, rows split i = 4 J + k: (x, k, J, n); a box is six rows
    // J of one plane k, 232 pixels wide from x = -4 (copy A) or x = -6 (copy B)
    CUtensorMap tmap;
    const uint64_t dims[4] = {(uint64_t)IMG, 4u, (uint64_t)IMG / 4, (uint64_t)a.n};
    const uint64_t strides[3] = {(uint64_t)IMG * 4, (uint64_t)IMG * 16, (uint64_t)IMG * IMG * 4};
    const uint32_t box[4] = {(uint32_t)(ROW_BYTES / 4), 1u, (uint32_t)PLANE_ROWS, 1u};
    const int rc = encode_tmap_u32_4d(&tmap, a.in, dims, strides, box);
    if (rc != F8_OK) return rc;
    long long grid = (long long)a.n * TILES_IMG;
    if (grid > num_sms) grid = num_sms;
    static const bool want_stats = debug_env("F8_STATS") != nullptr;
    static long long *stats_dev = nullptr;
    if (want_stats) {
        if (!stats_dev) F8_CUDA(cudaMalloc(&stats_dev, 16 * 1024 * sizeof(long long)));
        F8_CUDA(cudaMemsetAsync(stats_dev, 0, 16 * 1024 * sizeof(long long), s));
        g.stats = stats_dev;
    }
    note_kernel("head_pool2");
    if (a.in_signed) F8_CUDA(f8host::launch_pdl(head_pool2_kernel<true>, (unsigned)grid, THREADS, SMEM_BYTES, s, g, ep, tmap));
    else F8_CUDA(f8host::launch_pdl(head_pool2_kernel<false>, (unsigned)grid, THREADS, SMEM_BYTES, s, g, ep, tmap));
    F8_CUDA(cudaGetLastError());
    if (want_stats) {
        static long long host[16 * 1024];
        F8_CUDA(cudaStreamSynchronize(s));
        F8_CUDA(cudaMemcpy(host, stats_dev, sizeof(host), cudaMemcpyDeviceToHost));
        double acc[16] = {0};
        for (long long b = 0; b < grid; ++b)
            for (int k = 0; k < 16; ++k) acc[k] += (double)host[b * 16 + k] / (double)grid;
        fprintf(stderr,
                "[f8 stats] head_pool2 tiles/cta=%.1f | mma total %.0f wait_stage %.0f wait_stageb %.0f wait_acc %.0f | "
                "epi total %.0f wait_full %.0f | shift total %.0f wait_stage %.0f (cycles, mean per CTA)\n",
                (double)a.n * TILES_IMG / (double)grid, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]);
    }
    return F8_OK;
}

}  // namespace f8host
