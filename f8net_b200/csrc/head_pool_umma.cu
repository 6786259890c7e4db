// head_pool_umma.cu -- the ResNet head in ONE kernel: 7x7 / stride 2 / pad 3 convolution of the
// 3-channel image (tcgen05.mma kind::i8), + bias, ReLU, the float32 round trip and the 3x3 /
// stride 2 / pad 1 max-pool, then the consumer-side requantisation(s) and the int32 carry.
// The 112x112x64 int32 head activation (3.2 MB per image) never leaves the SM.
//
// Replaces: IntModel.forward head,  x = self.head[:-1](x); x = self.head[-1](x.float()).int()
// (/root/reference/models/fix_resnet.py:355-362; head = [int Conv2d 7x7 s2 p3, ReLU,
// MaxPool2d(3, 2, 1)], fix_resnet.py:434-440) and the int_op_only_fix_quant of the first
// block's convolutions (fix_resnet.py:28-33, :57-58).
//
// One tile = 4 pooled rows x 56 pooled columns of one image = conv rows 2*pr0-1 .. 2*pr0+7
// (9 rows x 112 = 1008 conv pixels = 8 MMA segments of 128) -- one conv row of overlap between
// neighbouring tiles is recomputed.  Pipeline per tile:
//   builders (warps 16-19)  stage the 23 x 232-pixel input patch (NHWC4 bytes, zero fill outside
//             the image) with cp.async, then expand it segment by segment into the im2col
//             operand [K/16][128 rows][16 B] (K = 7 filter rows x 8-pixel window x 4 B = 224,
//             padded to 256) with shared->shared copies: global memory is read once.
//   MMA (warp 20)  8 K=32 MMAs per segment into TMEM columns [64*s, 64*s+64); the 16 KB weight
//             image stays resident in shared memory for the whole kernel.
//   epilogue (warps 0-15) four passes of 16 channels: TMEM -> +bias, ReLU, int->float, horizontal
//             3-max with the neighbouring lanes (shuffles) -> shared staging tile; then each of
//             448 threads takes the vertical 3-max of 8 channels of one pooled pixel (on the integers: int -> float
//             is monotone), applies the float32 round trip with the x86 cvttss2si semantics of
//             .int(), and writes carry / 8-bit images.
#include <cstdio>
#include <cstdlib>

#include "umma_common.cuh"

namespace {

using namespace f8u;

constexpr int IMG = 224, CONV = 112, POOLED = 56, COUT = 64;
constexpr int TP = 4;                       // pooled rows per tile
constexpr int CROWS = 2 * TP + 1;           // conv rows per tile (9)
constexpr int CPIX = CROWS * CONV;          // 1008 conv pixels
constexpr int SEGS = 8;                     // MMA segments of 128 rows
constexpr int PROWS = 2 * CROWS + 5;        // input rows per tile (23)
constexpr int PPITCH = (IMG + 8) * 4;       // patch row pitch in bytes: 4 px pad left and right
constexpr int PATCH_BYTES = PROWS * PPITCH; // 21344
constexpr int KPAD = 256;                   // K bytes per conv pixel (224 used)
constexpr int A_STAGE = 128 * KPAD;         // 32 KB: [16 chunks][128][16]
constexpr int A_CHUNK = 128 * 16;
constexpr int SA = 2;
constexpr int W_BYTES = COUT * KPAD;        // 16 KB: [16 chunks][64][16]
constexpr int STAGE_BYTES = CROWS * POOLED * 64 + 32 * 64;   // horizontally pooled tile (16 channels x 4 B per entry) + side buffer
// the epilogue is instruction-bound (~8 integer instructions per conv element): 16 warps
constexpr int EPI_THREADS = 512;
constexpr int BUILD_WARP0 = 16, BUILDERS = 128;
constexpr int MMA_WARP = 20;
constexpr int THREADS = 21 * 32;

constexpr int OFF_PATCH = 0;                                   // 2 patch buffers
constexpr int OFF_A = OFF_PATCH + 2 * ((PATCH_BYTES + 127) / 128 * 128);
constexpr int OFF_W = OFF_A + SA * A_STAGE;
constexpr int OFF_STAGE = OFF_W + W_BYTES;
constexpr int OFF_BAR = OFF_STAGE + STAGE_BYTES;
constexpr int NBARS = 2 * SA + 3;                              // a_full, a_empty, acc_full, acc_empty, w_full
constexpr int OFF_MISC = OFF_BAR + (NBARS * 8 + 15) / 16 * 16;  // tmem slot (16 B) + bias (64 ints)
static_assert(OFF_MISC % 16 == 0 && OFF_STAGE % 16 == 0 && OFF_A % 128 == 0, "smem carve-up alignment");
constexpr int SMEM_BYTES = OFF_MISC + 16 + COUT * 4;

struct HGeom {
    const uint8_t *in;      // NHWC4 8-bit [N][224][224][4]
    const uint8_t *wpack;   // [16][wrows][16]
    int wrows;
    int N;
    long long *stats;       // debug (F8_STATS=1)
};

#define F8_TIMED(acc, stmt)                      \
    do {                                         \
        if (g.stats) {                           \
            const long long _t0 = clock64();     \
            stmt;                                \
            acc += clock64() - _t0;              \
        } else {                                 \
            stmt;                                \
        }                                        \
    } while (0)

__device__ __forceinline__ uint32_t stage_off(int m, int chunk) {
    // 64-byte record per conv pixel, 16-byte chunks XOR-swizzled so that both the per-pixel
    // writes (consecutive m) and the stride-2 pooled reads spread over the banks
    return (uint32_t)(m * 64 + ((chunk ^ ((m >> 1) & 3)) << 4));
}

template <bool A_SIGNED>
__global__ void __launch_bounds__(THREADS, 1)
head_pool_kernel(const HGeom g, const f8::Epilogue ep) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t bar_base = smem_base + OFF_BAR;
    auto a_full = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto a_empty = [&](int s) { return bar_base + (uint32_t)(SA + s) * 8; };
    const uint32_t acc_full = bar_base + 2 * SA * 8;
    const uint32_t acc_empty = bar_base + (2 * SA + 1) * 8;
    const uint32_t w_full = bar_base + (2 * SA + 2) * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_MISC);
    int32_t *sbias = reinterpret_cast<int32_t *>(smem + OFF_MISC + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = POOLED / TP;                  // 14
    const int total_tiles = g.N * tiles_per_img;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < SA; ++s) { mbar_init(a_full(s), BUILDERS); mbar_init(a_empty(s), 1); }
            mbar_init(acc_full, 1);
            mbar_init(acc_empty, EPI_THREADS);
            mbar_init(w_full, 1);
            fence_barrier_init();
            // resident weights: 16 chunks of 64 rows x 16 B
            mbar_expect_tx(w_full, W_BYTES);
            mbar_arrive(w_full);
            for (int j = 0; j < 16; ++j)
                bulk_g2s(smem_base + OFF_W + j * (COUT * 16), g.wpack + (size_t)j * g.wrows * 16,
                         COUT * 16, w_full);
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 512);
    }
    if (tid < COUT) sbias[tid] = __ldg(ep.bias + tid);
    // K chunks 14, 15 of every operand stage are padding: zero them once
    for (int i = tid; i < SA * 2 * 128; i += THREADS) {
        const int s = i / 256, r = i % 256;
        *reinterpret_cast<uint4 *>(smem + OFF_A + s * A_STAGE + 14 * A_CHUNK + r * 16) =
            make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= BUILD_WARP0 && warp < MMA_WARP) {
        // =========================== builders =====================================
        const int bt = tid - BUILD_WARP0 * 32;
        auto load_patch = [&](int t, int pbuf) {
            const int img = t / tiles_per_img;
            const int pr0 = (t - img * tiles_per_img) * TP;
            const int in_row0 = 2 * (2 * pr0 - 1) - 3;       // input row of patch row 0
            const uint32_t dst0 = smem_base + OFF_PATCH + pbuf * ((PATCH_BYTES + 127) / 128 * 128);
            const uint8_t *src_img = g.in + (size_t)img * IMG * IMG * 4;
            constexpr int CPR = PPITCH / 16;                  // 58 chunks per patch row
            for (int c = bt; c < PROWS * CPR; c += BUILDERS) {
                const int pr = c / CPR, j = c - pr * CPR;
                const int y = in_row0 + pr;
                const int x = 4 * (j - 1);                    // chunk = 4 pixels
                const bool ok = (unsigned)y < (unsigned)IMG && (unsigned)x < (unsigned)IMG;
                const uint8_t *src = ok ? src_img + ((size_t)y * IMG + x) * 4 : g.in;
                cp_async16(dst0 + pr * PPITCH + j * 16, src, ok);
            }
            cp_async_commit();
        };
        int slot = 0, phase = 0, pbuf = 0;
        long long w_patch = 0, w_aempty = 0;
        const long long t_begin = clock64();
        if ((int)blockIdx.x < total_tiles) load_patch(blockIdx.x, 0);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            F8_TIMED(w_patch, cp_async_wait<0>(); asm volatile("bar.sync 2, %0;" ::"n"(BUILDERS) : "memory"));   // whole patch visible
            if (t + (int)gridDim.x < total_tiles) load_patch(t + gridDim.x, pbuf ^ 1);
            const uint8_t *patch = smem + OFF_PATCH + pbuf * ((PATCH_BYTES + 127) / 128 * 128);
            for (int s = 0; s < SEGS; ++s) {
                F8_TIMED(w_aempty, mbar_wait(a_empty(slot), phase ^ 1));
                uint8_t *sa = smem + OFF_A + slot * A_STAGE;
                const int m = s * 128 + bt;
                if (m < CPIX) {
                    const int cr = m / CONV, cq = m - cr * CONV;
                    const uint8_t *win = patch + (2 * cr) * PPITCH + 8 * cq;   // 8-byte aligned
#pragma unroll
                    for (int fr = 0; fr < 7; ++fr) {
                        const uint2 *w8 = reinterpret_cast<const uint2 *>(win + fr * PPITCH);
                        const uint2 a = w8[0], b = w8[1], c = w8[2], d = w8[3];
                        *reinterpret_cast<uint4 *>(sa + (2 * fr) * A_CHUNK + bt * 16) =
                            make_uint4(a.x, a.y, b.x, b.y);
                        *reinterpret_cast<uint4 *>(sa + (2 * fr + 1) * A_CHUNK + bt * 16) =
                            make_uint4(c.x, c.y, d.x, d.y);
                    }
                }
                fence_proxy_async();
                mbar_arrive(a_full(slot));
                if (++slot == SA) { slot = 0; phase ^= 1; }
            }
            // every builder is done reading this patch buffer before it is refilled two tiles on
            asm volatile("bar.sync 2, %0;" ::"n"(BUILDERS) : "memory");
            pbuf ^= 1;
        }
        cp_async_wait<0>();
        if (g.stats && bt == 0) {
            g.stats[blockIdx.x * 16 + 0] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 1] = w_patch;
            g.stats[blockIdx.x * 16 + 2] = w_aempty;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===================================
        constexpr uint32_t idesc = instr_desc(A_SIGNED, COUT);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
        constexpr uint32_t a_lbo = ((uint32_t)A_CHUNK >> 4) << 16;
        constexpr uint32_t b_lbo = ((uint32_t)(COUT * 16) >> 4) << 16;
        mbar_wait(w_full, 0);
        const uint32_t b_lo0 = (((smem_base + OFF_W) & 0x3ffffu) >> 4) | b_lbo;
        int slot = 0, phase = 0, tphase = 0;
        long long w_acc = 0, w_a = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            F8_TIMED(w_acc, mbar_wait(acc_empty, tphase ^ 1));            // epilogue has drained the previous tile
            tc_fence_after();
            for (int s = 0; s < SEGS; ++s) {
                F8_TIMED(w_a, mbar_wait(a_full(slot), phase));
                tc_fence_after();
                const uint32_t a_lo0 = (((smem_base + OFF_A + slot * A_STAGE) & 0x3ffffu) >> 4) | a_lbo;
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < KPAD / 32; ++kk)
                        umma_i8_lohi(tmem_base + (uint32_t)(s * COUT),
                                     a_lo0 + (uint32_t)kk * ((2 * A_CHUNK) >> 4), desc_hi,
                                     b_lo0 + (uint32_t)kk * ((2 * COUT * 16) >> 4), desc_hi, idesc,
                                     kk ? 1u : 0u);
                    umma_commit(a_empty(slot));
                }
                __syncwarp();
                if (++slot == SA) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
            tphase ^= 1;
        }
        if (g.stats && lane == 0) {
            g.stats[blockIdx.x * 16 + 3] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 4] = w_acc;
            g.stats[blockIdx.x * 16 + 5] = w_a;
        }
    } else {
        // =========================== epilogue (warps 0-7) =========================
        const int lg = warp & 3;                 // TMEM lane group
        const int sq = warp >> 2;                // segments 2*sq, 2*sq+1
        uint8_t *stage = smem + OFF_STAGE;
        uint8_t *side = smem + OFF_STAGE + CROWS * POOLED * 64;   // raw lane-31 values, 32 x 64 B
        int tphase = 0;
        long long w_full = 0, t_p1 = 0, t_p2 = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int img = t / tiles_per_img;
            const int pr0 = (t - img * tiles_per_img) * TP;
            F8_TIMED(w_full, mbar_wait(acc_full, tphase));
            tc_fence_after();
            for (int grp = 0; grp < 4; ++grp) {
                const long long tp0 = g.stats ? clock64() : 0;
                // ---- phase 1: accumulators -> relu -> float bits -> HORIZONTAL 3-max -> staging ----
                // Conv rows start at even lanes (112 = 3.5 warps), so the pooling centres (even
                // columns) are the even lanes and their neighbours are the adjacent lanes: two
                // shuffles per value.  Only a centre at lane 0 lacks its left neighbour (lane 31 of
                // the previous 32-pixel run): every lane 31 parks its raw values in a side buffer
                // that phase 2 folds in.  The staging tile shrinks to 9 x 56 entries of 64 B.
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int s = sq * 2 + j;
                    const int m = s * 128 + lg * 32 + lane;
                    const int cr = m / CONV, col = m - cr * CONV;
                    int32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(s * COUT + grp * 16), v);
                    tmem_ld_wait();
                    int32_t x[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int4 b = *reinterpret_cast<const int4 *>(sbias + grp * 16 + 4 * q);
                        // int -> float is monotone, so the pooling maximum is taken on the integers and
                        // the float32 round trip is applied once per pooled value in phase 2
                        x[4 * q + 0] = max((int32_t)((uint32_t)v[4 * q + 0] + (uint32_t)b.x), 0);
                        x[4 * q + 1] = max((int32_t)((uint32_t)v[4 * q + 1] + (uint32_t)b.y), 0);
                        x[4 * q + 2] = max((int32_t)((uint32_t)v[4 * q + 2] + (uint32_t)b.z), 0);
                        x[4 * q + 3] = max((int32_t)((uint32_t)v[4 * q + 3] + (uint32_t)b.w), 0);
                    }
                    if (lane == 31) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<int4 *>(side + (s * 4 + lg) * 64 + q * 16) =
                                make_int4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                    }
                    const bool use_left = lane > 0 && col > 0;
                    int32_t hm[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int32_t l = __shfl_up_sync(0xffffffffu, x[i], 1);
                        const int32_t r = __shfl_down_sync(0xffffffffu, x[i], 1);
                        hm[i] = max(max(x[i], r), use_left ? l : 0);
                    }
                    if (m < CPIX && !(col & 1)) {
                        const int e = cr * POOLED + (col >> 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<int4 *>(stage + stage_off(e, q)) =
                                make_int4(hm[4 * q], hm[4 * q + 1], hm[4 * q + 2], hm[4 * q + 3]);
                    }
                }
                if (grp == 3) {
                    tc_fence_before();
                    mbar_arrive(acc_empty);          // TMEM of this tile fully read
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                const long long tp1 = g.stats ? clock64() : 0;
                t_p1 += tp1 - tp0;
                // ---- phase 2: vertical 3-max, .int(), carry + requantised images ----
                if (tid < 2 * TP * POOLED) {
                    // thread = (pooled pixel, 8-channel half of the 16-channel group)
                    const int pp = tid >> 1, hf = tid & 1;
                    const int pr = pp / POOLED, pq = pp - pr * POOLED;
                    int4 mx[2];
                    mx[0] = make_int4(0, 0, 0, 0);
                    mx[1] = make_int4(0, 0, 0, 0);
#pragma unroll
                    for (int dr = 0; dr < 3; ++dr) {
                        const int cr = 2 * pr + dr;                       // local conv row
                        if (2 * pr0 - 1 + cr < 0) continue;               // conv row -1 (padding)
                        const int e = cr * POOLED + pq;
                        const int mc = cr * CONV + 2 * pq;                // conv pixel of the centre
                        const bool need_side = (mc & 31) == 0 && pq > 0;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int4 xv = *reinterpret_cast<const int4 *>(stage + stage_off(e, 2 * hf + q));
                            mx[q].x = max(mx[q].x, xv.x); mx[q].y = max(mx[q].y, xv.y);
                            mx[q].z = max(mx[q].z, xv.z); mx[q].w = max(mx[q].w, xv.w);
                            if (need_side) {
                                const int4 sv = *reinterpret_cast<const int4 *>(side + ((mc - 1) >> 5) * 64 + (2 * hf + q) * 16);
                                mx[q].x = max(mx[q].x, sv.x); mx[q].y = max(mx[q].y, sv.y);
                                mx[q].z = max(mx[q].z, sv.z); mx[q].w = max(mx[q].w, sv.w);
                            }
                        }
                    }
                    int32_t r[8];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        r[4 * q + 0] = f8::f2i_x86((float)mx[q].x);
                        r[4 * q + 1] = f8::f2i_x86((float)mx[q].y);
                        r[4 * q + 2] = f8::f2i_x86((float)mx[q].z);
                        r[4 * q + 3] = f8::f2i_x86((float)mx[q].w);
                    }
                    const size_t opix = ((size_t)img * POOLED + (pr0 + pr)) * POOLED + pq;
                    const int ch0 = grp * 16 + hf * 8;
                    const size_t o = opix * ep.cout_pad + ch0;
                    if (ep.carry_out) {
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            *reinterpret_cast<int4 *>(ep.carry_out + f8::carry_off(opix, ch0 + 4 * q, ep.cout_pad)) =
                                make_int4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                    }
                    if (ep.out0)
                        *reinterpret_cast<uint2 *>(ep.out0 + o) =
                            make_uint2(f8::requant_pack4(r[0], r[1], r[2], r[3], ep.shift0, ep.signed0),
                                       f8::requant_pack4(r[4], r[5], r[6], r[7], ep.shift0, ep.signed0));
                    if (ep.out1)
                        *reinterpret_cast<uint2 *>(ep.out1 + o) =
                            make_uint2(f8::requant_pack4(r[0], r[1], r[2], r[3], ep.shift1, ep.signed1),
                                       f8::requant_pack4(r[4], r[5], r[6], r[7], ep.shift1, ep.signed1));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (g.stats) t_p2 += clock64() - tp1;
            }
            tphase ^= 1;
        }
        if (g.stats && tid == 0) {
            g.stats[blockIdx.x * 16 + 6] = clock64() - t_begin;
            g.stats[blockIdx.x * 16 + 7] = w_full;
            g.stats[blockIdx.x * 16 + 8] = t_p1;
            g.stats[blockIdx.x * 16 + 9] = t_p2;
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
int launch_head_pool_v1(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 7 || a.kw != 7 || a.stride != 2 || a.pad != 3 || a.cin_pad != 4 || a.cout != COUT ||
        a.cout_pad != COUT || a.hin != IMG || a.win != IMG || a.hout != POOLED || a.wout != POOLED ||
        a.carry_in != nullptr || a.out_f32 != nullptr)
        return F8_ERR_UNSUPPORTED;
    const DensePack pk = dense_pack_geometry(4, COUT, 7, 7);
    if (pk.mode != 1 || pk.row_bytes != 32 || pk.shift_px != 1 || pk.K_pad != KPAD) return F8_ERR_UNSUPPORTED;
    HGeom g{};
    g.in = static_cast<const uint8_t *>(a.in);
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
    static bool attr_done = false;
    static int num_sms = 0;
    if (!attr_done) {
        F8_CUDA(cudaFuncSetAttribute(head_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        F8_CUDA(cudaFuncSetAttribute(head_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int dev = 0;
        F8_CUDA(cudaGetDevice(&dev));
        F8_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    long long grid = (long long)a.n * (POOLED / TP);
    if (grid > num_sms) grid = num_sms;
    static const bool want_stats = getenv("F8_STATS") != nullptr;
    static long long *stats_dev = nullptr;
    if (want_stats) {
        if (!stats_dev) F8_CUDA(cudaMalloc(&stats_dev, 16 * 1024 * sizeof(long long)));
        F8_CUDA(cudaMemsetAsync(stats_dev, 0, 16 * 1024 * sizeof(long long), s));
        g.stats = stats_dev;
    }
    if (a.in_signed) head_pool_kernel<true><<<(unsigned)grid, THREADS, SMEM_BYTES, s>>>(g, ep);
    else head_pool_kernel<false><<<(unsigned)grid, THREADS, SMEM_BYTES, s>>>(g, ep);
    F8_CUDA(cudaGetLastError());
    if (want_stats) {
        static long long host[16 * 1024];
        F8_CUDA(cudaStreamSynchronize(s));
        F8_CUDA(cudaMemcpy(host, stats_dev, sizeof(host), cudaMemcpyDeviceToHost));
        double acc[16] = {0};
        for (long long b = 0; b < grid; ++b)
            for (int k = 0; k < 16; ++k) acc[k] += (double)host[b * 16 + k] / (double)grid;
        fprintf(stderr,
                "[f8 stats] head_pool tiles/cta=%.1f | build total %.0f wait_patch %.0f wait_aempty %.0f | mma total "
                "%.0f wait_acc %.0f wait_a %.0f | epi total %.0f wait_full %.0f phase1 %.0f phase2 %.0f\n",
                (double)a.n * (POOLED / TP) / (double)grid, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6],
                acc[7], acc[8], acc[9]);
    }
    return F8_OK;
}

}  // namespace f8host
