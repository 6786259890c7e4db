// head3x3_umma.cu -- the MobileNet head: 3x3 stride-2 pad-1 int8 convolution of the 4-byte-pixel
// input image (Cin = 3 -> 32 channels) + bias + ReLU + requant on tcgen05, without im2col.
//
// Replaces, per launch: head[0] = int nn.Conv2d(3, 32, 3, 2, 1) built by int_conv()
// (/root/reference/models/fix_quant_ops.py:680-714) and the ReLU + consumer-side
// int_op_only_fix_quant that follow it in IntModel.forward
// (/root/reference/models/fix_mobilenet_v1.py:120-131, fix_mobilenet_v2.py:207-218).
//
// Space-to-depth view.  Group the NHWC4 image into 2x2 pixel blocks: packed pixel (P, Q) is the
// 16 bytes  [px(2P,2Q) px(2P,2Q+1) | px(2P+1,2Q) px(2P+1,2Q+1)].  Output pixel (p, q) of the
// stride-2 convolution reads image rows 2p-1..2p+1 and columns 2q-1..2q+1, i.e. the four packed
// pixels (p-1|p, q-1|q): a 2x2 STRIDE-1 convolution over 16-byte "channels", K = 4 x 16 = 64 bytes,
// with zero weights on the sub-positions a tap does not touch.
//
// Padded linear space (as in conv3x3_umma.cu): one zero packed row above every image and one zero
// packed column before every row, pitch PW = Wout + 1:
//     slot   s(img, P, Q) = (img*(Hout+1) + P + 1)*PW + Q + 1          (16 bytes each)
//     output m(img, p, q) = (img*(Hout+1) + p    )*PW + q
// so output m reads slots m, m+1 (packed row p-1) and m+PW, m+PW+1 (packed row p).  With the
// patch resident in shared memory as [slot][16 B], 128 consecutive outputs are a canonical
// K-major operand whose core matrices are 8 consecutive slots (SBO = 128 B) and whose second
// K chunk is simply the NEXT slot (LBO = 16 B, overlapping core matrices): one M128 x N32 x K32
// MMA per packed row, two per 128 outputs, eight per 512-output tile -- against 32 cp.async
// gathers per output pixel in the generic small-C path.
//
// Persistent CTA, one per SM: 4 loader warps (two 8-byte cp.async per slot: the two image rows of
// a packed pixel), 1 MMA warp, 16 epilogue warps (warp = 32 outputs x 32 channels); 2 accumulator
// sets x 4 segments x 32 columns = 256 TMEM columns.  The 2 KB weight tile is built once per CTA
// from the layer's row-window weight image (K order of plan.cu's mode-1 pack).
#include <cstdlib>
#include <cstring>

#include "umma_common.cuh"

namespace {

using namespace f8u;

constexpr int COUT_PAD = 32;
constexpr int TM = 512;                    // outputs per tile: 4 segments of 128
constexpr int SEGS = TM / 128;
constexpr int SA = 4;                      // patch ring depth
constexpr int LAG = 2;                     // a stage is signalled once LAG younger ones are issued
constexpr int EPI_WARPS = 16, LOAD_WARPS = 4;
constexpr int LOADERS = LOAD_WARPS * 32;
constexpr int THREADS = (EPI_WARPS + LOAD_WARPS + 1) * 32;
constexpr int B_BYTES = 4 * COUT_PAD * 16; // [4 chunks (dP, dQ)][32 rows][16 B]

struct H3Geom {
    const uint8_t *in;      // NHWC4 8-bit [N, Hin, Win, 4]
    const uint8_t *wpack;   // mode-1 dense pack: k = r*row_bytes + (s + shift_px)*4 + c
    int wrows, row_bytes, shift_px;
    int N, Hin, Win, Hout, Wout;
    int PW, HP;             // Wout + 1, Hout + 1
    uint32_t mPW, mHP;      // floor(2^32 / d) + 1
    int stage_slots;        // TM + PW + 2 rounded up to 8
    int ntiles;
};

template <bool A_SIGNED>
__global__ void __launch_bounds__(THREADS, 1)
head3x3s2_kernel(const H3Geom g, const f8::Epilogue ep) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (f8::smem_u32(smem_raw) & 127u)) & 127u);
    const int stage_bytes = g.stage_slots * 16;
    const uint32_t smem_base = f8::smem_u32(smem);
    const uint32_t b_base = smem_base + SA * stage_bytes;
    const uint32_t bar_base = b_base + B_BYTES;
    auto full_bar = [&](int s) { return bar_base + (uint32_t)s * 8; };
    auto empty_bar = [&](int s) { return bar_base + (uint32_t)(SA + s) * 8; };
    auto acc_full = [&](int b) { return bar_base + (uint32_t)(2 * SA + b) * 8; };
    auto acc_empty = [&](int b) { return bar_base + (uint32_t)(2 * SA + 2 + b) * 8; };
    uint8_t *after = smem + SA * stage_bytes + B_BYTES + (2 * SA + 4) * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(after);
    int32_t *sbias = reinterpret_cast<int32_t *>(after + 16);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int MMA_WARP = EPI_WARPS + LOAD_WARPS;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < SA; ++s) { mbar_init(full_bar(s), LOAD_WARPS); mbar_init(empty_bar(s), 1); }
            for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), EPI_WARPS); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(f8::smem_u32(tmem_slot), 256);
    }
    // weight tile: chunk j = (dP, dQ), byte b = (sub-row, sub-column, channel) of the packed pixel
    // -> filter tap r = 2 dP + sub-row - 1, s = 2 dQ + sub-column - 1 (zero when outside the 3x3)
    for (int idx = tid; idx < B_BYTES; idx += THREADS) {
        const int j = idx >> 9, o = (idx >> 4) & 31, b = idx & 15;
        const int r = 2 * (j >> 1) + (b >> 3) - 1, s = 2 * (j & 1) + ((b >> 2) & 1) - 1;
        int8_t w = 0;
        if (r >= 0 && s >= 0 && o < ep.cout_pad) {
            const int k = r * g.row_bytes + (s + g.shift_px) * 4 + (b & 3);
            w = (int8_t)__ldg(g.wpack + ((size_t)(k >> 4) * g.wrows + o) * 16 + (k & 15));
        }
        smem[SA * stage_bytes + idx] = (uint8_t)w;
    }
    if (tid < COUT_PAD)      // plain path: bias + half (f8_common.cuh epilogue16_plain_u8)
        sbias[tid] = (int32_t)((uint32_t)__ldg(ep.bias + tid) + (1u << (ep.shift0 - 1)));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    f8::pdl_trigger();
    f8::pdl_wait();                     // the image is the previous launch's output

    if (warp >= EPI_WARPS && warp < MMA_WARP) {
        // =========================== patch loaders ================================
        const int lt = tid - EPI_WARPS * 32;
        const size_t row_b = (size_t)g.Win * 4;
        int slot = 0, phase = 0, aslot = 0, issued = 0;
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            mbar_wait(empty_bar(slot), phase ^ 1);
            const uint32_t sa = smem_base + slot * stage_bytes;
            const int s0 = t * TM;
            for (int i = lt; i < g.stage_slots; i += LOADERS) {
                const int sg = s0 + i;
                const int Y = (int)__umulhi((uint32_t)sg, g.mPW);
                const int X = sg - Y * g.PW;
                const int img = (int)__umulhi((uint32_t)Y, g.mHP);
                const int P = Y - img * g.HP - 1, Q = X - 1;
                const bool ok = img < g.N && P >= 0 && Q >= 0;
                const uint8_t *src = ok ? g.in + ((size_t)(img * g.Hin + 2 * P) * g.Win + 2 * Q) * 4 : g.in;
                cp_async8(sa + i * 16, src, ok);
                cp_async8(sa + i * 16 + 8, src + (ok ? row_b : 0), ok);
            }
            cp_async_commit();
            if (++slot == SA) { slot = 0; phase ^= 1; }
            if (++issued > LAG) {
                cp_async_wait<LAG>();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(aslot));
                if (++aslot == SA) aslot = 0;
                --issued;
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        for (; issued > 0; --issued) {
            if (lane == 0) mbar_arrive(full_bar(aslot));
            if (++aslot == SA) aslot = 0;
        }
    } else if (warp == MMA_WARP) {
        // =========================== MMA issuer ===================================
        constexpr uint32_t idesc = instr_desc(A_SIGNED, COUT_PAD);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, version 1
        constexpr uint32_t a_lbo = (16u >> 4) << 16;                            // next K chunk = next slot
        constexpr uint32_t b_lbo = ((uint32_t)(COUT_PAD * 16) >> 4) << 16;
        const uint32_t b_lo = ((b_base & 0x3ffffu) >> 4) | b_lbo;
        int slot = 0, phase = 0, buf = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            mbar_wait(acc_empty(buf), acc_phase ^ 1);
            mbar_wait(full_bar(slot), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + slot * stage_bytes;
            const uint32_t a_lo = ((sa & 0x3ffffu) >> 4) | a_lbo;
            const uint32_t tacc = tmem_base + (uint32_t)(buf * SEGS * COUT_PAD);
            if (elect_one()) {
#pragma unroll
                for (int sg = 0; sg < SEGS; ++sg) {
                    // packed row p-1: slots m, m+1; packed row p: slots m+PW, m+PW+1
                    umma_i8_lohi(tacc + (uint32_t)(sg * COUT_PAD), a_lo + (uint32_t)(sg * 128), desc_hi, b_lo, desc_hi,
                                 idesc, 0u);
                    umma_i8_lohi(tacc + (uint32_t)(sg * COUT_PAD), a_lo + (uint32_t)(sg * 128 + g.PW), desc_hi,
                                 b_lo + (uint32_t)((2 * COUT_PAD * 16) >> 4), desc_hi, idesc, 1u);
                }
                umma_commit(empty_bar(slot));
                umma_commit(acc_full(buf));
            }
            __syncwarp();
            if (++slot == SA) { slot = 0; phase ^= 1; }
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
    } else {
        // =========================== epilogue (16 warps) ==========================
        const int lg = warp & 3, seg = warp >> 2;
        int buf = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
            const int m = t * TM + seg * 128 + lg * 32 + lane;
            const int Y = (int)__umulhi((uint32_t)m, g.mPW);
            const int q = m - Y * g.PW;
            const int img = (int)__umulhi((uint32_t)Y, g.mHP);
            const int p = Y - img * g.HP;
            const bool valid = q < g.Wout && p < g.Hout && img < g.N;
            uint8_t *dst = ep.out0 + ((size_t)(img * g.Hout + p) * g.Wout + q) * COUT_PAD;
            mbar_wait(acc_full(buf), acc_phase);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * SEGS + seg) * COUT_PAD);
            int32_t v0[16], v1[16];
            tmem_ld16(trow, v0);
            tmem_ld16(trow + 16, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));     // this warp's columns are in registers
            if (valid) {
                f8::epilogue16_plain_u8(v0, sbias, dst, ep.shift0);
                if (ep.cout_pad > 16) f8::epilogue16_plain_u8(v1, sbias + 16, dst + 16, ep.shift0);
            }
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

namespace f8host {

// F8_ERR_UNSUPPORTED => the caller takes the generic small-C path of conv_umma.cu.
int launch_head3x3s2(const f8_conv_args &a, cudaStream_t s) {
    static const bool off = getenv("F8_NO_HEAD3X3") != nullptr;
    if (off || a.kh != 3 || a.kw != 3 || a.stride != 2 || a.pad != 1 || a.cin_pad != 4 || a.cout_pad != COUT_PAD ||
        (a.hin & 1) || (a.win & 1) || a.hout != a.hin / 2 || a.wout != a.win / 2)
        return F8_ERR_UNSUPPORTED;
    f8::Epilogue ep{};
    ep.bias = a.bias;
    ep.carry_in = a.carry_in;
    ep.carry_out = a.carry_out;
    ep.out0 = static_cast<uint8_t *>(a.out[0]);
    ep.out1 = static_cast<uint8_t *>(a.out[1]);
    ep.out_f32 = a.out_f32;
    ep.carry_shift = a.carry_shift;
    ep.relu = a.relu;
    ep.shift0 = a.out_shift[0]; ep.signed0 = a.out_signed[0];
    ep.shift1 = a.out_shift[1]; ep.signed1 = a.out_signed[1];
    ep.cout = a.cout;
    ep.cout_pad = a.cout_pad;
    if (!f8::epilogue_is_plain_u8(ep)) return F8_ERR_UNSUPPORTED;     // bias + ReLU + one unsigned right-shift consumer
    const DensePack pk = dense_pack_geometry(a.cin_pad, a.cout_pad, a.kh, a.kw);
    if (pk.mode != 1) return F8_ERR_UNSUPPORTED;
    H3Geom g{};
    g.in = static_cast<const uint8_t *>(a.in);
    g.wpack = static_cast<const uint8_t *>(a.wpack);
    g.wrows = pk.rows; g.row_bytes = pk.row_bytes; g.shift_px = pk.shift_px;
    g.N = a.n; g.Hin = a.hin; g.Win = a.win; g.Hout = a.hout; g.Wout = a.wout;
    g.PW = a.wout + 1; g.HP = a.hout + 1;
    g.mPW = (uint32_t)(0x100000000ULL / (uint32_t)g.PW) + 1u;
    g.mHP = (uint32_t)(0x100000000ULL / (uint32_t)g.HP) + 1u;
    g.stage_slots = (TM + g.PW + 2 + 7) / 8 * 8;
    const long long lin = (long long)a.n * g.HP * g.PW;
    // the magic-number divisions need (lin + stage) * max(PW, HP) < 2^32
    if ((lin + g.stage_slots + TM) * (long long)(g.PW > g.HP ? g.PW : g.HP) >= 0xffffffffLL) return F8_ERR_UNSUPPORTED;
    g.ntiles = (int)((lin + TM - 1) / TM);
    size_t smem = (size_t)SA * g.stage_slots * 16 + B_BYTES + (2 * SA + 4) * 8 + 16 + COUT_PAD * 4 + 128;
    if (smem > 200 * 1024) return F8_ERR_UNSUPPORTED;
    if (smem < 120 * 1024) smem = 120 * 1024;                          // one CTA per SM (register budget)
    static DeviceOnce once;
    int num_sms = 0;
    {
        const int rc = device_once(once, &num_sms, []() -> int {
            F8_CUDA(cudaFuncSetAttribute(head3x3s2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            F8_CUDA(cudaFuncSetAttribute(head3x3s2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            return F8_OK;
        });
        if (rc) return rc;
    }
    const unsigned grid = (unsigned)(g.ntiles < num_sms ? g.ntiles : num_sms);
    note_kernel("head3x3s2");
    if (a.in_signed) F8_CUDA(launch_pdl(head3x3s2_kernel<true>, grid, (unsigned)THREADS, smem, s, g, ep));
    else F8_CUDA(launch_pdl(head3x3s2_kernel<false>, grid, (unsigned)THREADS, smem, s, g, ep));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

}  // namespace f8host
