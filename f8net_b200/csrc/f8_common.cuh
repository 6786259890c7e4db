// f8_common.cuh -- device helpers shared by every kernel of libf8b200.so.
//
// The arithmetic here is the bit-exact device form of the reference's integer rules:
//   requant()      int_op_only_fix_quant            /root/reference/models/fix_quant_ops.py:90-114
//   residual step  IntBlock.forward shift/add/clamp /root/reference/models/fix_resnet.py:40-76,
//                                                   /root/reference/models/fix_mobilenet_v2.py:34-48
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/f8b200.h"

namespace f8 {

// round-half-to-even right shift (n > 0) or wrapping left shift (n <= 0), then saturate.
//   n > 0 : t = x + 2^(n-1) (wraps);  tie <=> (x mod 2^n) == 2^(n-1);
//           tie ? ((t >> (n+1)) << 1) : (t >> n)
// ((t >> (n+1)) << 1) == (t >> n) & ~1, which is what is computed below.
__device__ __forceinline__ int32_t requant(int32_t x, int n, int is_signed) {
    int32_t r;
    if (n > 0) {
        const uint32_t half = 1u << (n - 1);
        const uint32_t mask = (half << 1) - 1u;
        const int32_t t = (int32_t)((uint32_t)x + half);
        const bool tie = ((uint32_t)x & mask) == half;
        r = t >> n;
        if (tie) r &= ~1;
    } else {
        r = (int32_t)((uint32_t)x << (-n));
    }
    const int lo = is_signed ? -127 : 0;
    const int hi = is_signed ? 127 : 255;
    return max(lo, min(hi, r));
}

// requant() without the final clamp: the saturating byte pack below supplies it.
//   tie <=> (x mod 2^n) == 2^(n-1) <=> the low n bits of t = x + 2^(n-1) are all zero
__device__ __forceinline__ int32_t requant_shift(int32_t x, int n) {
    if (n > 0) {
        const uint32_t half = 1u << (n - 1);
        const uint32_t mask = (half << 1) - 1u;
        const uint32_t t = (uint32_t)x + half;
        int32_t r = (int32_t)t >> n;
        if ((t & mask) == 0u) r &= ~1;
        return r;
    }
    return (int32_t)((uint32_t)x << (-n));
}

// four requantised values -> four saturated bytes (a in the low byte).  cvt.pack.sat clamps to
// [0,255] (u8) or [-128,127] (s8); the reference's signed range is [-127,127], hence the max.
__device__ __forceinline__ uint32_t requant_pack4(int32_t a, int32_t b, int32_t c, int32_t d, int n,
                                                  int is_signed) {
    a = requant_shift(a, n); b = requant_shift(b, n);
    c = requant_shift(c, n); d = requant_shift(d, n);
    uint32_t hi, out;
    if (is_signed) {
        a = max(a, -127); b = max(b, -127); c = max(c, -127); d = max(d, -127);
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(b), "r"(a), "r"(hi));
    } else {
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(b), "r"(a), "r"(hi));
    }
    return out;
}

// int32 carry tensors (residual carries, max-pool / avg-pool inputs) use a pixel-interleaved
// layout private to the engine: with p = pixel index inside the launch (image-major NHW order)
// and C = cout_pad, element (p, c) lives at
//     ((p >> 7) * (C >> 2) + (c >> 2)) * 512 + (p & 127) * 4 + (c & 3)
// i.e. blocks of 128 pixels x 4 channels.  The tensor-core epilogues own one pixel per thread
// (TMEM lane = pixel), so a warp's 16-byte accesses to 32 consecutive pixels are one contiguous
// 512-byte run instead of 32 scattered pieces of an NHWC row.
__host__ __device__ __forceinline__ size_t carry_off(size_t p, int c, int C) {
    return ((p >> 7) * (size_t)(C >> 2) + (size_t)(c >> 2)) * 512 + (p & 127) * 4 + (size_t)(c & 3);
}

// Epilogue parameters, identical for every producing kernel (see f8_op in f8b200.h).
struct Epilogue {
    const int32_t *bias;      // [cout_pad]
    const int32_t *carry_in;  // int32 [M, cout_pad] or nullptr
    int32_t *carry_out;       // int32 [M, cout_pad] or nullptr
    uint8_t *out0;            // 8-bit [M, cout_pad] or nullptr
    uint8_t *out1;
    float *out_f32;           // [M, out_f32_ld] or nullptr
    int out_f32_ld;
    int carry_shift;
    int relu;
    int shift0, signed0;
    int shift1, signed1;
    int cout;                 // logical channel count (float output bound)
    int cout_pad;             // row pitch of the NHWC outputs
    int int_pool;             // max-pool kernels: FXQMaxPool2d (integer max, no float round trip)
};

// acc (already including bias) + optional residual carry -> int32 value every output derives from
__device__ __forceinline__ int32_t residual_relu(int32_t v, bool has_carry, int32_t carry,
                                                 int carry_shift, int relu) {
    if (has_carry) {
        if (carry_shift >= 0) carry = (int32_t)((uint32_t)carry << carry_shift);
        else v = (int32_t)((uint32_t)v << (-carry_shift));
        v = (int32_t)((uint32_t)v + (uint32_t)carry);
        v = max(v, -2147483647);        // clamp_(min=-(1<<31)+1); the max bound is a no-op
    }
    if (relu) v = max(v, 0);
    return v;
}

// streaming 16-byte load of a residual carry (read once: keep it out of L1)
__device__ __forceinline__ int4 ld_stream_int4(const int32_t *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16-byte read-only load of a residual carry through L1 (adjacent 16-byte pieces of a sector
// are requested by the same thread back to back)
__device__ __forceinline__ int4 ld_carry_int4(const int32_t *p) {
    return __ldg(reinterpret_cast<const int4 *>(p));
}

// The fused epilogue for 16 consecutive output channels of ONE output pixel held by one
// thread (tcgen05 kernels: TMEM lane = pixel).  v: raw accumulators; bias16: 16 ints (shared
// memory); o = pixel * cout_pad + first channel; gc = first channel.  carry16: the 16 residual
// carry values, loaded by the caller ahead of time (software prefetch) when PRELOADED.
template <bool PRELOADED>
__device__ __forceinline__ void epilogue16_t(int32_t (&v)[16], const int32_t *bias16,
                                             const Epilogue &ep, size_t o, int gc, size_t pixel,
                                             const int4 *carry16) {
    const bool has_carry = ep.carry_in != nullptr;
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
        const int4 b = *reinterpret_cast<const int4 *>(bias16 + q);
        v[q + 0] = (int32_t)((uint32_t)v[q + 0] + (uint32_t)b.x);
        v[q + 1] = (int32_t)((uint32_t)v[q + 1] + (uint32_t)b.y);
        v[q + 2] = (int32_t)((uint32_t)v[q + 2] + (uint32_t)b.z);
        v[q + 3] = (int32_t)((uint32_t)v[q + 3] + (uint32_t)b.w);
        if (has_carry) {
            int4 c;
            if (PRELOADED) c = carry16[q >> 2];
            else c = ld_stream_int4(ep.carry_in + carry_off(pixel, gc + q, ep.cout_pad));
            v[q + 0] = residual_relu(v[q + 0], true, c.x, ep.carry_shift, ep.relu);
            v[q + 1] = residual_relu(v[q + 1], true, c.y, ep.carry_shift, ep.relu);
            v[q + 2] = residual_relu(v[q + 2], true, c.z, ep.carry_shift, ep.relu);
            v[q + 3] = residual_relu(v[q + 3], true, c.w, ep.carry_shift, ep.relu);
        } else if (ep.relu) {
            v[q + 0] = max(v[q + 0], 0); v[q + 1] = max(v[q + 1], 0);
            v[q + 2] = max(v[q + 2], 0); v[q + 3] = max(v[q + 3], 0);
        }
        if (ep.carry_out)
            *reinterpret_cast<int4 *>(ep.carry_out + carry_off(pixel, gc + q, ep.cout_pad)) =
                make_int4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    }
    if (ep.out0) {
        uint4 w;
        w.x = requant_pack4(v[0], v[1], v[2], v[3], ep.shift0, ep.signed0);
        w.y = requant_pack4(v[4], v[5], v[6], v[7], ep.shift0, ep.signed0);
        w.z = requant_pack4(v[8], v[9], v[10], v[11], ep.shift0, ep.signed0);
        w.w = requant_pack4(v[12], v[13], v[14], v[15], ep.shift0, ep.signed0);
        *reinterpret_cast<uint4 *>(ep.out0 + o) = w;
    }
    if (ep.out1) {
        uint4 w;
        w.x = requant_pack4(v[0], v[1], v[2], v[3], ep.shift1, ep.signed1);
        w.y = requant_pack4(v[4], v[5], v[6], v[7], ep.shift1, ep.signed1);
        w.z = requant_pack4(v[8], v[9], v[10], v[11], ep.shift1, ep.signed1);
        w.w = requant_pack4(v[12], v[13], v[14], v[15], ep.shift1, ep.signed1);
        *reinterpret_cast<uint4 *>(ep.out1 + o) = w;
    }
    if (ep.out_f32) {
        float *f = ep.out_f32 + pixel * ep.out_f32_ld + gc;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (gc + i < ep.cout) f[i] = (float)v[i];
    }
}
__device__ __forceinline__ void epilogue16(int32_t (&v)[16], const int32_t *bias16,
                                           const Epilogue &ep, size_t o, int gc, size_t pixel) {
    epilogue16_t<false>(v, bias16, ep, o, gc, pixel, nullptr);
}

// The arithmetic of the epilogue alone (no memory traffic) for 16 channels of one pixel, written
// branch-free on warp-uniform constants so that an element costs ~10 integer instructions:
//   v += bias;  [v = (v << sv) + (carry << sc), clamp]  ;  v = max(v, floor)
// where floor folds the residual clamp (INT_MIN+1) and the ReLU (0) into one max.
struct EpiConst {
    int sv, sc;        // residual alignment: v <<= sv ; carry <<= sc   (one of them is 0)
    uint32_t mv, mc;   // the same as wrapping multipliers 2^sv, 2^sc: v * mv + carry * mc is two multiply-adds
                       // (FMA pipe) instead of two shifts and an add on the ALU pipe the requantisation lives on
    int floor;         // lower bound applied after bias / residual: 0 (ReLU), INT_MIN+1 (clamp) or INT_MIN
};
__device__ __forceinline__ EpiConst epi_const(const Epilogue &ep, bool has_carry) {
    EpiConst k;
    k.sv = has_carry && ep.carry_shift < 0 ? -ep.carry_shift : 0;
    k.sc = has_carry && ep.carry_shift > 0 ? ep.carry_shift : 0;
    k.mv = 1u << k.sv;          // shifts are at most 30 (check_shift, plan.cu)
    k.mc = 1u << k.sc;
    k.floor = ep.relu ? 0 : (has_carry ? -2147483647 : (int)0x80000000);
    return k;
}
__device__ __forceinline__ void epilogue16_math(int32_t (&v)[16], const int32_t *bias16,
                                                const EpiConst &k, const int4 *c, bool has_carry) {
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
        const int4 b = *reinterpret_cast<const int4 *>(bias16 + q);
        uint32_t t0 = (uint32_t)v[q + 0] + (uint32_t)b.x, t1 = (uint32_t)v[q + 1] + (uint32_t)b.y;
        uint32_t t2 = (uint32_t)v[q + 2] + (uint32_t)b.z, t3 = (uint32_t)v[q + 3] + (uint32_t)b.w;
        if (has_carry) {                       // warp-uniform
            const int4 cc = c[q >> 2];
            t0 = t0 * k.mv + (uint32_t)cc.x * k.mc;
            t1 = t1 * k.mv + (uint32_t)cc.y * k.mc;
            t2 = t2 * k.mv + (uint32_t)cc.z * k.mc;
            t3 = t3 * k.mv + (uint32_t)cc.w * k.mc;
        }
        v[q + 0] = max((int32_t)t0, k.floor); v[q + 1] = max((int32_t)t1, k.floor);
        v[q + 2] = max((int32_t)t2, k.floor); v[q + 3] = max((int32_t)t3, k.floor);
    }
}
// 16 values -> 16 requantised, saturated bytes; the shift direction and signedness are
// warp-uniform and tested once per call
__device__ __forceinline__ uint4 requant_pack16(const int32_t (&v)[16], int n, int is_signed) {
    int32_t r[16];
    if (n > 0) {
        const uint32_t half = 1u << (n - 1);
        const uint32_t mask = (half << 1) - 1u;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t t = (uint32_t)v[i] + half;
            r[i] = (int32_t)t >> n;
            if ((t & mask) == 0u) r[i] &= ~1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = (int32_t)((uint32_t)v[i] << (-n));
    }
    uint32_t w[4];
    if (is_signed) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t hi;
            const int32_t a = max(r[4 * q], -127), b = max(r[4 * q + 1], -127);
            const int32_t c = max(r[4 * q + 2], -127), d = max(r[4 * q + 3], -127);
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w[q]) : "r"(b), "r"(a), "r"(hi));
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t hi;
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r[4 * q + 3]), "r"(r[4 * q + 2]), "r"(0));
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w[q]) : "r"(r[4 * q + 1]), "r"(r[4 * q]), "r"(hi));
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Fast path of the same epilogue for the most common launch: no residual, no int32 carry
// out, ONE unsigned 8-bit consumer reached by a right shift (n > 0).  ReLU is then absorbed
// by the unsigned clamp (requant(relu(v)) == requant(v) for n > 0, SURVEY.md "exactness
// traps"), and bias + 2^(n-1) is one pre-merged addend (wrapping adds associate), so an
// element costs add, tie test, shift, tie mask and half a saturating pack.
// bh16: 16 ints in shared memory holding bias + half.
__host__ __device__ inline bool epilogue_is_plain_u8(const Epilogue &ep) {
    return ep.carry_in == nullptr && ep.carry_out == nullptr && ep.out1 == nullptr &&
           ep.out_f32 == nullptr && ep.out0 != nullptr && ep.shift0 > 0 && !ep.signed0;
}
__device__ __forceinline__ void epilogue16_plain_u8(const int32_t (&v)[16], const int32_t *bh16,
                                                    uint8_t *dst, int n) {
    const uint32_t mask = (1u << n) - 1u;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int4 b = *reinterpret_cast<const int4 *>(bh16 + 4 * q);
        int32_t r[4];
        const uint32_t t0 = (uint32_t)v[4 * q + 0] + (uint32_t)b.x;
        const uint32_t t1 = (uint32_t)v[4 * q + 1] + (uint32_t)b.y;
        const uint32_t t2 = (uint32_t)v[4 * q + 2] + (uint32_t)b.z;
        const uint32_t t3 = (uint32_t)v[4 * q + 3] + (uint32_t)b.w;
        r[0] = (int32_t)t0 >> n; if ((t0 & mask) == 0u) r[0] &= ~1;
        r[1] = (int32_t)t1 >> n; if ((t1 & mask) == 0u) r[1] &= ~1;
        r[2] = (int32_t)t2 >> n; if ((t2 & mask) == 0u) r[2] &= ~1;
        r[3] = (int32_t)t3 >> n; if ((t3 & mask) == 0u) r[3] &= ~1;
        uint32_t hi;
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(r[3]), "r"(r[2]), "r"(0));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w[q]) : "r"(r[1]), "r"(r[0]), "r"(hi));
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// float32 -> int32 as the reference's x86 CPU path does (.int()): truncate, and the x86
// "integer indefinite" 0x80000000 when out of range.
__device__ __forceinline__ int32_t f2i_x86(float f) {
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(f);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Programmatic dependent launch (griddepcontrol): a kernel launched through f8host::launch_pdl may
// start -- barrier init, TMEM allocation, tensor-map prefetch, weight-ring fill -- while the
// previous layer's launch is still draining; pdl_wait() returns once that launch has completed
// and its writes are visible, so it must precede the first access to any activation / carry /
// logits buffer.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
#ifdef F8_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

}  // namespace f8

// host-side error plumbing (plan.cu)
namespace f8host {
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
// every launcher names the kernel template it is about to launch (thread local, printf style);
// f8_plan_profile keeps the name per op (f8_plan_kernel_name) so that measurements can be
// attributed to kernel templates instead of op kinds
void note_kernel(const char *fmt, ...);
}  // namespace f8host

#define F8_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return f8host::cuda_fail(_e, #call);       \
    } while (0)

// Debug instrumentation (in-kernel wait counters F8_STATS, timing probes F8_PROBE / F8_RPROBE that
// skip work and therefore give WRONG results) exists only in a -DF8_DEBUG_PROBES build
// (F8_DEBUG_PROBES=1 python -m f8net_b200.build --force).  In the shipping library the switches
// are not read at all and every probe branch is compiled out.
#ifdef F8_DEBUG_PROBES
#define F8_DBG 1
#else
#define F8_DBG 0
#endif

// kernel launchers implemented in the .cu files, called by plan.cu
namespace f8host {
inline const char *debug_env(const char *name) { return F8_DBG ? getenv(name) : nullptr; }

// One-time setup of a kernel family PER DEVICE: cudaFuncSetAttribute (the > 48 KB dynamic shared
// memory opt-in) applies to the current device only, and so does the SM count a persistent grid is
// sized with -- a process may hold plans on several GPUs and launch from several host threads.
struct DeviceOnce {
    static constexpr int kMaxDevices = 64;
    std::mutex m;
    bool done[kMaxDevices] = {};
    int sms[kMaxDevices] = {};
};
template <typename Setup>
inline int device_once(DeviceOnce &st, int *num_sms, Setup &&setup) {
    int dev = 0;
    F8_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= DeviceOnce::kMaxDevices) {
        set_error("device index %d outside the supported range", dev);
        return F8_ERR_UNSUPPORTED;
    }
    std::lock_guard<std::mutex> lk(st.m);
    if (!st.done[dev]) {
        const int rc = setup();
        if (rc) return rc;
        F8_CUDA(cudaDeviceGetAttribute(&st.sms[dev], cudaDevAttrMultiProcessorCount, dev));
        st.done[dev] = true;
    }
    *num_sms = st.sms[dev];
    return F8_OK;
}

// F8_PDL=0 turns programmatic dependent launch off (plain stream order between layers)
inline bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("F8_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
// Launch with the programmatic-stream-serialization attribute: only for kernels whose every
// access to the previous launch's output comes after f8::pdl_wait().
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster` CTAs along x (1 = no cluster attribute)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                      int cluster, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = (unsigned)cluster;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = cluster > 1 ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
int launch_conv_mma(const f8_conv_args &a, cudaStream_t s);
int launch_conv_umma(const f8_conv_args &a, cudaStream_t s);
int launch_conv3x3_umma(const f8_conv_args &a, cudaStream_t s);
int launch_head_pool(const f8_conv_args &a, cudaStream_t s);
int launch_head3x3s2(const f8_conv_args &a, cudaStream_t s);
int launch_dw3x3(const f8_conv_args &a, cudaStream_t s);
int launch_conv3x3_dw(const f8_conv_args &a, cudaStream_t s);
// depthwise weight pack = [12][cpad/4] dp4a words, then (256-byte aligned) one block-diagonal
// dense image [36 chunks][64 rows][16 B] per 64-channel group for the tensor-core path
inline size_t dw_dense_offset(int cpad) { return ((size_t)12 * (size_t)cpad + 255) / 256 * 256; }
inline size_t dw_pack_bytes(int cpad) { return dw_dense_offset(cpad) + (size_t)((cpad + 63) / 64) * 36 * 64 * 16; }
int launch_maxpool(const f8_conv_args &a, cudaStream_t s);
int launch_pool_requant(const f8_conv_args &a, cudaStream_t s);
int launch_pool_fc(const f8_conv_args &a, cudaStream_t s);
// range_flag: host-mapped word raised when a value lies outside the head's 8 bits ([-128,127] signed,
// [0,255] unsigned); nullptr = no check
int launch_convert_input(const int32_t *x, void *out, int n, int h, int w, int is_signed,
                         cudaStream_t s, int *range_flag = nullptr);
int launch_requant_i32(const int32_t *x, int32_t *y, size_t count, int shift, int is_signed,
                       cudaStream_t s);
int launch_integerize_f32(const float *x, void *out, int n, int h, int w, int normalize, int fraclen,
                          cudaStream_t s, int is_signed = 0, int *range_flag = nullptr);
int launch_integerize_u8(const uint8_t *x, const uint8_t *lut_dev, void *out, int n, int h, int w,
                         cudaStream_t s);

// Dense weight image geometry (shared by the packer and the kernels)
struct DensePack {
    int mode;        // 0: K = (r*kw+s)*cin_pad + c ; 1: small-C row-window mode (cin_pad == 4)
    int row_bytes;   // mode 1: bytes of one filter row window
    int shift_px;    // mode 1: extra pixels on the left of the window
    int K;           // logical K in bytes
    int K_pad;       // multiple of 64
    int rows;        // cout_pad rounded up to 256
    // image: int8 [K_pad/16][rows][16] -- byte k of output row o at ((k/16)*rows + o)*16 + k%16
};
DensePack dense_pack_geometry(int cin_pad, int cout_pad, int kh, int kw);
}  // namespace f8host
