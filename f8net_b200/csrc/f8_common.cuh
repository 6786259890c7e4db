// f8_common.cuh -- device helpers shared by every kernel of libf8b200.so.
//
// The arithmetic here is the bit-exact device form of the reference's integer rules:
//   requant()      int_op_only_fix_quant            /root/reference/models/fix_quant_ops.py:90-114
//   residual step  IntBlock.forward shift/add/clamp /root/reference/models/fix_resnet.py:40-76,
//                                                   /root/reference/models/fix_mobilenet_v2.py:34-48
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

// float32 -> int32 as the reference's x86 CPU path does (.int()): truncate, and the x86
// "integer indefinite" 0x80000000 when out of range.
__device__ __forceinline__ int32_t f2i_x86(float f) {
    if (!(f >= -2147483648.0f && f < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(f);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

}  // namespace f8

// host-side error plumbing (plan.cu)
namespace f8host {
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
}  // namespace f8host

#define F8_CUDA(call)                                                     \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) return f8host::cuda_fail(_e, #call);       \
    } while (0)

// kernel launchers implemented in the .cu files, called by plan.cu
namespace f8host {
int launch_conv_mma(const f8_conv_args &a, cudaStream_t s);
int launch_conv_umma(const f8_conv_args &a, cudaStream_t s);
int launch_conv3x3_umma(const f8_conv_args &a, cudaStream_t s);
int launch_dw3x3(const f8_conv_args &a, cudaStream_t s);
int launch_maxpool(const f8_conv_args &a, cudaStream_t s);
int launch_pool_requant(const f8_conv_args &a, cudaStream_t s);
int launch_convert_input(const int32_t *x, void *out, int n, int h, int w, int is_signed,
                         cudaStream_t s);
int launch_requant_i32(const int32_t *x, int32_t *y, size_t count, int shift, int is_signed,
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
