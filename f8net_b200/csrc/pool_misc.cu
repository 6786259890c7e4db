// pool_misc.cu -- the small memory-bound kernels of the int_op_only path:
//   maxpool3x3s2   ResNet head max-pool with the float32 round trip
//                  x = self.head[-1](x.float()).int()   /root/reference/models/fix_resnet.py:358-359
//   pool_requant   FXQAvgPool2d int branch + classifier requant
//                  /root/reference/models/fix_quant_ops.py:126-134, fix_resnet.py:367-374
//   convert_input  int32 NCHW (reference tensor) -> NHWC4 8-bit
//   requant_i32    int_op_only_fix_quant as a standalone op, fix_quant_ops.py:90-114
// All are HBM-bound streaming kernels: 16-byte vector accesses, channel-fastest thread
// mapping so every warp touches contiguous memory.
#include "f8_common.cuh"

namespace {

constexpr int THREADS = 256;

// in: int32 [n,hin,win,cpad] in the carry layout of f8_common.cuh (post-ReLU head accumulator).  One thread = 4 channels
// of one output pixel.  int -> float is monotone, so max-then-convert equals the
// reference's convert-then-max; -inf padding never wins because every window holds at
// least one real element.
__global__ void __launch_bounds__(THREADS)
maxpool_kernel(const int32_t *__restrict__ in, int n, int hin, int win, int hout, int wout,
               int cpad, const f8::Epilogue ep) {
    const int c4n = cpad >> 2;
    const long long total = (long long)n * hout * wout * c4n;
    for (long long idx = blockIdx.x * (long long)THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * THREADS) {
        const int c4 = (int)(idx % c4n);
        long long t = idx / c4n;
        const int q = (int)(t % wout);
        t /= wout;
        const int p = (int)(t % hout);
        const int img = (int)(t / hout);
        int4 m = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int ih = p * 2 - 1 + r;
            if ((unsigned)ih >= (unsigned)hin) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int iw = q * 2 - 1 + s;
                if ((unsigned)iw >= (unsigned)win) continue;
                const int4 v = __ldg(reinterpret_cast<const int4 *>(
                    in + f8::carry_off(((size_t)img * hin + ih) * win + iw, c4 * 4, cpad)));
                m.x = max(m.x, v.x); m.y = max(m.y, v.y);
                m.z = max(m.z, v.z); m.w = max(m.w, v.w);
            }
        }
        // nn.MaxPool2d on x.float(), then .int() (fix_resnet.py:358-359) -- or FXQMaxPool2d's integer max
        // (fix_quant_ops.py:141-157; its zero padding never wins over the post-ReLU values)
        int32_t v[4] = {m.x, m.y, m.z, m.w};
        if (!ep.int_pool) {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = f8::f2i_x86((float)v[c]);
        }
        const size_t o = (((size_t)img * hout + p) * wout + q) * cpad + c4 * 4;
        if (ep.relu) {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = max(v[c], 0);
        }
        if (ep.carry_out)
            *reinterpret_cast<int4 *>(ep.carry_out + f8::carry_off(o / cpad, c4 * 4, cpad)) =
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

// in: int32 [n, hw, cpad] in the carry layout -> out0: 8-bit [n, cpad].  Sum wraps mod 2^32,
// which equals the reference's int64 sum followed by .int() (fix_quant_ops.py:130-133).
// One thread = 4 channels of one image; consecutive pixels of a channel quad are 16 B apart.
__global__ void __launch_bounds__(THREADS)
pool_requant_kernel(const int32_t *__restrict__ in, int n, int hw, int cpad,
                    const f8::Epilogue ep) {
    const int c4n = cpad >> 2;
    const long long total = (long long)n * c4n;
    for (long long idx = blockIdx.x * (long long)THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * THREADS) {
        const int c4 = (int)(idx % c4n);
        const int img = (int)(idx / c4n);
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int i = 0; i < hw; ++i) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(
                in + f8::carry_off((size_t)img * hw + i, c4 * 4, cpad)));
            a0 += (uint32_t)v.x; a1 += (uint32_t)v.y; a2 += (uint32_t)v.z; a3 += (uint32_t)v.w;
        }
        const int32_t v[4] = {(int32_t)a0, (int32_t)a1, (int32_t)a2, (int32_t)a3};
        const size_t o = (size_t)img * cpad + c4 * 4;
        if (ep.carry_out)       // plain [n, cpad] int32 (tests only)
            *reinterpret_cast<int4 *>(ep.carry_out + o) = make_int4(v[0], v[1], v[2], v[3]);
        if (ep.out0) {
            uint32_t pk = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                pk |= ((uint32_t)f8::requant(v[c], ep.shift0, ep.signed0) & 0xffu) << (8 * c);
            *reinterpret_cast<uint32_t *>(ep.out0 + o) = pk;
        }
    }
}

// The network tail in one launch: FXQAvgPool2d (sum over the hw pixels, wrapping) + requant to
// 8 bit + the classifier nn.Linear + .float()  (fix_quant_ops.py:126-134, fix_resnet.py:367-383).
// One CTA = TAIL_IMGS images (2, or 4 when the weight matrix is large: fewer re-reads from L2): phase 1 pools and requantises their channel vectors into shared
// memory, phase 2 gives every thread output neurons o = tid, tid + THREADS, ... for all the images
// of the CTA: a weight chunk (16 K bytes of row o, from the dense pack [K/16][rows][16]) is loaded
// once and multiplied with the matching 16 bytes of every image (dp4a, int32 wrap == the reference's
// int32 addmm).  Consecutive threads read consecutive 16-byte weight pieces.
// 1024 threads: one output neuron per thread and twice the weight loads in flight per SM (the phase is bound by
// L2 latency: 27 -> 19 us for ResNet18, 90 -> 57 us for ResNet50 against 512 threads)
constexpr int TAIL_THREADS = 1024;

template <bool Q_SIGNED, int TAIL_IMGS>
__global__ void __launch_bounds__(TAIL_THREADS)
pool_fc_kernel(const int32_t *__restrict__ in, int n, int hw, int cpad, const uint4 *__restrict__ w, int wrows,
               const int32_t *__restrict__ bias, int cout, int shift, float *__restrict__ out, int out_ld) {
    extern __shared__ __align__(16) uint8_t q[];            // [TAIL_IMGS][cpad]
    f8::pdl_trigger();                                      // programmatic dependent launch:
    f8::pdl_wait();                                         // the last conv's carry is read below
    const int img0 = blockIdx.x * TAIL_IMGS;
    const int nimg = min(TAIL_IMGS, n - img0);
    const int c4n = cpad >> 2;
    // phase 1: four adjacent lanes share one (image, channel quad): lane part p sums pixels p, p+4, ...
    // (consecutive pixels of a quad are 16 B apart in the carry layout: the four lanes read 64
    // contiguous bytes), then two shuffles finish the sum
    for (int idx = threadIdx.x; idx < TAIL_IMGS * c4n * 4; idx += TAIL_THREADS) {
        const int item = idx >> 2, part = idx & 3;
        const int li = item / c4n, c4 = item - li * c4n;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (li < nimg) {
#pragma unroll 4
            for (int i = part; i < hw; i += 4) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(
                    in + f8::carry_off((size_t)(img0 + li) * hw + i, c4 * 4, cpad)));
                a0 += (uint32_t)v.x; a1 += (uint32_t)v.y; a2 += (uint32_t)v.z; a3 += (uint32_t)v.w;
            }
        }
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d);
            a2 += __shfl_xor_sync(0xffffffffu, a2, d); a3 += __shfl_xor_sync(0xffffffffu, a3, d);
        }
        if (part == 0) {
            uint32_t pk = 0;
            if (li < nimg)
                pk = ((uint32_t)f8::requant((int32_t)a0, shift, Q_SIGNED) & 0xffu) |
                     (((uint32_t)f8::requant((int32_t)a1, shift, Q_SIGNED) & 0xffu) << 8) |
                     (((uint32_t)f8::requant((int32_t)a2, shift, Q_SIGNED) & 0xffu) << 16) |
                     (((uint32_t)f8::requant((int32_t)a3, shift, Q_SIGNED) & 0xffu) << 24);
            reinterpret_cast<uint32_t *>(q)[item] = pk;
        }
    }
    __syncthreads();
    const int kchunks = cpad >> 4;
    for (int o = threadIdx.x; o < cout; o += TAIL_THREADS) {
        int32_t acc[TAIL_IMGS];
#pragma unroll
        for (int i = 0; i < TAIL_IMGS; ++i) acc[i] = 0;
#pragma unroll 8
        for (int kc = 0; kc < kchunks; ++kc) {
            const uint4 wv = __ldg(w + (size_t)kc * wrows + o);
#pragma unroll
            for (int i = 0; i < TAIL_IMGS; ++i) {
                const uint4 x = *reinterpret_cast<const uint4 *>(q + i * cpad + kc * 16);     // broadcast
                if (Q_SIGNED) {
                    asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.x), "r"(wv.x));
                    asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.y), "r"(wv.y));
                    asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.z), "r"(wv.z));
                    asm("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.w), "r"(wv.w));
                } else {
                    asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.x), "r"(wv.x));
                    asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.y), "r"(wv.y));
                    asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.z), "r"(wv.z));
                    asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x.w), "r"(wv.w));
                }
            }
        }
        const uint32_t b = (uint32_t)__ldg(bias + o);
#pragma unroll
        for (int i = 0; i < TAIL_IMGS; ++i)
            if (i < nimg) out[(size_t)(img0 + i) * out_ld + o] = (float)(int32_t)((uint32_t)acc[i] + b);
    }
}

// x int32 [n,3,h,w] -> out 8-bit [n,h,w,4]; channel 3 = 0.  Keeps the low byte: u8 0..255
// and s8 -128..127 both survive the truncation unchanged.  The reference's head conv consumes the
// full int32 (fix_resnet.py:355): a value outside [lo, lo + 255] would give different logits, so
// it raises *range_flag (host-mapped word of the plan; nullptr = standalone call, no check).
__global__ void __launch_bounds__(THREADS)
convert_input_kernel(const int32_t *__restrict__ x, uint32_t *__restrict__ out, int n, int hw, int lo,
                     int *range_flag) {
    const long long total = (long long)n * hw;
    // programmatic dependent launch: the NHWC4 buffer written here may still be read by the
    // head conv of the previous chunk / pass
    f8::pdl_trigger();
    f8::pdl_wait();
    uint32_t wide = 0;
    for (long long idx = blockIdx.x * (long long)THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * THREADS) {
        const int img = (int)(idx / hw);
        const int px = (int)(idx - (long long)img * hw);
        const int32_t *src = x + (size_t)img * 3 * hw + px;
        const uint32_t v0 = (uint32_t)__ldg(src), v1 = (uint32_t)__ldg(src + hw), v2 = (uint32_t)__ldg(src + 2 * (size_t)hw);
        wide |= (v0 - (uint32_t)lo) | (v1 - (uint32_t)lo) | (v2 - (uint32_t)lo);
        out[idx] = (v0 & 0xffu) | ((v1 & 0xffu) << 8) | ((v2 & 0xffu) << 16);
    }
    if (range_flag && (wide & ~0xffu)) {                       // never taken for well-formed inputs
        *reinterpret_cast<volatile int *>(range_flag) = 1;
        __threadfence_system();
    }
}

// forward_loss's integerisation (fix_train.py:676-692) of the float32 NCHW tensor, in float32
// with round-half-even exactly as torch does: (255 * x).round().int()  or
// clamp(round(x * 2^fl), -127, 127); the low byte is kept and a value outside the head's 8 bits
// (x < 0 -- the reference asserts input >= 0, fix_train.py:689 --, x > 1, NaN) raises *range_flag
// (like the int32 path).
__global__ void __launch_bounds__(THREADS)
integerize_f32_kernel(const float *__restrict__ x, uint32_t *__restrict__ out, int n, int hw, int normalize,
                      float scale, int lo, int *range_flag) {
    const long long total = (long long)n * hw;
    uint32_t wide = 0;
    for (long long idx = blockIdx.x * (long long)THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * THREADS) {
        const int img = (int)(idx / hw);
        const int px = (int)(idx - (long long)img * hw);
        const float *src = x + (size_t)img * 3 * hw + px;
        uint32_t q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float r = rintf(__fmul_rn(__ldg(src + (size_t)c * hw), scale));
            int32_t v;
            if (normalize) v = (int32_t)fminf(fmaxf(r, -127.0f), 127.0f);
            else v = f8::f2i_x86(r);
            // NaN: torch's clamp keeps it and .int() gives INT_MIN; fmaxf drops it -- out of range either way
            wide |= ((uint32_t)v - (uint32_t)lo) | (r != r ? 0x100u : 0u);
            q[c] = (uint32_t)v & 0xffu;
        }
        out[idx] = q[0] | (q[1] << 8) | (q[2] << 16);
    }
    if (range_flag && (wide & ~0xffu)) {
        *reinterpret_cast<volatile int *>(range_flag) = 1;
        __threadfence_system();
    }
}

// decoded image bytes [n, hw, 3] through the per-channel table -> NHWC4
__global__ void __launch_bounds__(THREADS)
integerize_u8_kernel(const uint8_t *__restrict__ x, const uint8_t *__restrict__ lut, uint32_t *__restrict__ out,
                     long long total) {
    __shared__ uint8_t slut[768];
    for (int i = threadIdx.x; i < 768; i += THREADS) slut[i] = lut[i];
    __syncthreads();
    // four pixels (12 bytes = three aligned words) per thread
    const long long quads = total >> 2;
    for (long long qd = blockIdx.x * (long long)THREADS + threadIdx.x; qd < quads;
         qd += (long long)gridDim.x * THREADS) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(x) + qd * 3;
        const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
        const uint32_t b[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
        uint4 o;
        o.x = slut[b[0]] | (slut[256 + b[1]] << 8) | (slut[512 + b[2]] << 16);
        o.y = slut[b[3]] | (slut[256 + b[4]] << 8) | (slut[512 + b[5]] << 16);
        o.z = slut[b[6]] | (slut[256 + b[7]] << 8) | (slut[512 + b[8]] << 16);
        o.w = slut[b[9]] | (slut[256 + b[10]] << 8) | (slut[512 + b[11]] << 16);
        reinterpret_cast<uint4 *>(out)[qd] = o;
    }
    // tail (total % 4 pixels)
    for (long long i = (quads << 2) + blockIdx.x * (long long)THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * THREADS)
        out[i] = slut[x[i * 3]] | (slut[256 + x[i * 3 + 1]] << 8) | (slut[512 + x[i * 3 + 2]] << 16);
}

__global__ void __launch_bounds__(THREADS)
requant_i32_kernel(const int32_t *__restrict__ x, int32_t *__restrict__ y, size_t count,
                   int shift, int is_signed) {
    for (size_t i = blockIdx.x * (size_t)THREADS + threadIdx.x; i < count;
         i += (size_t)gridDim.x * THREADS)
        y[i] = f8::requant(x[i], shift, is_signed);
}

unsigned grid_for(long long total) {
    long long b = (total + THREADS - 1) / THREADS;
    const long long cap = 148LL * 8 * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

f8::Epilogue make_ep(const f8_conv_args &a) {
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
    ep.int_pool = (a.flags & F8_OPF_INT_MAXPOOL) != 0;
    return ep;
}

}  // namespace

namespace f8host {

int launch_maxpool(const f8_conv_args &a, cudaStream_t s) {
    if (a.kh != 3 || a.kw != 3 || a.stride != 2 || a.pad != 1 || a.cin_pad != a.cout_pad ||
        a.cin_pad % 4 != 0) {
        set_error("maxpool: only 3x3 stride 2 pad 1 (fix_resnet.py:439)");
        return F8_ERR_UNSUPPORTED;
    }
    const long long total = (long long)a.n * a.hout * a.wout * (a.cin_pad >> 2);
    note_kernel("maxpool");
    maxpool_kernel<<<grid_for(total), THREADS, 0, s>>>(static_cast<const int32_t *>(a.in), a.n,
                                                       a.hin, a.win, a.hout, a.wout, a.cin_pad,
                                                       make_ep(a));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_pool_requant(const f8_conv_args &a, cudaStream_t s) {
    const long long total = (long long)a.n * (a.cin_pad >> 2);
    note_kernel("pool_requant");
    pool_requant_kernel<<<grid_for(total), THREADS, 0, s>>>(static_cast<const int32_t *>(a.in),
                                                            a.n, a.hin * a.win, a.cin_pad,
                                                            make_ep(a));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_pool_fc(const f8_conv_args &a, cudaStream_t s) {
    const DensePack pk = dense_pack_geometry(a.cin_pad, a.cout_pad, 1, 1);
    if (pk.mode != 0 || a.cin_pad % 16 != 0 || !a.out_f32 || a.out[0] != nullptr || a.carry_in != nullptr ||
        (size_t)4 * a.cin_pad > 96 * 1024) {
        set_error("pool_fc: unsupported geometry (cin_pad %d)", a.cin_pad);
        return F8_ERR_UNSUPPORTED;
    }
    const int imgs = 2;        // (4 per CTA halves the weight re-reads but measured slower, also with 1024 threads: 50 -> 86 us for ResNet50)
    const unsigned grid = (unsigned)((a.n + imgs - 1) / imgs);
    const size_t smem = (size_t)imgs * a.cin_pad;
    auto kern = imgs == 4 ? (a.out_signed[0] ? pool_fc_kernel<true, 4> : pool_fc_kernel<false, 4>)
                          : (a.out_signed[0] ? pool_fc_kernel<true, 2> : pool_fc_kernel<false, 2>);
    if (smem > 48 * 1024) F8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    note_kernel("pool_fc");
    F8_CUDA(launch_pdl(kern, grid, TAIL_THREADS, smem, s, static_cast<const int32_t *>(a.in), a.n, a.hin * a.win,
                       a.cin_pad, static_cast<const uint4 *>(a.wpack), pk.rows, a.bias, a.cout, a.out_shift[0],
                       a.out_f32, a.out_f32_ld));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_convert_input(const int32_t *x, void *out, int n, int h, int w, int is_signed, cudaStream_t s,
                         int *range_flag) {
    const long long total = (long long)n * h * w;
    note_kernel("convert_input");
    F8_CUDA(launch_pdl(convert_input_kernel, grid_for(total), THREADS, 0, s, x, static_cast<uint32_t *>(out), n, h * w,
                       is_signed ? -128 : 0, range_flag));
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_integerize_f32(const float *x, void *out, int n, int h, int w, int normalize, int fraclen,
                          cudaStream_t s, int is_signed, int *range_flag) {
    const long long total = (long long)n * h * w;
    const float scale = normalize ? ldexpf(1.0f, fraclen) : 255.0f;
    note_kernel("integerize_f32");
    integerize_f32_kernel<<<grid_for(total), THREADS, 0, s>>>(x, static_cast<uint32_t *>(out), n, h * w,
                                                              normalize, scale, is_signed ? -128 : 0, range_flag);
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_integerize_u8(const uint8_t *x, const uint8_t *lut_dev, void *out, int n, int h, int w,
                         cudaStream_t s) {
    const long long total = (long long)n * h * w;
    note_kernel("integerize_u8");
    integerize_u8_kernel<<<grid_for((total + 3) / 4), THREADS, 0, s>>>(x, lut_dev, static_cast<uint32_t *>(out),
                                                                        total);
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

int launch_requant_i32(const int32_t *x, int32_t *y, size_t count, int shift, int is_signed,
                       cudaStream_t s) {
    requant_i32_kernel<<<grid_for((long long)count), THREADS, 0, s>>>(x, y, count, shift,
                                                                      is_signed);
    F8_CUDA(cudaGetLastError());
    return F8_OK;
}

}  // namespace f8host
