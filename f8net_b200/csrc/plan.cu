// plan.cu -- the C ABI of libf8b200.so (include/f8b200.h): weight repacking, the execution
// plan (an ordered list of fused launches over caller-owned device buffers) and the
// per-kernel entry points used by the layer-level parity tests.
//
// What a plan replaces in the reference: the hand-over of Model.int_model()
// (/root/reference/fix_train.py:930-935) and every output = model(input) call
// (/root/reference/fix_train.py:693) on the int_op_only branch of IntModel.forward
// (/root/reference/models/fix_resnet.py:352-383, fix_mobilenet_v1.py:120-147,
// fix_mobilenet_v2.py:207-241).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "f8_common.cuh"
#include "host_pack.h"
#ifdef F8_WITH_UMMA
#include "tma_common.cuh"
#endif

namespace f8host {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local char g_kernel[96] = "";
void note_kernel(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_kernel, sizeof(g_kernel), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    (void)cudaGetLastError();
    return F8_ERR_CUDA;
}

#ifdef F8_WITH_UMMA
// cuTensorMapEncodeTiled through the runtime's driver entry point: no link dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static int encode_4d(CUtensorMap *out, CUtensorMapDataType dt, const void *base, const uint64_t dims[4],
                     const uint64_t strides[3], const uint32_t box[4], CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return F8_ERR_CUDA; }
    cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = fn(out, dt, 4, const_cast<void *>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu box %u %u %u %u", (int)r,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                  (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
        return F8_ERR_CUDA;
    }
    return F8_OK;
}
int encode_tmap_u8_4d(CUtensorMap *out, const void *base, const uint64_t dims[4], const uint64_t strides[3],
                      const uint32_t box[4], CUtensorMapSwizzle swizzle) {
    return encode_4d(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, base, dims, strides, box, swizzle);
}
int encode_tmap_u32_4d(CUtensorMap *out, const void *base, const uint64_t dims[4], const uint64_t strides[3],
                       const uint32_t box[4]) {
    return encode_4d(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}
#endif

}  // namespace f8host

using f8host::set_error;

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct PlanOp {
    f8_op op;
    size_t w_off = 0;   // byte offset of the packed weights inside the device blob
    size_t b_off = 0;   // byte offset of the padded bias
    size_t ws_off = (size_t)-1;   // dense 3x3: byte offset of the stage-major weight image (none: -1)
};

}  // namespace

struct f8_plan {
    int device = 0;
    int backend = 0;
    std::vector<PlanOp> ops;
    std::vector<f8_buffer> bufs;
    int64_t ws_per_image = 0;
    int image_h = 0, image_w = 0, num_classes = 0, head_signed = 0;
    int convert_op = -1;        // index of the F8_OP_CONVERT_INPUT op, if any
    uint8_t *blob = nullptr;    // device: packed weights + biases of every op
    size_t blob_bytes = 0;
    // forward_loss's input integerisation (f8_plan_set_input_prep)
    int prep_normalize = 0, prep_fraclen = 8;
    uint8_t *lut_dev = nullptr; // [3][256] table for F8_IN_NHWC3_U8
    // f8_plan_run_host, int32 NCHW input: pinned staging of the host-side 8-bit repack
    uint8_t *host_stage = nullptr;
    size_t host_stage_bytes = 0;
    cudaEvent_t host_stage_free = nullptr;   // recorded after the H2D copy that reads the staging
    cudaEvent_t dma_idle = nullptr;          // "the copy engine has executed everything enqueued so far"
    f8hp::Pool *pack_pool = nullptr;         // helper threads of the host-side narrowing (owned)
    int last_raw_images = 0;                 // images of the last run_host the copy engine shipped un-narrowed
    // input range check: a host-mapped word the narrowing / integerising kernels raise when a value lies
    // outside the head's 8 bits (no copy, no launch: a store that never executes for well-formed inputs),
    // and its device alias; host_out_of_range is the same finding of the host-side narrowing
    int *range_flag = nullptr, *range_flag_dev = nullptr;
    bool host_out_of_range = false;
    std::vector<std::string> kernel_names;   // per op: the kernel template of the last f8_plan_profile
};

// ---------------------------------------------------------------------------------------
// weight packing (host -> host)
// ---------------------------------------------------------------------------------------
extern "C" size_t f8_pack_weights_bytes(int kind, int cin, int cout, int cin_pad, int cout_pad,
                                        int kh, int kw) {
    (void)cin; (void)cout;
    // [12][cpad/4] u32 for the CUDA-core kernel + block-diagonal 64 x 64 images per channel
    // group for the tensor-core kernel (f8_common.cuh)
    if (kind == F8_OP_CONV_DW) return f8host::dw_pack_bytes(cin_pad);
    if (kind == F8_OP_CONV_DENSE) {
        const f8host::DensePack p = f8host::dense_pack_geometry(cin_pad, cout_pad, kh, kw);
        return (size_t)p.rows * (size_t)p.K_pad;
    }
    return 0;
}

extern "C" size_t f8_pack_weights_stage3x3_bytes(int cin_pad, int cout_pad) {
    if (cin_pad <= 0 || cin_pad % 64 || cout_pad <= 0) return 0;
    const int T = cout_pad > 64 ? 128 : 64;
    return (size_t)((cout_pad + T - 1) / T) * T * (size_t)cin_pad * 9;
}

extern "C" int f8_pack_weights_stage3x3(const void *dense_pack, int cin_pad, int cout_pad, void *dst_host) {
    if (!dense_pack || !dst_host || !f8_pack_weights_stage3x3_bytes(cin_pad, cout_pad)) {
        set_error("pack_weights_stage3x3: bad arguments (cin_pad must be a multiple of 64)");
        return F8_ERR_ARG;
    }
    const f8host::DensePack pk = f8host::dense_pack_geometry(cin_pad, cout_pad, 3, 3);
    const int T = cout_pad > 64 ? 128 : 64, NT = (cout_pad + T - 1) / T, NCG = cin_pad / 64;
    const uint8_t *src = static_cast<const uint8_t *>(dense_pack);
    uint8_t *dst = static_cast<uint8_t *>(dst_host);
    size_t chunk = 0;
    for (int t = 0; t < NT; ++t)
        for (int cg = 0; cg < NCG; ++cg)
            for (int tap = 0; tap < 9; ++tap)
                for (int j = 0; j < 4; ++j, ++chunk) {
                    const size_t kc = (size_t)(tap * cin_pad + cg * 64) / 16 + j;         // chunk of the (r, s, c) K order
                    for (int n = 0; n < T; ++n) {
                        const int row = t * T + n;
                        uint8_t *d = dst + (chunk * T + n) * 16;
                        if (row < pk.rows) std::memcpy(d, src + (kc * pk.rows + row) * 16, 16);
                        else std::memset(d, 0, 16);
                    }
                }
    return F8_OK;
}

extern "C" int f8_pack_weights(int kind, const int32_t *weight, int cin, int cout, int cin_pad,
                               int cout_pad, int kh, int kw, void *dst_host) {
    if (!weight || !dst_host || cin <= 0 || cout <= 0 || cin_pad < cin || cout_pad < cout) {
        set_error("pack_weights: bad arguments");
        return F8_ERR_ARG;
    }
    const size_t bytes = f8_pack_weights_bytes(kind, cin, cout, cin_pad, cout_pad, kh, kw);
    if (!bytes) {
        set_error("pack_weights: kind %d has no weights", kind);
        return F8_ERR_ARG;
    }
    std::memset(dst_host, 0, bytes);
    if (kind == F8_OP_CONV_DW) {
        // reference layout [C,1,3,3]; dp4a operand k of channel ch holds taps 4k..4k+3
        if (kh != 3 || kw != 3 || cin != cout || cin_pad != cout_pad || (cin_pad & 3)) {
            set_error("pack_weights: depthwise needs 3x3, cin == cout, cpad %% 4 == 0");
            return F8_ERR_UNSUPPORTED;
        }
        const int c4n = cin_pad >> 2;
        uint32_t *dst = static_cast<uint32_t *>(dst_host);
        for (int ch = 0; ch < cin; ++ch) {
            const int c4 = ch >> 2, c = ch & 3;
            for (int t = 0; t < 9; ++t) {
                const int32_t w = weight[(size_t)ch * 9 + t];
                if (w < -128 || w > 127) {
                    set_error("pack_weights: weight %d outside the 8-bit range", w);
                    return F8_ERR_UNSUPPORTED;
                }
                const int k = t >> 2, byte = t & 3;
                dst[(size_t)(c * 3 + k) * c4n + c4] |= ((uint32_t)w & 0xffu) << (8 * byte);
            }
        }
        // tensor-core form: group gi = ch / 64 is a dense 64 -> 64 3x3 conv whose weight matrix is
        // diagonal: K byte k = tap * 64 + c of output row o = c, image [36 chunks][64 rows][16 B]
        int8_t *dense = static_cast<int8_t *>(dst_host) + f8host::dw_dense_offset(cin_pad);
        for (int ch = 0; ch < cin; ++ch) {
            const int gi = ch >> 6, c = ch & 63;
            for (int t = 0; t < 9; ++t) {
                const size_t kc = (size_t)t * 4 + (c >> 4);
                dense[(size_t)gi * (36 * 64 * 16) + (kc * 64 + c) * 16 + (c & 15)] = (int8_t)weight[(size_t)ch * 9 + t];
            }
        }
        return F8_OK;
    }
    const f8host::DensePack p = f8host::dense_pack_geometry(cin_pad, cout_pad, kh, kw);
    int8_t *dst = static_cast<int8_t *>(dst_host);
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int r = 0; r < kh; ++r)
                for (int s = 0; s < kw; ++s) {
                    const int32_t w = weight[(((size_t)o * cin + c) * kh + r) * kw + s];
                    if (w < -128 || w > 127) {
                        set_error("pack_weights: weight %d outside the 8-bit range", w);
                        return F8_ERR_UNSUPPORTED;
                    }
                    const size_t k = p.mode == 1
                                         ? (size_t)r * p.row_bytes + (size_t)(s + p.shift_px) * 4 + c
                                         : ((size_t)r * kw + s) * cin_pad + c;
                    dst[((k >> 4) * (size_t)p.rows + o) * 16 + (k & 15)] = (int8_t)w;
                }
    return F8_OK;
}

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
static bool buf_ok(const f8_model_desc *d, int b) { return b >= -1 && b < d->n_buffers; }

static int check_shift(int s, const char *what) {
    if (s > 30 || s < -30) {   // what the reference's `1 << net_fl` on int32 can express
        set_error("%s shift %d is outside +-30", what, s);
        return F8_ERR_UNSUPPORTED;
    }
    return F8_OK;
}

extern "C" int f8_plan_create(const f8_model_desc *desc, int device, f8_plan **out) {
    if (!desc || !out) { set_error("plan_create: null argument"); return F8_ERR_ARG; }
    *out = nullptr;
    if (desc->abi_version != F8_ABI_VERSION) {
        set_error("plan_create: ABI version %d, library is %d", desc->abi_version, F8_ABI_VERSION);
        return F8_ERR_ARG;
    }
    if (desc->n_ops <= 0 || !desc->ops || desc->n_buffers < 0 || (desc->n_buffers && !desc->buffers)) {
        set_error("plan_create: empty descriptor");
        return F8_ERR_ARG;
    }
    f8_plan *p = new (std::nothrow) f8_plan();
    if (!p) { set_error("plan_create: out of host memory"); return F8_ERR_NOMEM; }
    p->device = device;
    p->ws_per_image = desc->workspace_per_image;
    p->image_h = desc->image_h; p->image_w = desc->image_w;
    p->num_classes = desc->num_classes; p->head_signed = desc->head_signed;
    p->bufs.assign(desc->buffers, desc->buffers + desc->n_buffers);
    for (const f8_buffer &b : p->bufs) {
        if (b.bytes_per_image < 0 || b.offset_per_image < 0 || (b.offset_per_image % 256) ||
            b.offset_per_image + b.bytes_per_image > desc->workspace_per_image) {
            set_error("plan_create: buffer table inconsistent with workspace_per_image");
            delete p;
            return F8_ERR_ARG;
        }
    }
    // pass 1: validate, size the blob
    size_t blob = 0;
    p->ops.resize(desc->n_ops);
    for (int i = 0; i < desc->n_ops; ++i) {
        PlanOp &po = p->ops[i];
        po.op = desc->ops[i];
        const f8_op &o = po.op;
        int rc = F8_OK;
        if (!buf_ok(desc, o.carry_in_buf) || !buf_ok(desc, o.carry_out_buf) ||
            !buf_ok(desc, o.out_buf[0]) || !buf_ok(desc, o.out_buf[1]) ||
            !(o.in_buf >= -2 && o.in_buf < desc->n_buffers)) {
            set_error("plan_create: op %d refers to a buffer outside the table", i);
            rc = F8_ERR_ARG;
        }
        if (!rc) rc = check_shift(o.carry_shift, "residual");
        if (!rc) rc = check_shift(o.out_shift[0], "requant");
        if (!rc) rc = check_shift(o.out_shift[1], "requant");
        if (!rc && (o.kind == F8_OP_CONV_DENSE || o.kind == F8_OP_CONV_DW || o.kind == F8_OP_HEAD_POOL ||
                    o.kind == F8_OP_POOL_FC)) {
            if (!o.weight || !o.bias) {
                set_error("plan_create: op %d has no weight / bias", i);
                rc = F8_ERR_ARG;
            } else {
                po.w_off = blob;
                const bool fc = o.kind == F8_OP_POOL_FC;       // the classifier: a 1x1 dense pack
                blob += align_up(f8_pack_weights_bytes(o.kind == F8_OP_CONV_DW ? o.kind : F8_OP_CONV_DENSE, o.cin,
                                                       o.cout, o.cin_pad, o.cout_pad, fc ? 1 : o.kh, fc ? 1 : o.kw),
                                 kAlign);
                po.b_off = blob;
                blob += align_up((size_t)o.cout_pad * sizeof(int32_t), kAlign);
                if (o.kind == F8_OP_CONV_DENSE && o.kh == 3 && o.kw == 3 && o.pad == 1 && o.cin_pad % 64 == 0) {
                    po.ws_off = blob;
                    blob += align_up(f8_pack_weights_stage3x3_bytes(o.cin_pad, o.cout_pad), kAlign);
                }
            }
        }
        if (!rc && o.kind == F8_OP_CONVERT_INPUT) {
            if (p->convert_op >= 0 || o.out_buf[0] < 0) {
                set_error("plan_create: at most one CONVERT_INPUT op, with out_buf[0] set");
                rc = F8_ERR_ARG;
            }
            p->convert_op = i;
        }
        if (rc) { delete p; return rc; }
    }
    // pass 2: pack on the host, one upload
    std::vector<uint8_t> host(blob ? blob : 1, 0);
    for (PlanOp &po : p->ops) {
        const f8_op &o = po.op;
        if (o.kind != F8_OP_CONV_DENSE && o.kind != F8_OP_CONV_DW && o.kind != F8_OP_HEAD_POOL &&
            o.kind != F8_OP_POOL_FC)
            continue;
        const bool fc = o.kind == F8_OP_POOL_FC;
        int rc = f8_pack_weights(o.kind == F8_OP_CONV_DW ? o.kind : F8_OP_CONV_DENSE, o.weight, o.cin, o.cout,
                                 o.cin_pad, o.cout_pad, fc ? 1 : o.kh, fc ? 1 : o.kw, host.data() + po.w_off);
        if (rc) { delete p; return rc; }
        if (po.ws_off != (size_t)-1) {
            rc = f8_pack_weights_stage3x3(host.data() + po.w_off, o.cin_pad, o.cout_pad, host.data() + po.ws_off);
            if (rc) { delete p; return rc; }
        }
        std::memcpy(host.data() + po.b_off, o.bias, (size_t)o.cout * sizeof(int32_t));
        po.op.weight = nullptr;   // host pointers are not kept: the caller owns them
        po.op.bias = nullptr;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&p->blob), host.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->blob, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (p->blob) cudaFree(p->blob);
        delete p;
        return f8host::cuda_fail(e, "plan_create upload");
    }
    p->blob_bytes = host.size();
    e = cudaHostAlloc(reinterpret_cast<void **>(&p->range_flag), 64, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *p->range_flag = 0;
        e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&p->range_flag_dev), p->range_flag, 0);
    }
    if (e != cudaSuccess) {
        f8_plan_destroy(p);
        return f8host::cuda_fail(e, "plan_create range flag");
    }
    *out = p;
    return F8_OK;
}

extern "C" void f8_plan_destroy(f8_plan *plan) {
    if (!plan) return;
    if (plan->blob) cudaFree(plan->blob);
    if (plan->lut_dev) cudaFree(plan->lut_dev);
    if (plan->host_stage) cudaFreeHost(plan->host_stage);
    if (plan->range_flag) cudaFreeHost(plan->range_flag);
    if (plan->host_stage_free) cudaEventDestroy(plan->host_stage_free);
    if (plan->dma_idle) cudaEventDestroy(plan->dma_idle);
    delete plan->pack_pool;
    delete plan;
}

// uint8 pixel p of channel c -> the 8-bit value forward_loss hands to the head, with torch's
// float32 operation order: ToTensor p / 255, Normalize (x - mean) / std (fix_train.py:299-318),
// then (255 x).round() or clamp(round(x * 2^fl), -127, 127) (fix_train.py:676-692).
extern "C" int f8_make_input_lut(int normalize, int fraclen, const float *mean3, const float *std3,
                                 uint8_t *lut) {
    if (!lut || fraclen < 0 || fraclen > 8) { set_error("make_input_lut: bad arguments"); return F8_ERR_ARG; }
    for (int c = 0; c < 3; ++c) {
        const float mean = (normalize && mean3) ? mean3[c] : 0.0f;
        const float sd = (normalize && std3) ? std3[c] : 1.0f;
        for (int p = 0; p < 256; ++p) {
            volatile float x = (float)p / 255.0f;            // volatile: one IEEE rounding per operation
            if (normalize) {
                x = x - mean;
                x = x / sd;
                volatile float r = x * ldexpf(1.0f, fraclen);
                r = nearbyintf(r);
                float q = r < -127.0f ? -127.0f : (r > 127.0f ? 127.0f : r);
                lut[c * 256 + p] = (uint8_t)(int8_t)(int)q;
            } else {
                volatile float r = 255.0f * x;
                r = nearbyintf(r);
                lut[c * 256 + p] = (uint8_t)((int)r & 0xff);
            }
        }
    }
    return F8_OK;
}

extern "C" int f8_plan_set_input_prep(f8_plan *plan, int normalize, int fraclen, const float *mean3,
                                      const float *std3) {
    if (!plan || fraclen < 0 || fraclen > 8) { set_error("set_input_prep: bad arguments"); return F8_ERR_ARG; }
    uint8_t lut[768];
    int rc = f8_make_input_lut(normalize, fraclen, mean3, std3, lut);
    if (rc) return rc;
    int cur = -1;
    F8_CUDA(cudaGetDevice(&cur));
    if (cur != plan->device) F8_CUDA(cudaSetDevice(plan->device));
    if (!plan->lut_dev) F8_CUDA(cudaMalloc(&plan->lut_dev, 768));
    F8_CUDA(cudaMemcpy(plan->lut_dev, lut, 768, cudaMemcpyHostToDevice));
    plan->prep_normalize = normalize ? 1 : 0;
    plan->prep_fraclen = fraclen;
    return F8_OK;
}

extern "C" int f8_integerize_f32(const float *x, void *out, int n, int h, int w, int normalize, int fraclen,
                                 void *stream) {
    if (!x || !out || n <= 0 || fraclen < 0 || fraclen > 8) { set_error("integerize_f32: bad arguments"); return F8_ERR_ARG; }
    return f8host::launch_integerize_f32(x, out, n, h, w, normalize, fraclen, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_integerize_u8(const uint8_t *x, const uint8_t *lut_dev, void *out, int n, int h, int w,
                                void *stream) {
    if (!x || !lut_dev || !out || n <= 0) { set_error("integerize_u8: bad arguments"); return F8_ERR_ARG; }
    return f8host::launch_integerize_u8(x, lut_dev, out, n, h, w, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_plan_workspace_bytes(const f8_plan *plan, int max_batch, size_t *bytes) {
    if (!plan || !bytes || max_batch <= 0) { set_error("workspace_bytes: bad arguments"); return F8_ERR_ARG; }
    *bytes = (size_t)plan->ws_per_image * (size_t)max_batch;
    return F8_OK;
}

extern "C" int f8_plan_set_backend(f8_plan *plan, int backend) {
    if (!plan || backend < 0 || backend > 2) { set_error("set_backend: bad arguments"); return F8_ERR_ARG; }
    if (backend >= 1 && !f8_has_umma(plan->device)) {
        set_error("set_backend: tcgen05 backend not available on device %d", plan->device);
        return F8_ERR_UNSUPPORTED;
    }
    plan->backend = backend;
    return F8_OK;
}

extern "C" int f8_plan_launch_count(const f8_plan *plan, int x_layout, int n, int chunk) {
    if (!plan || n <= 0) return 0;
    if (chunk <= 0 || chunk > n) chunk = n;
    const int passes = (n + chunk - 1) / chunk;
    int per = (int)plan->ops.size();
    if (x_layout == F8_IN_NHWC4_8 && plan->convert_op >= 0) --per;   // every other layout runs the convert op
    return per * passes;
}

static int run_op(const f8_plan *p, const PlanOp &po, int n, const uint8_t *x, int x_layout,
                  float *logits, uint8_t *ws, int cap, cudaStream_t s) {
    const bool x_is_nhwc4 = x_layout == F8_IN_NHWC4_8;
    const f8_op &o = po.op;
    auto buf = [&](int b) -> uint8_t * {
        if (b < 0) return nullptr;
        if (x_is_nhwc4 && p->convert_op >= 0 && b == p->ops[p->convert_op].op.out_buf[0])
            return const_cast<uint8_t *>(x);
        return ws + (size_t)p->bufs[b].offset_per_image * (size_t)cap;
    };
    if (o.kind == F8_OP_CONVERT_INPUT) {
        if (x_is_nhwc4) return F8_OK;
        if (x_layout == F8_IN_NCHW_F32)
            return f8host::launch_integerize_f32(reinterpret_cast<const float *>(x), buf(o.out_buf[0]), n, o.hin,
                                                 o.win, p->prep_normalize, p->prep_fraclen, s, p->head_signed,
                                                 p->range_flag_dev);
        if (x_layout == F8_IN_NHWC3_U8) {
            if (!p->lut_dev) { set_error("plan_run: call f8_plan_set_input_prep before a uint8 image input"); return F8_ERR_ARG; }
            return f8host::launch_integerize_u8(x, p->lut_dev, buf(o.out_buf[0]), n, o.hin, o.win, s);
        }
        return f8host::launch_convert_input(reinterpret_cast<const int32_t *>(x), buf(o.out_buf[0]),
                                            n, o.hin, o.win, p->head_signed, s, p->range_flag_dev);
    }
    f8_conv_args a{};
    a.n = n;
    a.cin = o.cin; a.cout = o.cout; a.cin_pad = o.cin_pad; a.cout_pad = o.cout_pad;
    a.kh = o.kh; a.kw = o.kw; a.stride = o.stride; a.pad = o.pad;
    a.hin = o.hin; a.win = o.win; a.hout = o.hout; a.wout = o.wout;
    a.in_signed = o.in_signed;
    a.in = o.in_buf == -2 ? x : buf(o.in_buf);
    a.wpack = p->blob + po.w_off;
    a.wpack_stage = po.ws_off != (size_t)-1 ? p->blob + po.ws_off : nullptr;
    a.bias = reinterpret_cast<const int32_t *>(p->blob + po.b_off);
    a.carry_in = reinterpret_cast<const int32_t *>(buf(o.carry_in_buf));
    a.carry_shift = o.carry_shift;
    a.relu = o.relu;
    a.flags = o.flags;
    a.carry_out = reinterpret_cast<int32_t *>(buf(o.carry_out_buf));
    for (int j = 0; j < 2; ++j) {
        a.out[j] = buf(o.out_buf[j]);
        a.out_shift[j] = o.out_shift[j];
        a.out_signed[j] = o.out_signed[j];
    }
    a.out_f32 = o.out_f32 ? logits : nullptr;
    a.out_f32_ld = p->num_classes;
    switch (o.kind) {
        case F8_OP_CONV_DENSE: return f8_conv_dense(&a, p->backend, s);
        case F8_OP_CONV_DW: return f8host::launch_dw3x3(a, s);
        case F8_OP_MAXPOOL: return f8host::launch_maxpool(a, s);
        case F8_OP_POOL_REQUANT: return f8host::launch_pool_requant(a, s);
        case F8_OP_POOL_FC: return f8host::launch_pool_fc(a, s);
        case F8_OP_HEAD_POOL: {
            if (p->backend != 1) {
                set_error("plan_run: the plan fuses head conv + max-pool (tcgen05 backend); rebuild it "
                          "for backend %d", p->backend);
                return F8_ERR_UNSUPPORTED;
            }
            int rc = f8host::launch_head_pool(a, s);
            if (rc == F8_ERR_UNSUPPORTED) set_error("plan_run: fused head geometry not supported");
            return rc;
        }
        default: set_error("plan_run: unknown op kind %d", o.kind); return F8_ERR_ARG;
    }
}

static int plan_run_impl(f8_plan *plan, const void *x_dev, int x_layout, int n,
                         float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                         int chunk, void *stream, std::vector<cudaEvent_t> *events);

static size_t input_image_bytes(const f8_plan *plan, int x_layout) {
    const size_t hw = (size_t)plan->image_h * plan->image_w;
    switch (x_layout) {
        case F8_IN_NHWC4_8: return hw * 4;
        case F8_IN_NHWC3_U8: return hw * 3;
        default: return hw * 3 * 4;          // int32 / float32 NCHW
    }
}

extern "C" int f8_plan_run(f8_plan *plan, const void *x_dev, int x_layout, int n,
                           float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                           int chunk, void *stream) {
    return plan_run_impl(plan, x_dev, x_layout, n, logits_dev, workspace_dev, workspace_bytes,
                         chunk, stream, nullptr);
}

// Per-launch device times of one f8_plan_run: CUDA events recorded on `stream` right before
// and after every launch.  op_ms[i] = time of op i summed over the passes (chunks).
extern "C" int f8_plan_profile(f8_plan *plan, const void *x_dev, int x_layout, int n,
                               float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                               int chunk, void *stream, float *op_ms, int n_ops) {
    if (!plan || !op_ms || n_ops != (int)plan->ops.size()) {
        set_error("plan_profile: op_ms must hold one float per plan op (%d)",
                  plan ? (int)plan->ops.size() : -1);
        return F8_ERR_ARG;
    }
    std::vector<cudaEvent_t> ev;
    plan->kernel_names.assign(plan->ops.size(), std::string());
    int rc = plan_run_impl(plan, x_dev, x_layout, n, logits_dev, workspace_dev, workspace_bytes,
                           chunk, stream, &ev);
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    for (int i = 0; i < n_ops; ++i) op_ms[i] = 0.f;
    if (!rc && e == cudaSuccess) {
        const size_t per = 2 * plan->ops.size();
        for (size_t base = 0; base + per <= ev.size(); base += per)
            for (int i = 0; i < n_ops; ++i) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, ev[base + 2 * i], ev[base + 2 * i + 1]) == cudaSuccess)
                    op_ms[i] += ms;
            }
    }
    for (cudaEvent_t x : ev) cudaEventDestroy(x);
    if (rc) return rc;
    if (e != cudaSuccess) return f8host::cuda_fail(e, "plan_profile sync");
    return F8_OK;
}

static int plan_run_impl(f8_plan *plan, const void *x_dev, int x_layout, int n,
                         float *logits_dev, void *workspace_dev, size_t workspace_bytes,
                         int chunk, void *stream, std::vector<cudaEvent_t> *events) {
    if (!plan || !x_dev || !logits_dev || !workspace_dev || n <= 0) {
        set_error("plan_run: bad arguments");
        return F8_ERR_ARG;
    }
    if (x_layout < F8_IN_NCHW_I32 || x_layout > F8_IN_NHWC3_U8) {
        set_error("plan_run: unknown input layout %d", x_layout);
        return F8_ERR_ARG;
    }
    if (x_layout != F8_IN_NHWC4_8 && plan->convert_op < 0) {
        set_error("plan_run: plan has no CONVERT_INPUT op for input layout %d", x_layout);
        return F8_ERR_ARG;
    }
    if (chunk <= 0 || chunk > n) chunk = n;
    if ((size_t)plan->ws_per_image * (size_t)chunk > workspace_bytes) {
        set_error("plan_run: workspace of %zu bytes is too small for %d images per pass (%zu needed)",
                  workspace_bytes, chunk, (size_t)plan->ws_per_image * (size_t)chunk);
        return F8_ERR_ARG;
    }
    int cur = -1;
    F8_CUDA(cudaGetDevice(&cur));
    if (cur != plan->device) F8_CUDA(cudaSetDevice(plan->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t x_img = input_image_bytes(plan, x_layout);
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int nn = (n - i0 < chunk) ? (n - i0) : chunk;
        const uint8_t *x = static_cast<const uint8_t *>(x_dev) + (size_t)i0 * x_img;
        float *lg = logits_dev + (size_t)i0 * plan->num_classes;
        for (const PlanOp &po : plan->ops) {
            if (events) {
                cudaEvent_t a, b;
                F8_CUDA(cudaEventCreate(&a));
                events->push_back(a);
                F8_CUDA(cudaEventCreate(&b));
                events->push_back(b);
                F8_CUDA(cudaEventRecord(a, s));
            }
            if (events) f8host::g_kernel[0] = 0;
            int rc = run_op(plan, po, nn, x, x_layout, lg, static_cast<uint8_t *>(workspace_dev),
                            chunk, s);
            if (rc) return rc;
            if (events) {
                F8_CUDA(cudaEventRecord(events->back(), s));
                plan->kernel_names[&po - plan->ops.data()] = f8host::g_kernel;
            }
        }
    }
    return F8_OK;
}

// ---------------------------------------------------------------------------------------
// Host side of the reference-facing call (host_pack.cpp: SIMD narrowing, per-plan helper threads)
// ---------------------------------------------------------------------------------------
extern "C" int f8_pack_input_host(const int32_t *x, int n, int h, int w, void *dst, int threads, int head_signed) {
    if (!x || !dst || n <= 0 || h <= 0 || w <= 0) { set_error("pack_input_host: bad arguments"); return F8_ERR_ARG; }
    static f8hp::Pool *pool = new f8hp::Pool;                // standalone entry point: one process-wide pool
    if (!pool->run(x, static_cast<uint8_t *>(dst), h, w, 0, (long long)n * h, threads, head_signed ? -128 : 0)) {
        set_error("pack_input_host: input values outside [%d, %d]", head_signed ? -128 : 0, head_signed ? 127 : 255);
        return F8_ERR_RANGE;
    }
    return F8_OK;
}

extern "C" const char *f8_host_pack_info(int *threads) {
    if (threads) *threads = f8hp::default_threads();
    return f8hp::isa_name();
}

static bool is_pinned_host(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// F8_IN_NCHW_I32 from host memory.  The tensor holds 8-bit-range integers in int32: 602 KB per image, of which
// the engine needs 200 KB.  Two resources can turn it into device-resident NHWC4 bytes, and they work AT THE
// SAME TIME, in sub-batches of 16 images:
//   * the host cores narrow a sub-batch into pinned staging (SIMD, streaming stores) and its 3.2 MB are copied;
//   * the copy engine ships a sub-batch as it is (9.6 MB, pinned source only) and a device kernel narrows it.
// The host takes sub-batches from the front, the copy engine from the back; a raw sub-batch is handed to the
// copy engine whenever an event shows its queue has drained, so the split follows the measured speeds of the
// two (many ranks sharing the host cores push it towards the copy engine, a lone rank towards the cores).
static int stage_int32_input(f8_plan *plan, const int32_t *x_host, int n, uint8_t *x_stage_dev, int nthreads, cudaStream_t s) {
    const int H = plan->image_h, W = plan->image_w;
    const size_t img4 = (size_t)H * W * 4, img_raw = (size_t)H * W * 3 * 4;
    const size_t bytes = (size_t)n * img4;
    if (plan->host_stage_bytes < bytes) {
        if (plan->host_stage) { F8_CUDA(cudaStreamSynchronize(s)); F8_CUDA(cudaFreeHost(plan->host_stage)); }
        plan->host_stage = nullptr;
        plan->host_stage_bytes = 0;
        F8_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&plan->host_stage), bytes, cudaHostAllocDefault));
        plan->host_stage_bytes = bytes;
    }
    if (!plan->host_stage_free) {
        F8_CUDA(cudaEventCreateWithFlags(&plan->host_stage_free, cudaEventDisableTiming));
        F8_CUDA(cudaEventCreateWithFlags(&plan->dma_idle, cudaEventDisableTiming));
    } else {
        F8_CUDA(cudaEventSynchronize(plan->host_stage_free));      // the previous call's copies have read the staging
    }
    if (!plan->pack_pool) plan->pack_pool = new f8hp::Pool;
    constexpr int SUB = 16;
    static const bool raw_allowed = [] { const char *e = getenv("F8_HOST_RAW_DMA"); return !(e && e[0] == '0'); }();
    // raw sub-batches land behind the NHWC4 region of the staging buffer (sized for n int32 images)
    uint8_t *raw_dev = x_stage_dev + bytes;
    const int raw_cap = (int)(((size_t)n * img_raw - bytes) / img_raw);
    const bool raw_ok = raw_allowed && n >= 4 * SUB && is_pinned_host(x_host);
    int lo = 0, hi = n, raw_used = 0;
    bool have_evt = false;
    auto ship_raw = [&]() -> int {
        const int cnt = SUB;
        hi -= cnt;
        uint8_t *dst = raw_dev + (size_t)raw_used * img_raw;
        F8_CUDA(cudaMemcpyAsync(dst, x_host + (size_t)hi * H * W * 3, (size_t)cnt * img_raw, cudaMemcpyHostToDevice, s));
        F8_CUDA(cudaEventRecord(plan->dma_idle, s));
        have_evt = true;
        raw_used += cnt;
        return f8host::launch_convert_input(reinterpret_cast<const int32_t *>(dst), x_stage_dev + (size_t)hi * img4, cnt, H, W,
                                            plan->head_signed, s, plan->range_flag_dev);
    };
    if (raw_ok && raw_cap >= SUB) { const int rc = ship_raw(); if (rc) return rc; }
    while (lo < hi) {
        const int cnt = hi - lo < SUB ? hi - lo : SUB;
        if (!plan->pack_pool->run(x_host, plan->host_stage, H, W, (long long)lo * H, (long long)(lo + cnt) * H, nthreads,
                                  plan->head_signed ? -128 : 0))
            plan->host_out_of_range = true;
        F8_CUDA(cudaMemcpyAsync(x_stage_dev + (size_t)lo * img4, plan->host_stage + (size_t)lo * img4, (size_t)cnt * img4,
                                cudaMemcpyHostToDevice, s));
        lo += cnt;
        // the copy engine has caught up with everything enqueued so far: give it a raw sub-batch from the back
        if (raw_ok && hi - lo >= 2 * SUB && raw_used + SUB <= raw_cap && have_evt && cudaEventQuery(plan->dma_idle) == cudaSuccess) {
            const int rc = ship_raw();
            if (rc) return rc;
        } else if (raw_ok) {
            (void)cudaGetLastError();                       // cudaErrorNotReady from the query is not an error
            F8_CUDA(cudaEventRecord(plan->dma_idle, s));
            have_evt = true;
        }
    }
    F8_CUDA(cudaEventRecord(plan->host_stage_free, s));
    plan->last_raw_images = raw_used;
    return F8_OK;
}

extern "C" int f8_plan_run_host(f8_plan *plan, const void *x_host, int x_layout, int n,
                                float *logits_host, void *x_stage_dev, float *logits_dev,
                                void *workspace_dev, size_t workspace_bytes, int chunk,
                                int sync, void *stream) {
    if (!plan || !x_host || !logits_host || !x_stage_dev || !logits_dev || n <= 0) {
        set_error("plan_run_host: bad arguments");
        return F8_ERR_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t x_img = input_image_bytes(plan, x_layout);
    int cur = -1;
    F8_CUDA(cudaGetDevice(&cur));
    if (cur != plan->device) F8_CUDA(cudaSetDevice(plan->device));
    const int nthreads = f8hp::default_threads();
    if (x_layout == F8_IN_NCHW_I32 && nthreads > 0) {
        const int rc0 = stage_int32_input(plan, static_cast<const int32_t *>(x_host), n, static_cast<uint8_t *>(x_stage_dev), nthreads, s);
        if (rc0) return rc0;
        x_layout = F8_IN_NHWC4_8;
    } else {
        F8_CUDA(cudaMemcpyAsync(x_stage_dev, x_host, x_img * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    int rc = f8_plan_run(plan, x_stage_dev, x_layout, n, logits_dev, workspace_dev,
                         workspace_bytes, chunk, stream);
    if (rc) return rc;
    F8_CUDA(cudaMemcpyAsync(logits_host, logits_dev,
                            (size_t)n * plan->num_classes * sizeof(float),
                            cudaMemcpyDeviceToHost, s));
    if (sync) {
        F8_CUDA(cudaStreamSynchronize(s));
        if (f8_plan_input_range(plan, 1)) {
            set_error("plan_run_host: input values outside the head's 8-bit range [%d, %d]: the reference's head conv "
                      "consumes the full int32 and would give different logits", plan->head_signed ? -128 : 0,
                      plan->head_signed ? 127 : 255);
            return F8_ERR_RANGE;
        }
    }
    return F8_OK;
}

extern "C" int f8_plan_input_range(f8_plan *plan, int clear) {
    if (!plan) return 0;
    const bool bad = plan->host_out_of_range || (plan->range_flag && *reinterpret_cast<volatile int *>(plan->range_flag));
    if (bad && clear) {
        plan->host_out_of_range = false;
        if (plan->range_flag) *reinterpret_cast<volatile int *>(plan->range_flag) = 0;
    }
    return bad ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// per-kernel entry points
// ---------------------------------------------------------------------------------------
static int check_args(const f8_conv_args *a, const char *who) {
    if (!a || !a->in || a->n <= 0) { set_error("%s: bad arguments", who); return F8_ERR_ARG; }
    int rc = check_shift(a->carry_shift, who);
    if (!rc) rc = check_shift(a->out_shift[0], who);
    if (!rc) rc = check_shift(a->out_shift[1], who);
    return rc;
}

extern "C" int f8_conv_dense(const f8_conv_args *a, int backend, void *stream) {
    int rc = check_args(a, "conv_dense");
    if (rc) return rc;
    if (!a->wpack || !a->bias) { set_error("conv_dense: no weights / bias"); return F8_ERR_ARG; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (backend >= 1) {
        // backend 1: resident-patch kernel for 3x3 s1, gather kernel for the rest;
        // backend 2: gather kernel only (A/B comparisons)
        if (backend == 1) {
            rc = f8host::launch_conv3x3_umma(*a, s);
            if (rc != F8_ERR_UNSUPPORTED) return rc;
        }
        rc = f8host::launch_conv_umma(*a, s);
        if (rc != F8_ERR_UNSUPPORTED) return rc;   // shapes outside the tcgen05 kernels: IMMA path
    }
    return f8host::launch_conv_mma(*a, s);
}

extern "C" int f8_conv_dw3x3(const f8_conv_args *a, void *stream) {
    int rc = check_args(a, "conv_dw3x3");
    if (rc) return rc;
    if (!a->wpack || !a->bias) { set_error("conv_dw3x3: no weights / bias"); return F8_ERR_ARG; }
    return f8host::launch_dw3x3(*a, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_maxpool3x3s2(const f8_conv_args *a, void *stream) {
    int rc = check_args(a, "maxpool3x3s2");
    if (rc) return rc;
    return f8host::launch_maxpool(*a, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_pool_fc(const f8_conv_args *a, void *stream) {
    int rc = check_args(a, "pool_fc");
    if (rc) return rc;
    if (!a->wpack || !a->bias || !a->out_f32) { set_error("pool_fc: no weights / bias / output"); return F8_ERR_ARG; }
    return f8host::launch_pool_fc(*a, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_head_pool(const f8_conv_args *a, void *stream) {
    int rc = check_args(a, "head_pool");
    if (rc) return rc;
    if (!a->wpack || !a->bias) { set_error("head_pool: no weights / bias"); return F8_ERR_ARG; }
    rc = f8host::launch_head_pool(*a, static_cast<cudaStream_t>(stream));
    if (rc == F8_ERR_UNSUPPORTED) set_error("head_pool: only the 224x224 7x7 s2 p3 3->64 head on sm_100");
    return rc;
}

extern "C" int f8_pool_requant(const f8_conv_args *a, void *stream) {
    int rc = check_args(a, "pool_requant");
    if (rc) return rc;
    return f8host::launch_pool_requant(*a, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_convert_input(const int32_t *x, void *out, int n, int h, int w, void *stream) {
    if (!x || !out || n <= 0 || h <= 0 || w <= 0) { set_error("convert_input: bad arguments"); return F8_ERR_ARG; }
    return f8host::launch_convert_input(x, out, n, h, w, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_requant_i32(const int32_t *x, int32_t *y, size_t count, int fl, int input_fl,
                              int is_signed, void *stream) {
    if (!x || !y) { set_error("requant_i32: null pointer"); return F8_ERR_ARG; }
    const int shift = input_fl - fl;
    int rc = check_shift(shift, "requant_i32");
    if (rc) return rc;
    if (!count) return F8_OK;
    return f8host::launch_requant_i32(x, y, count, shift, is_signed, static_cast<cudaStream_t>(stream));
}

extern "C" int f8_plan_read_buffer(const f8_plan *plan, int buf_index, int n, int chunk, const void *workspace_dev,
                                   void *dst_host, size_t dst_bytes, void *stream) {
    if (!plan || !workspace_dev || !dst_host || n <= 0 || chunk < n || buf_index < 0 ||
        buf_index >= (int)plan->bufs.size()) {
        set_error("plan_read_buffer: bad arguments");
        return F8_ERR_ARG;
    }
    const f8_buffer &b = plan->bufs[buf_index];
    const size_t bytes = (size_t)b.bytes_per_image * (size_t)n;
    if (bytes > dst_bytes) {
        set_error("plan_read_buffer: buffer %d holds %zu bytes for %d images, destination has %zu", buf_index, bytes, n,
                  dst_bytes);
        return F8_ERR_ARG;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint8_t *src = static_cast<const uint8_t *>(workspace_dev) + (size_t)b.offset_per_image * (size_t)chunk;
    F8_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, s));
    F8_CUDA(cudaStreamSynchronize(s));
    return F8_OK;
}

extern "C" int f8_plan_last_raw_images(const f8_plan *plan) { return plan ? plan->last_raw_images : 0; }

extern "C" int f8_plan_kernel_name(const f8_plan *plan, int op_index, char *dst, int cap) {
    if (!plan || !dst || cap <= 0 || op_index < 0 || op_index >= (int)plan->ops.size()) {
        set_error("plan_kernel_name: bad arguments");
        return F8_ERR_ARG;
    }
    const std::string name = op_index < (int)plan->kernel_names.size() ? plan->kernel_names[op_index] : std::string();
    snprintf(dst, (size_t)cap, "%s", name.c_str());
    return F8_OK;
}

extern "C" const char *f8_last_error(void) { return f8host::g_err; }
extern "C" int f8_abi_version(void) { return F8_ABI_VERSION; }

extern "C" int f8_has_umma(int device) {
#ifdef F8_WITH_UMMA
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return major == 10;
#else
    (void)device;
    return 0;
#endif
}

#ifndef F8_WITH_UMMA
namespace f8host {
int launch_conv_umma(const f8_conv_args &, cudaStream_t) { return F8_ERR_UNSUPPORTED; }
int launch_conv3x3_umma(const f8_conv_args &, cudaStream_t) { return F8_ERR_UNSUPPORTED; }
int launch_head_pool(const f8_conv_args &, cudaStream_t) { return F8_ERR_UNSUPPORTED; }
}  // namespace f8host
#endif
