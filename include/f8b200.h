/*
 * f8b200.h -- C ABI of libf8b200.so, the B200 (sm_100a) engine for F8Net's int_op_only
 * forward path.
 *
 * The reference (snap-research/F8Net) has no FFI: its int_op_only path is Python calling
 * torch ATen CPU int32 kernels.  The entry points below are what a binding for that path
 * binds instead; each one names the reference interface it replaces.  INTEGRATION.md shows
 * the reference-side stub (ctypes) a maintainer would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types; every call returns 0 (F8_OK)
 *     or a negative f8_status, never throws; f8_last_error() gives the message (thread local).
 *   - device memory for activations, logits and the workspace is OWNED BY THE CALLER
 *     (torch tensors on the host side); a plan owns only its repacked weights.
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*); no hidden
 *     synchronisation.  A plan is immutable after creation, so f8_plan_run is re-entrant
 *     across streams as long as each caller brings its own workspace.
 *   - activations between layers are NHWC, 8 bit (u8 or s8 as the consumer's
 *     input_symmetric says), channels padded to a multiple of 16 with zeros.
 *   - int32 tensors that stay inside the engine (residual carries, the inputs of the max-pool
 *     and the average pool) use a pixel-interleaved layout: with p the pixel index in
 *     image-major NHW order inside the launch and C the padded channel count, element (p, c) is
 *     at int32 index ((p >> 7) * (C / 4) + (c >> 2)) * 512 + (p & 127) * 4 + (c & 3)  -- blocks
 *     of 128 pixels x 4 channels, so that a warp owning 32 consecutive pixels moves contiguous
 *     512-byte runs.  A buffer holds ceil(pixels / 128) * 128 * C elements.
 */
#ifndef F8B200_H_
#define F8B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F8_ABI_VERSION 2

#if defined(__GNUC__)
#define F8_API __attribute__((visibility("default")))
#else
#define F8_API
#endif

typedef enum f8_status {
    F8_OK = 0,
    F8_ERR_ARG = -1,          /* malformed descriptor / null pointer / shape mismatch          */
    F8_ERR_CUDA = -2,         /* a CUDA runtime call failed; see f8_last_error()               */
    F8_ERR_UNSUPPORTED = -3,  /* valid for the reference but outside this engine (e.g. shift>30)*/
    F8_ERR_NOMEM = -4,
    F8_ERR_RANGE = -5         /* input values outside the head's 8 bits: the logits were computed from the low
                                 bytes and differ from the reference's (f8_plan_input_range)                  */
} f8_status;

/* Layout of the input handed to f8_plan_run (reference call surface: model(x) with x int32
 * NCHW, fix_resnet.py:352 / fix_mobilenet_v1.py:120 / fix_mobilenet_v2.py:207). */
typedef enum f8_input_layout {
    F8_IN_NCHW_I32 = 0,   /* the reference's own tensor: int32 [N,3,H,W], values in 8-bit range */
    F8_IN_NHWC4_8 = 1,    /* engine-native: 8-bit [N,H,W,4] (channel 3 = 0), u8 or s8 per head   */
    /* the two tensors BEFORE forward_loss's integerisation (fix_train.py:676-692): the engine
     * applies that step on the device, bit-exactly (see f8_plan_set_input_prep) */
    F8_IN_NCHW_F32 = 2,   /* the DataLoader's tensor: float32 [N,3,H,W] (ToTensor + Normalize)    */
    F8_IN_NHWC3_U8 = 3    /* decoded image bytes: uint8 [N,H,W,3]; ToTensor + Normalize +
                             integerisation folded into one 3 x 256 table                        */
} f8_input_layout;

typedef enum f8_op_kind {
    F8_OP_CONVERT_INPUT = 0,  /* NCHW int32 -> NHWC4 8 bit                                      */
    F8_OP_CONV_DENSE = 1,     /* groups == 1 conv or nn.Linear, tensor cores, fused epilogue    */
    F8_OP_CONV_DW = 2,        /* depthwise 3x3, CUDA-core int MAC, fused epilogue               */
    F8_OP_MAXPOOL = 3,        /* ResNet head 3x3 s2 p1 max-pool with float32 round trip         */
    F8_OP_POOL_REQUANT = 4,   /* FXQAvgPool2d sum over HxW + requant for the classifier         */
    F8_OP_HEAD_POOL = 5,      /* ResNet head fused: 7x7 s2 conv + ReLU + float round trip +
                                 3x3 s2 max-pool (tcgen05 backend only); hout/wout = pooled size  */
    F8_OP_POOL_FC = 6         /* network tail fused: FXQAvgPool2d sum over hin x win + requant
                                 (out_shift[0] / out_signed[0], no buffer) + nn.Linear + .float()
                                 (fix_quant_ops.py:126-134, fix_resnet.py:367-383)                */
} f8_op_kind;

/* f8_op.flags / f8_conv_args.flags */
typedef enum f8_op_flags {
    /* MAXPOOL / HEAD_POOL: the head pool is FXQMaxPool2d (FLAGS.quant_maxpool, fix_quant_ops.py:141-157,
     * fix_resnet.py:331-334, :355-356): a pure integer max, WITHOUT the x.float() ... .int() round trip
     * of nn.MaxPool2d (fix_resnet.py:358-359) */
    F8_OPF_INT_MAXPOOL = 1
} f8_op_flags;

/*
 * One fused launch.  "Epilogue" = everything the reference does between one int layer's
 * accumulator and the 8-bit input of the next int layer(s), in the reference's order:
 *
 *   v  = acc + bias                                   nn.Conv2d/nn.Linear bias (fix_quant_ops.py:706)
 *   if carry_in:  d = carry_shift                     IntBlock residual, fix_resnet.py:40-76,
 *       d >= 0 ? carry <<= d : v <<= -d  (wrap)       fix_mobilenet_v2.py:34-48
 *       v = max(v + carry (wrap), INT32_MIN+1)
 *   if relu:      v = max(v, 0)                        nn.ReLU in body / post_relu / head / tail
 *   if carry_out: carry_out = v                        int32 tensor a later residual add reads
 *   for j in 0,1 if out[j]:                            int_op_only_fix_quant of the consumer,
 *       out[j] = requant(v, out_shift[j], out_signed[j])   fix_quant_ops.py:90-114
 *   if out_f32:   out_f32 = (float) v                  classifier(x).float(), fix_resnet.py:383
 *
 * requant(v, n, s): n > 0: round-half-even(v / 2^n) by the reference's exact formula;
 * n <= 0: wrapping v << -n; then clamp to [-127,127] (s) or [0,255].
 * Buffers are referred to by index into the plan's buffer table; -1 = unused.
 */
typedef struct f8_op {
    int32_t kind;              /* f8_op_kind */
    /* geometry (per image) */
    int32_t cin, cout;         /* logical channels                                   */
    int32_t cin_pad, cout_pad; /* channel counts of the NHWC buffers                 */
    int32_t kh, kw, stride, pad;
    int32_t hin, win, hout, wout;
    int32_t in_signed;         /* activation operand: 1 = s8 (input_symmetric), 0 = u8 */
    /* operands */
    int32_t in_buf;            /* 8-bit NHWC input (int32 NHWC for MAXPOOL / POOL_REQUANT);
                                  -2 = the plan input x (CONVERT_INPUT, or head conv when
                                  the caller passes F8_IN_NHWC4_8)                     */
    const int32_t *weight;     /* HOST, reference layout [O, C/g, kh, kw] or [O, K]  */
    const int32_t *bias;       /* HOST, [O]                                          */
    /* epilogue */
    int32_t carry_in_buf;
    int32_t carry_shift;
    int32_t relu;
    int32_t carry_out_buf;
    int32_t out_buf[2];
    int32_t out_shift[2];
    int32_t out_signed[2];
    int32_t out_f32;           /* 1: write float logits to the plan output          */
    int32_t flags;             /* f8_op_flags                                        */
} f8_op;

/* One workspace buffer: bytes per image and its offset (in per-image bytes) inside the
 * workspace; the address at run time is  workspace + offset_per_image * N.  Offsets are
 * chosen by the caller (the host planner reuses dead buffers); multiples of 256. */
typedef struct f8_buffer {
    int64_t bytes_per_image;
    int64_t offset_per_image;
} f8_buffer;

typedef struct f8_model_desc {
    int32_t abi_version;       /* F8_ABI_VERSION */
    int32_t n_ops;
    const f8_op *ops;
    int32_t n_buffers;
    const f8_buffer *buffers;
    int64_t workspace_per_image;  /* max over buffers of offset + bytes (multiple of 256) */
    int32_t image_h, image_w;     /* 224 x 224 */
    int32_t num_classes;          /* 1000 */
    int32_t head_signed;          /* head conv input_symmetric */
} f8_model_desc;

typedef struct f8_plan f8_plan;

/* Replaces: Model.int_model() + .cpu() hand-over (fix_train.py:930-935) -- the engine takes
 * the int32 tensors of IntModel.state_dict() as they are and repacks them to int8 tiles. */
F8_API int f8_plan_create(const f8_model_desc *desc, int device, f8_plan **out);
F8_API void f8_plan_destroy(f8_plan *plan);

/* Bytes of caller-owned device workspace needed to run up to max_batch images at once. */
F8_API int f8_plan_workspace_bytes(const f8_plan *plan, int max_batch, size_t *bytes);

/* Replaces: output = model(input) (fix_train.py:693) i.e. IntModel.forward, int_op_only
 * branch.  x_dev: device input in x_layout; logits_dev: float32 [n, num_classes] holding the
 * exact int32 logits (fix_resnet.py:383).  chunk: images per pass through the layer list
 * (<= the batch the workspace was sized for; 0 = n). */
F8_API int f8_plan_run(f8_plan *plan, const void *x_dev, int x_layout, int n, float *logits_dev,
                void *workspace_dev, size_t workspace_bytes, int chunk, void *stream);

/* Replaces: the input preparation of forward_loss (fix_train.py:676-692) and the transform in
 * front of it (ToTensor + Normalize, fix_train.py:299-318) for the F8_IN_NCHW_F32 / F8_IN_NHWC3_U8
 * layouts.  normalize = FLAGS.normalize:
 *   0: x_int = (255 * x).round().int()                    (unsigned head, fraclen 8)
 *   1: x_int = clamp(round(x * 2^fraclen), -127, 127)      (signed head, fraclen = head.input_fraclen)
 * with float32 arithmetic and round-half-even exactly as torch computes them; for uint8 pixels p
 * the float tensor is x = (p / 255 - mean[c]) / std[c] (mean 0 / std 1 when normalize = 0).
 * Defaults without this call: normalize = the head's signedness, fraclen = 8 (unsigned) or the
 * head's input_fraclen given here.  The low byte of x_int is kept, like F8_IN_NCHW_I32, and a value
 * outside the head's 8 bits is recorded (f8_plan_input_range). */
F8_API int f8_plan_set_input_prep(f8_plan *plan, int normalize, int fraclen, const float *mean3,
                           const float *std3);

/* Same, with HOST buffers (pinned or pageable): copies x host->device, runs, copies the
 * logits back, all on `stream`, then (sync != 0) synchronises the stream.  x_stage_dev must hold the
 * input (n * 3*H*W*4 bytes for NCHW_I32 / NCHW_F32, n * H*W*4 for NHWC4_8, n * H*W*3 for NHWC3_U8).  sync = 0 leaves the stream
 * unsynchronised (x_host / logits_host must then be pinned and stay alive) so that a caller
 * can overlap the copies of one batch with the compute of another on a second stream.  This
 * is the entry the reference-facing call surface uses for CPU tensors.
 * F8_IN_NCHW_I32: the tensor holds 8-bit-range integers (fix_train.py:682-692), so the call first
 * narrows it on the host cores to NHWC4 bytes (the low byte of every value, as the device-side
 * conversion keeps) in a plan-owned pinned staging buffer and copies a third of the bytes; x_host
 * is read by the host cores AND, when it is pinned, directly by the copy engine: the batch is split in
 * sub-batches of 16 images which the host narrows from the front while the copy engine ships raw
 * sub-batches from the back (narrowed by a device kernel), each at its own pace.  With sync = 0 the
 * tensor must therefore stay unmodified until `stream` has executed the call's work.  Environment:
 * F8_HOST_PACK_THREADS = host threads (default min(16, usable cores / LOCAL_WORLD_SIZE); 0 = copy the
 * int32 tensor as is), F8_HOST_RAW_DMA=0 = host narrowing only.  One call at a time per plan. */
F8_API int f8_plan_run_host(f8_plan *plan, const void *x_host, int x_layout, int n, float *logits_host,
                     void *x_stage_dev, float *logits_dev, void *workspace_dev,
                     size_t workspace_bytes, int chunk, int sync, void *stream);

/* Input range check, always on.  The reference hands IntModel.forward 8-bit-range integers in int32
 * (fix_train.py:682-692) but its head conv consumes the full int32 (fix_resnet.py:355), and forward_loss
 * asserts input >= 0 on the float tensor (fix_train.py:689); the engine keeps the low byte.  Every narrowing
 * step of a plan (host SIMD narrowing, the device kernels behind F8_IN_NCHW_I32 and F8_IN_NCHW_F32) therefore
 * records whether it saw a value outside [0, 255] (unsigned head) or [-128, 127] (signed head; NaN counts as
 * outside), at no cost for well-formed inputs: the device kernels raise a host-mapped word.
 *   f8_plan_run_host(sync != 0) returns F8_ERR_RANGE for such a call (the logits are written, computed from
 *     the low bytes) and clears the record;
 *   asynchronous calls (f8_plan_run, f8_plan_run_host(sync = 0)) cannot know yet: the record stays raised
 *     until this function reads it, after the caller has synchronised the stream.
 * Returns 1 when an input since the last clear was out of range, else 0; clear != 0 resets the record. */
F8_API int f8_plan_input_range(f8_plan *plan, int clear);

/* How the host side of f8_plan_run_host narrows int32 inputs on this machine: returns the SIMD body
 * ("avx512" | "avx2" | "sse2" | "scalar"), *threads = helper threads per plan. */
F8_API const char *f8_host_pack_info(int *threads);
/* Images of the plan's most recent f8_plan_run_host(F8_IN_NCHW_I32) that the copy engine shipped
 * un-narrowed (the rest were narrowed by the host cores). */
F8_API int f8_plan_last_raw_images(const f8_plan *plan);

/* Measurement aid (the reference's only timing device is a wall-clock decorator,
 * fix_train.py:41-53): one f8_plan_run with a CUDA event pair around every launch on
 * `stream`; synchronises, then op_ms[i] = device time of op i summed over the passes. */
F8_API int f8_plan_profile(f8_plan *plan, const void *x_dev, int x_layout, int n, float *logits_dev,
                    void *workspace_dev, size_t workspace_bytes, int chunk, void *stream,
                    float *op_ms, int n_ops);

/* Name of the kernel template that served op `op_index` in the most recent f8_plan_profile of this
 * plan ("" when that op launched nothing, e.g. CONVERT_INPUT with an engine-native input): lets a
 * measurement be attributed per kernel template rather than per op kind. */
F8_API int f8_plan_kernel_name(const f8_plan *plan, int op_index, char *dst, int cap);

/* Debug / parity aid (the reference's equivalent is a forward hook on a module, e.g.
 * register_forward_pre_hook on an int nn.Conv2d to see its 8-bit input): copies plan buffer
 * `buf_index` as the last pass of f8_plan_run(n <= chunk images, chunk) left it in `workspace_dev`
 * -- n * bytes_per_image bytes from workspace + offset_per_image * chunk -- to `dst_host`, on
 * `stream`, and synchronises.  Meaningful for every buffer only when the planner gave each buffer
 * its own range (host planner: keep_buffers=True); with liveness-based reuse a buffer holds what
 * its last writer left. */
F8_API int f8_plan_read_buffer(const f8_plan *plan, int buf_index, int n, int chunk,
                               const void *workspace_dev, void *dst_host, size_t dst_bytes, void *stream);

/* Number of kernel launches one f8_plan_run(x_layout, n, chunk) enqueues. */
F8_API int f8_plan_launch_count(const f8_plan *plan, int x_layout, int n, int chunk);
/* Which dense-conv backend the plan uses: 0 = mma.sync (legacy IMMA); 1 = tcgen05 (resident-
 * patch kernel for 3x3 stride 1, gather kernel otherwise); 2 = tcgen05 gather kernel only */
F8_API int f8_plan_set_backend(f8_plan *plan, int backend);

/* ------------------------------------------------------------------------------------
 * Per-kernel entry points (layer-level parity tests; same kernels the plan launches).
 * All pointers are DEVICE pointers except where noted.
 * ---------------------------------------------------------------------------------- */
typedef struct f8_conv_args {
    int32_t n;
    int32_t cin, cout, cin_pad, cout_pad;
    int32_t kh, kw, stride, pad;
    int32_t hin, win, hout, wout;
    int32_t in_signed;
    const void *in;            /* 8-bit NHWC [n,hin,win,cin_pad] (int32 for maxpool / pool) */
    const void *wpack;         /* from f8_pack_weights                                    */
    const int32_t *bias;       /* [cout_pad], zero padded                                 */
    const int32_t *carry_in;   /* int32, carry layout (see top), n*hout*wout pixels, or NULL */
    int32_t carry_shift;
    int32_t relu;
    int32_t *carry_out;        /* same layout, or NULL                                    */
    void *out[2];              /* 8-bit NHWC [n,hout,wout,cout_pad] or NULL               */
    int32_t out_shift[2];
    int32_t out_signed[2];
    float *out_f32;            /* [n*hout*wout, out_f32_ld] or NULL                       */
    int32_t out_f32_ld;
    int32_t flags;             /* f8_op_flags                                              */
    const void *wpack_stage;   /* optional, dense 3x3 with cin_pad % 64 == 0: the same weights in the
                                  stage-major order of f8_pack_weights_stage3x3 (one contiguous bulk copy
                                  per pipeline stage instead of twelve strided ones); NULL = not used   */
} f8_conv_args;

/* Bytes of the packed weight image for a layer and the packing itself (host -> host).
 * kind: F8_OP_CONV_DENSE or F8_OP_CONV_DW.  weight: reference layout int32. */
F8_API size_t f8_pack_weights_bytes(int kind, int cin, int cout, int cin_pad, int cout_pad, int kh,
                             int kw);
F8_API int f8_pack_weights(int kind, const int32_t *weight, int cin, int cout, int cin_pad,
                    int cout_pad, int kh, int kw, void *dst_host);

/* Stage-major copy of a dense 3x3 pack for the resident-patch tcgen05 kernel: tiles of T = 128 (cout_pad > 64) or
 * 64 output rows; per tile, per 64-channel input group, per filter row, per filter column: four 16-byte K chunks of
 * T rows -- exactly the order and granularity in which the kernel's weight ring consumes them.  dense_pack is the
 * image f8_pack_weights produced for the same layer (host pointers). */
F8_API size_t f8_pack_weights_stage3x3_bytes(int cin_pad, int cout_pad);
F8_API int f8_pack_weights_stage3x3(const void *dense_pack, int cin_pad, int cout_pad, void *dst_host);

/* Replaces: int nn.Conv2d.__call__ (groups == 1) / nn.Linear.__call__ + the consumer-side
 * int_op_only_fix_quant, ReLU and residual add around it (fix_resnet.py:28-77). */
F8_API int f8_conv_dense(const f8_conv_args *a, int backend, void *stream);
/* Replaces: int nn.Conv2d.__call__ with groups == in_channels (fix_mobilenet_v1.py:33,
 * fix_mobilenet_v2.py:28) + consumer-side requant / ReLU. */
F8_API int f8_conv_dw3x3(const f8_conv_args *a, void *stream);
/* Replaces: self.head[-1](x.float()).int()  (fix_resnet.py:358-359). in = int32, carry layout. */
F8_API int f8_maxpool3x3s2(const f8_conv_args *a, void *stream);
/* Replaces: x = self.head[:-1](x); x = self.head[-1](x.float()).int() in one launch
 * (fix_resnet.py:355-362): 7x7 s2 p3 conv of the NHWC4 image + bias + ReLU + float32 round trip
 * + 3x3 s2 p1 max-pool + consumer requant(s) / int32 carry.  a->hout/wout = 56 (pooled);
 * sm_100 only; F8_ERR_UNSUPPORTED for any other geometry. */
F8_API int f8_head_pool(const f8_conv_args *a, void *stream);
/* Replaces: FXQAvgPool2d.forward int branch + int_op_only_fix_quant for the classifier
 * (fix_quant_ops.py:126-134, fix_resnet.py:367-374). in = int32 carry layout, n*h*w pixels,
 * out[0] = 8-bit [n,c_pad]; carry_out (tests) = plain int32 [n,c_pad]. */
F8_API int f8_pool_requant(const f8_conv_args *a, void *stream);
/* Replaces: the same + the classifier nn.Linear + .float() (fix_resnet.py:367-383) in one launch.
 * in = int32 carry layout [n*h*w pixels, cin_pad]; wpack = dense pack of the [cout, cin] weight;
 * out_shift[0] / out_signed[0] = requant of the pooled sum; out_f32 = float32 [n, out_f32_ld]. */
F8_API int f8_pool_fc(const f8_conv_args *a, void *stream);
/* Replaces: the int32 NCHW tensor hand-over at model(x): repack to NHWC4 8 bit.
 * x int32 [n,3,h,w] -> out 8-bit [n,h,w,4]. */
F8_API int f8_convert_input(const int32_t *x, void *out, int n, int h, int w, void *stream);
/* Replaces: forward_loss's integerisation of the float tensor (fix_train.py:676-692), see
 * f8_plan_set_input_prep.  x float32 [n,3,h,w] -> out 8-bit [n,h,w,4]. */
F8_API int f8_integerize_f32(const float *x, void *out, int n, int h, int w, int normalize, int fraclen,
                      void *stream);
/* Replaces: ToTensor + Normalize + the same integerisation for decoded uint8 pixels.
 * x uint8 [n,h,w,3], lut_dev: device table [3][256] of 8-bit values -> out 8-bit [n,h,w,4].
 * f8_make_input_lut fills the (host) table for given normalize / fraclen / mean / std. */
F8_API int f8_integerize_u8(const uint8_t *x, const uint8_t *lut_dev, void *out, int n, int h, int w,
                     void *stream);
/* Host-only (no CUDA call): the narrowing f8_plan_run_host applies to an F8_IN_NCHW_I32 host tensor.
 * x int32 [n,3,h,w] -> dst bytes [n,h,w,4] = the low byte of each channel value, channel 3 = 0 (what
 * f8_convert_input produces on the device).  threads <= 1: the calling thread only.  Returns F8_ERR_RANGE
 * (dst is still written) when a value lies outside [0, 255] (head_signed = 0) or [-128, 127] (head_signed
 * != 0): the same pass that narrows also ORs (value - low bound) of everything it reads. */
F8_API int f8_pack_input_host(const int32_t *x, int n, int h, int w, void *dst, int threads, int head_signed);
F8_API int f8_make_input_lut(int normalize, int fraclen, const float *mean3, const float *std3,
                      uint8_t *lut_host768);
/* Replaces: int_op_only_fix_quant as a standalone op (fix_quant_ops.py:90-114):
 * y[i] = requant(x[i], input_fl - fl, is_signed), int32 in / int32 out. */
F8_API int f8_requant_i32(const int32_t *x, int32_t *y, size_t count, int fl, int input_fl,
                   int is_signed, void *stream);

F8_API const char *f8_last_error(void);
F8_API int f8_abi_version(void);
/* 1 when the library was built with the tcgen05 path and the device is sm_100 */
F8_API int f8_has_umma(int device);

#ifdef __cplusplus
}
#endif
#endif /* F8B200_H_ */
