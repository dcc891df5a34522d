/* keep_b200 — C-ABI of the Blackwell-native KEEP inference path (libkeep_b200.so).
 *
 * One call replaces the reference's `keep_net(clip, need_upscale=False)`:
 *   caller      modules/keep_processor.py:174,177,268,270   (self.keep_net(...))
 *   callee      modules/deps/wm_basicsr/archs/keep_arch.py:1008-1145 (KEEP.forward)
 *   lifecycle   modules/keep_model_loader.py:93-97 (construct), :120-121 (load_state_dict, eval),
 *               :28-48 (KEEPModelPack.load_device / offload -> .to(device))
 *
 * Plain pointers and sizes only; no torch types.  The Python shim (comfyui-keep_b200/keep_net.py)
 * binds these with ctypes and passes `tensor.data_ptr()` and the current CUDA stream handle.
 * Every function returns 0 on success or a negative error code and never aborts the process;
 * `keep_last_error()` returns the thread-local message (the reference's nodes catch exceptions
 * and return (None,), nodes.py:83-88,131-136).  There is no CPU fallback: without a CUDA device
 * `keep_create` fails.
 */
#ifndef KEEP_B200_H
#define KEEP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct keep_engine_s* keep_handle;

/* one entry of the reference state dict (after the loader's key renames, keep_model_loader.py:110-118):
 * `data` is HOST memory, fp32, contiguous, in the reference's own layout (conv OIHW, linear (out,in)). */
typedef struct {
    const char* name;
    const float* data;
    int32_t ndim;
    int64_t shape[4];
} keep_weight_desc;

/* KEEP_OUT_U8_BGR: (b, T, 512, 512, 3) uint8, BGR, HWC -- the image tensor2img(rgb2bgr=True, min_max=(-1, 1)) would make of
 * the fp32 output (clamp, (x+1)/2*255, round-half-even; B/utils/img_util.py:38-94) */
enum { KEEP_OUT_F32 = 0, KEEP_OUT_F16 = 1, KEEP_OUT_U8_BGR = 2 };
enum {
    KEEP_FLAG_DEFAULT = 0,
    KEEP_FLAG_FP16_FEATURES = 1, /* RESERVED: fp16 feature-map storage; keep_create rejects it (feature maps are fp32) */
    KEEP_FLAG_TCGEN05 = 2,       /* run eligible convolutions / GEMMs on the tcgen05 tensor-core kernel */
    KEEP_FLAG_TC_SPLIT3 = 4,     /* with TCGEN05: split-precision operands (A=Ah+Al, W=Wh+Wl, 3 MMAs) -> fp32-grade results */
    KEEP_FLAG_CUDA_GRAPH = 8,    /* capture one clip forward per T into a CUDA graph after the first (eager) call and replay it */
    KEEP_FLAG_TC_WIDE = 16,      /* with TC_SPLIT3: generator / CFT layers that read a raw (un-normalised) feature map take their
                                    activation operand as a bf16 pair (fp32 exponent range, 16 mantissa bits) instead of an fp16
                                    pair (22 bits, overflows beyond 65504) */
    KEEP_FLAG_BATCH_CLIPS = 32,  /* keep_forward with b > 1: groups of clips (2 by default, KEEP_BATCH_MAX) advance through the
                                    per-frame recurrence in lockstep -- one batched hq_encoder / code-transformer / generator pass
                                    per frame index instead of one per clip; engine-owned workspace only */
    KEEP_FLAG_PLAN_ONLY = 256    /* host-side planning only (strict key check + workspace sizing); keep_forward fails */
};

/* Build an engine on CUDA device `device` from the 896-tensor KEEP state dict: packs the weights
 * into kernel layouts and uploads them.  Replaces ARCH_REGISTRY.get('KEEP')(**cfg) +
 * load_state_dict(strict=True) + .to(device): missing or mis-shaped keys are an error. */
int keep_create(keep_handle* out, int device, const keep_weight_desc* weights, int n_weights, int flags);

/* Bytes of device workspace `keep_forward` needs for a (b, T) call (activations only; weights are
 * owned by the engine). */
size_t keep_workspace_bytes(keep_handle h, int b, int T);

/* The hot path.  x_dev: (b, T, 3, 512, 512) fp32 NCHW in [-1, 1] on the engine's device, not
 * mutated.  out_dev: (b, T, 3, 512, 512), fp32 or fp16 per `out_dtype`, unclamped — exactly what
 * KEEP.forward returns in eval mode (keep_arch.py:1138,1145).  T >= 2 (the reference duplicates a
 * single frame, keep_processor.py:173-175,266-268).  workspace may be NULL (the engine then owns
 * and grows its own); `stream` is a cudaStream_t (NULL = default stream).  Asynchronous: no host
 * synchronisation inside. */
int keep_forward(keep_handle h, const float* x_dev, int b, int T, void* out_dev, int out_dtype, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Clips per lockstep group for engines created with KEEP_FLAG_BATCH_CLIPS (default 2, or KEEP_BATCH_MAX; 1 = clip by clip).
 * Replaces nothing in the reference: its caller loops over clips one at a time (keep_processor.py:263-270). */
int keep_set_batch_clips(keep_handle h, int max_clips);

/* Same path with the caller-side conversions folded in: x_u8_dev is (b, T, 512, 512, 3) uint8 BGR HWC --
 * the aligned crops as OpenCV holds them -- converted on the device exactly as keep_processor.py:258-260 does on the host
 * (img2tensor(crop / 255., bgr2rgb=True) then normalize(0.5, 0.5)); out_u8_dev is KEEP_OUT_U8_BGR.  A 20-frame clip moves
 * 15.7 MB each way instead of 62.9 MB. */
int keep_forward_u8(keep_handle h, const unsigned char* x_u8_dev, int b, int T, unsigned char* out_u8_dev, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Free packed weights and any engine-owned workspace (KEEPModelPack.offload, keep_model_loader.py:45-61). */
int keep_destroy(keep_handle h);

const char* keep_last_error(void);

/* Sticky non-finite status of the engine since the last clearing read.  keep_forward never synchronises, so it cannot look
 * at its own result; instead the kernels at the joints of the path (flow warp, Kalman update, logits -> argmax, final frame
 * conversion) OR a bit into a device word when they meet inf / NaN -- e.g. an fp16-pair operand overflow on raw features
 * beyond 65504 (see KEEP_FLAG_TC_WIDE), which would otherwise reach the caller as a silently wrong code index or a NaN frame.
 * This call synchronises the device, so make it after your own sync (the caller's .cpu() in tensor2img, keep_processor.py:273).
 * Replaces nothing in the reference (PyTorch fp32 does not overflow here). */
enum { KEEP_STATUS_BAD_LATENT = 1, KEEP_STATUS_BAD_LOGITS = 2, KEEP_STATUS_BAD_PIXELS = 4, KEEP_STATUS_BAD_FLOW = 8 };
int keep_status(keep_handle h, int clear, int* status_out);

/* Number of kernels the engine launched since creation (bench.py `gpu_launches`). */
long long keep_launch_count(keep_handle h);

/* Per-launch CUDA-event timing of the conv/GEMM kernel family (bench.py's roofline leg): enable, run
 * keep_forward, then read out8 = {launches, ms, GFLOP, algorithmic GB} for the CUDA-core path followed by the
 * same four numbers for the tcgen05 path.  Reading synchronises the device and keeps the samples. */
int keep_profile_enable(keep_handle h, int enable);
int keep_profile_read(keep_handle h, double* out8);
int keep_profile_dump(keep_handle h, const char* csv_path); /* one row per conv/GEMM launch: shape, ms */

/* Host-side plan of one call, no device needed (works on a KEEP_FLAG_PLAN_ONLY engine): one text line per conv / linear /
 * GroupNorm / LayerNorm / attention op of a forward over `clips` clips of T frames (clips > 1: the lockstep path) with its
 * shape and kernel choice.  CPU test tier; replaces nothing in the reference. */
int keep_plan_dump(keep_handle h, int clips, int T, const char* path);

/* ---- test hooks (stage-wise teacher forcing and intermediate capture; tests/ only) -------------
 * what ∈ {"flows" (T-1,512,512,2) f32, "z_codes" (T,16,16,256) f32 NHWC, "gains" (T,256) f32,
 *         "logits" (T,256,1024) f32, "codes" (T,256) i32, "prev" (T,3,512,512) f32 NCHW (force only)}.
 * keep_debug_force with bytes == 0 clears the forcing.  Data is host memory. */
int keep_debug_capture(keep_handle h, int enable);
int keep_debug_force(keep_handle h, const char* what, const void* host_data, size_t bytes);
long long keep_debug_read(keep_handle h, const char* what, void* host_data, size_t bytes);

/* ---- op-level entry points (tests/ only): each runs ONE kernel family on device pointers --------
 * conv: x (n,h,w,cin) NHWC fp32, weight OIHW host fp32, bias host or NULL; pads (t,l,b,r); `up` nearest
 * factor; pre_scale/pre_shift (n,cin) device or NULL; res (n,ho,wo,cout) device or NULL; out device fp32.
 * use_tc: 0 exact-fp32 CUDA cores | 1 tcgen05, fp16 operands | 3 tcgen05, split precision | 19 (3|16) split precision with
 * bf16 activation pairs (KEEP_FLAG_TC_WIDE) | 4 the Cin=3 / Cout<=4 stem and head kernels. */
int keepop_conv2d(int use_tc, const float* x_dev, int n, int h, int w, int cin, const float* weight_host, const float* bias_host,
                  int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r, int up,
                  const float* pre_scale_dev, const float* pre_shift_dev, int pre_act, int act, const float* res_dev,
                  float* out_dev, void* stream);
/* same convolution on the tcgen05 path (use_tc 1 | 3) with the GroupNorm(32, eps 1e-6) statistics of its OUTPUT emitted by the
 * producing kernel (conv epilogue, or the split-K reduce) and finalized into the per-(n, channel) affine the next layer's
 * operand producers apply: scale_dev / shift_dev (n, cout), y = x * scale + shift  (normalize(), vqgan_arch.py:16-17) */
int keepop_conv2d_gn(int use_tc, const float* x_dev, int n, int h, int w, int cin, const float* weight_host, const float* bias_host,
                     int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r, int up,
                     const float* pre_scale_dev, const float* pre_shift_dev, int pre_act, int act, const float* res_dev,
                     float* out_dev, const float* gn_gamma_dev, const float* gn_beta_dev, float* gn_scale_dev, float* gn_shift_dev,
                     void* stream);
/* linear layer y = x W^T + b (+ res) over `rows` token rows on the split-precision tcgen05 path whose split-K reduce kernel also
 * writes the LayerNorm of every output row: ln_out = LN(y) * g + b (eps), and, when add2_dev is given, ln_out2 = ln_out +
 * add2[row % add2_rows] (the code transformer's norm -> (+ position_emb) pattern, keep_arch.py:423-440).  weight_host (cout, cin);
 * fails when the shape does not split along K (nothing to fuse into). */
int keepop_linear_ln(const float* x_dev, int rows, int cin, const float* weight_host, const float* bias_host, int cout,
                     const float* res_dev, const float* ln_g_dev, const float* ln_b_dev, float eps, const float* add2_dev, int add2_rows,
                     float* out_dev, float* ln_out_dev, float* ln_out2_dev, void* stream);
/* debug timeline (tools/timeline.py): every kernel appends %globaltimer at its start to dev_buf (uint64[1 + 65536], [0] = count;
 * NULL = off); the launch log records kernel names and grids on the host in enqueue order */
int keepop_kernel_stamps(unsigned long long* dev_buf);
int keepop_launch_log(int enable);
int keepop_launch_log_dump(const char* path);
int keepop_tc_trace(long long* dev_buf_160_i64); /* debug: per-role clock64 timeline of CTA 0 of the tcgen05 kernel */
/* tests: 1 / 0 = the attention hooks below pre-convert K / V with the pack kernel (TMA-fed stages) or let every query tile convert
 * them itself; -1 = follow KEEP_ATTN_PACK (default on) */
int keepop_attention_pack_mode(int mode);
int keepop_attn_trace(long long* dev_buf_320_i64); /* debug: the same for the fused attention kernel (8 roles x 40 stamps) */
int keepop_groupnorm_affine(const float* x_dev, int n, int hw, int c, int groups, float eps, const float* gamma_dev,
                            const float* beta_dev, float* scale_dev, float* shift_dev, void* stream);
int keepop_layernorm(const float* x_dev, int rows, int c, const float* g_dev, const float* b_dev, float eps, float* out_dev,
                     void* stream);
/* multi-head attention on packed (nb*L, heads*dh) fp32 matrices */
int keepop_attention(const float* q_dev, const float* k_dev, const float* v_dev, int nb, int Lq, int Lk, int heads, int dh,
                     float scale, float* out_dev, void* stream);
/* fused attention on tcgen05 (QK^T -> softmax -> PV in one kernel, scores never in HBM): q (nb, Lq, dh), k / v (nb, Lk, dh),
 * out (nb, Lq, dh) fp32 contiguous; region_dev: optional (n_win, Lk) uint8 region ids of a shifted swin-window layer (adds -100
 * where region[q] != region[k]; batch z uses row z % n_win; needs Lq == Lk), gmflow/transformer.py:19-43,88-91.  dh = 128,
 * Lq % 128 == 0, Lk % 64 == 0, Lk <= 1024.  Replaces softmax(q @ k^T * scale + mask) @ v (gmflow/transformer.py:8-16,78-98). */
int keepop_attention_fused(const float* q_dev, const float* k_dev, const float* v_dev, int nb, int Lq, int Lk, int dh, float scale,
                           const unsigned char* region_dev, int n_win, float* out_dev, void* stream);
/* the same kernel in multi-head mode, operands packed like keepop_attention: (nb * L, heads * dh) fp32, dh = 64 or 128 */
int keepop_attention_fused_heads(const float* q_dev, const float* k_dev, const float* v_dev, int nb, int Lq, int Lk, int heads, int dh,
                                 float scale, float* out_dev, void* stream);
/* the same kernel in swin-window mode: q / k / v / out are whole (nimg, map_w * map_w, dh) token maps; the partition into
 * (map_w / wsz)^2 windows of wsz x wsz tokens, the cyclic shift (torch.roll by -shift on both axes before, +shift after) and
 * the merge are index math inside the kernel; region_dev: ((map_w / wsz)^2, wsz^2) uint8 region ids when shift > 0, else
 * NULL.  Replaces split_feature + roll + attention + merge_splits + roll of gmflow/transformer.py:78-103. */
int keepop_attention_window(const float* q_dev, const float* k_dev, const float* v_dev, int nimg, int map_w, int wsz, int shift, int dh,
                            float scale, const unsigned char* region_dev, float* out_dev, void* stream);
int keepop_flow_warp(const float* img_dev, const float* flow_dev, float* out_dev, int n, int h, int w, int c, void* stream);
int keepop_convex_upsample8(const float* mask_dev, const float* flow_dev, float* out_dev, int n, int h, int w, void* stream);
int keepop_window_sine_pos(float* x_dev, int n, int h, int w, int c, int splits, void* stream);
int keepop_argmax_gather(const float* logits_dev, int tokens, int ncodes, const float* codebook_dev, int cdim, int* idx_dev,
                         float* quant_dev, void* stream);
/* Nearest-neighbour codebook lookup = VectorQuantizer.forward of the VQGAN (modules/deps/wm_basicsr/archs/vqgan_arch.py:37-76;
 * SURVEY.md §8f N4): z_dev (tokens, cdim) fp32 token-major (the reference's z.permute(0, 2, 3, 1).view(-1, emb_dim)),
 * codebook_dev (ncodes, cdim) = quantize.embedding.weight.  Writes idx_dev[t] = argmin_j (|z_t|^2 + |e_j|^2 - 2 z_t.e_j)
 * (int32, ties -> lowest j, as torch.argmin), zq_dev (tokens, cdim) = e[idx] -- with straight_through != 0 the forward
 * value z + (e[idx] - z) the reference returns (:61) -- and dmin_dev[t] = the winning distance.  zq_dev / dmin_dev may be
 * NULL.  Asynchronous on `stream`.  Returns 0, or a negative code with keep_last_error() set (cdim must be a multiple
 * of 128, <= 384). */
int keepop_vq_nearest(const float* z_dev, int tokens, int cdim, const float* codebook_dev, int ncodes, int straight_through,
                      int* idx_dev, float* zq_dev, float* dmin_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KEEP_B200_H */
