// keep_b200 — host-side launchers for the hand-written kernels (all on one stream).
#pragma once
#include "../../include/keep_b200.h"
#include "common.h"

namespace keep {

// ------------------------------------------------------------------------------------------
// convolution / linear as implicit GEMM over NHWC:  M = n*ho*wo pixels, N = cout, K = kh*kw*cin
// ------------------------------------------------------------------------------------------
struct ConvArgs {
    // input: up to two NHWC sources concatenated along channels (torch.cat([enc, dec], 1) of the
    // CFT block, keep_arch.py:467, is never materialised)
    const void* in0 = nullptr; int in0_dt = F32; int c0 = 0;
    int ld0 = 0;                   // tcgen05 path, 1x1 layers: row stride of in0 in elements (0: dense rows of c0) -- a column slice of a wider matrix
    const void* in1 = nullptr; int in1_dt = F32; int c1 = 0;
    int n = 0, h = 0, w = 0;     // physical input size
    int up = 1;                  // nearest-neighbour upsample factor applied on the fly (vqgan_arch.py:149)
    // prologue: per-(n, channel) affine (GroupNorm / InstanceNorm apply) then activation
    const float* pre_scale = nullptr; const float* pre_shift = nullptr; int pre_act = ACT_NONE;
    // weights packed [kh*kw*cin][cout] fp32 (k = (ky*kw + kx)*cin + ci), bias [cout] or null
    const float* wt = nullptr; const float* bias = nullptr;
    int kh = 1, kw = 1, stride = 1, pad_t = 0, pad_l = 0, cout = 0;
    int ho = 0, wo = 0;
    // epilogue: v = act(acc + bias) ; v += res
    int act = ACT_NONE;
    const void* res = nullptr; int res_dt = F32;
    void* out = nullptr; int out_dt = F32;
    // split-K (deterministic: partials to workspace, fixed-order reduce)
    int splitk = 1; float* partial = nullptr;
    int a_wide = 0;                // tcgen05 split-precision path: bf16 activation pairs (fp32 range) instead of fp16 pairs
    int pre_exact = 0;             // tcgen05 fp16-operand path: exact fp32 swish before the fp16 rounding (default: packed tanh.approx)
    long long wt_img_stride = 0;   // tcgen05 path only: per-image packed weight sets (batched A*B^T for attention)
    // tcgen05 path only: GroupNorm(32) partial statistics of the output, written by the producing kernel (the conv epilogue,
    // or the split-K reduce) as gn_part[n][32][gn_P][2] fp32 (sum, sum of squares) -- see conv_gn_slots()
    float* gn_part = nullptr; int gn_P = 0;
    // split-K layers with few slots: the reduce kernel's last block per image (ticket) also finalizes them into the consumer's
    // affine (gn_fin_scale / shift [n][cout], gamma / beta of the CONSUMING norm) -- no gn_finalize_parts launch.  gn_tickets:
    // >= n zeroed ints, left zeroed.
    const float* gn_fin_gamma = nullptr; const float* gn_fin_beta = nullptr; float* gn_fin_scale = nullptr; float* gn_fin_shift = nullptr;
    int* gn_tickets = nullptr;
    // split-K layers whose output rows (cout = 128 / 256 / 512 / 1024 channels) feed a LayerNorm: the reduce kernel also writes
    // ln_out = LN(out) * g + b and, optionally, ln_out2 = ln_out + ln_add2[row % ln_add2_rows] (a positional embedding)
    const float* ln_g = nullptr; const float* ln_b = nullptr; float ln_eps = 1e-5f; float* ln_out = nullptr;
    const float* ln_add2 = nullptr; int ln_add2_rows = 1; float* ln_out2 = nullptr;
    int wt_static = 0;             // tcgen05 path only: packed weights are older than the stream's previous kernel (see TcConvArgs)
    int no_reduce = 0;             // tcgen05 path only: leave the splitk partial tiles as the result -- with K = heads x dh and
                                   // splitk = heads, partial[h] IS the per-head product Q_h K_h^T (multi-head attention scores)
};
// slots per image the producing kernel writes for a layer run with `splitk` K-splits (0: this layer cannot emit statistics)
int conv_gn_slots(const ConvArgs& a, int splitk);
void conv2d_simt(const ConvArgs& a, cudaStream_t s);
// picks split-K for small-M layers; `scratch` must hold conv_splitk_scratch_floats(a) floats when splitk>1
int conv_pick_splitk(const ConvArgs& a);
// bandwidth-bound ends of the conv stacks: Cin = 3 stems (3x3 s1, 7x7 s2) and Cout <= 4 heads (3x3 s1)
bool conv_small_eligible(const ConvArgs& a);
void conv2d_small(const ConvArgs& a, cudaStream_t s);
void conv_small_configure_device();   // per-device kernel attributes (idempotent)

// ------------------------------------------------------------------------------------------
// batched strided GEMM (fp32), used for attention scores / PV
//   C[z][m][n] = alpha * sum_k A[z][m][k] * (transB ? B[z][n][k] : B[z][k][n])  (+ C if accumulate)
//   z = (z0, z1, z2) with per-operand strides (elements)
// ------------------------------------------------------------------------------------------
struct BGemmArgs {
    const float* A = nullptr; const float* B = nullptr; float* C = nullptr;
    int M = 0, N = 0, K = 0;
    int lda = 0, ldb = 0, ldc = 0;
    int transB = 1;
    float alpha = 1.0f; int accumulate = 0;
    int nz0 = 1, nz1 = 1, nz2 = 1;
    long long sA[3] = {0, 0, 0}, sB[3] = {0, 0, 0}, sC[3] = {0, 0, 0};
};
void bgemm_simt(const BGemmArgs& a, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// fused attention on tcgen05 (attn_tcgen05.cu): out[z] = softmax(q[z] k[z]^T * scale (+ region mask)) v[z], scores never in HBM.
//   q (Lq, dh), k / v (Lk, dh), out (Lq, dh) fp32 row-major with row strides ld* and batch strides *_bs (elements);
//   region: [n_win][Lk] uint8 region ids (shifted swin windows: -100 where region[q] != region[k]; batch z uses row z % n_win)
//   or null.  Supported: dh = 128 or 64, Lq % 128 == 0, Lk % 64 == 0, Lk <= 1024 (attention_tc_eligible).
// ------------------------------------------------------------------------------------------
bool attention_tc_eligible(int Lq, int Lk, int dh);
void attention_tc_configure_device();
void attention_tc(const float* q, int ldq, long long q_bs, const float* k, int ldk, long long k_bs, const float* v, int ldv, long long v_bs,
                  float* out, int ldo, long long o_bs, int nb, int Lq, int Lk, int dh, float scale, const unsigned char* region, int n_win,
                  cudaStream_t s, int win_side = 0, int wsz = 0, int map_w = 0, int shift = 0, int heads = 1, void* kv_pack = nullptr);
// kv_pack != null (attention_tc_pack_bytes(nb, Lk, dh) bytes, 128-byte aligned): K / V are converted to the kernel's (hi, lo) fp16
// stage images ONCE by a pack kernel and stream into the query tiles by 1-D TMA -- otherwise each of a batch element's Lq / 128
// query tiles converts the whole K and V itself (worth it from ~4 query tiles per batch element on)
size_t attention_tc_pack_bytes(int nb, int Lk, int dh);
// (heads > 1: nb = batches * heads, head h of batch b reads / writes columns [h * dh, (h + 1) * dh) of batch b's rows)
// (win_side > 0: swin window mode -- q / k / v / out are whole (map_w x map_w)-token maps, *_bs = image strides, nb = images *
//  win_side^2, window partition + cyclic shift + merge are index math inside the kernel; wsz = window side, Lq = Lk = wsz^2)

// ------------------------------------------------------------------------------------------
// normalisation
// ------------------------------------------------------------------------------------------
// per-(n, group) statistics over an NHWC tensor -> per-(n, channel) affine  y = x*scale + shift
//   groups of `cpg` consecutive channels; gamma/beta may be null (InstanceNorm, affine-free)
//   writes scale/shift at [n][c_total] + c_off  (c_total >= c, for concatenated inputs)
//   scratch: gn_scratch_doubles(...) doubles
//   ticket_buf: gn_ticket_count() zero-initialised ints owned by the caller, one buffer per concurrently used stream --
//   large maps then fuse the finalize stage into the partial kernel; null: two kernels
size_t gn_scratch_doubles(int n, int hw, int c);
int gn_ticket_count();
void gn_warmup();   // allocate the per-device ticket array (must happen outside CUDA-graph capture)
void groupnorm_affine(const void* x, int dt, int n, int hw, int c, int cpg, float eps,
                      const float* gamma, const float* beta, float* scale, float* shift,
                      int c_total, int c_off, double* scratch, cudaStream_t s, int* ticket_buf = nullptr);
// LayerNorm over the last dim of (rows, c) fp32; out = LN(x)*g+b (+ res) ; out2 = out + add2[row % add2_rows]
// statistics emitted by a producing kernel (ConvArgs::gn_part) -> per-(n, channel) affine; one small launch, fixed order
void gn_finalize_parts(const float* part, int n, int P, int hw, int c, float eps, const float* gamma, const float* beta,
                       float* scale, float* shift, cudaStream_t s);
void layernorm(const float* x, int rows, int c, const float* g, const float* b, float eps,
               const float* res, float* out, const float* add2, int add2_rows, float* out2, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// sticky non-finite status word (keep_status, include/keep_b200.h): kernels at the joints of the path OR a bit into an
// engine-owned device int when they see inf / NaN -- no host synchronisation, the caller reads it after its own sync
// ------------------------------------------------------------------------------------------
// (bits: KEEP_STATUS_* in the public header)

// ------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------
// out = act_o( act_a(A*sa+ba) + act_b(B*sb+bb) ), per-(n,c) affines optional, B optional
struct EwArgs {
    const void* A = nullptr; int a_dt = F32; const float* sa = nullptr; const float* ba = nullptr; int act_a = ACT_NONE;
    const void* B = nullptr; int b_dt = F32; const float* sb = nullptr; const float* bb = nullptr; int act_b = ACT_NONE;
    int act_o = ACT_NONE;
    void* out = nullptr; int o_dt = F32;
    int n = 0, hw = 0, c = 0;
};
void elementwise(const EwArgs& a, cudaStream_t s);
// CFT combine: out = dec + cond*(dec*scale + shift)          keep_arch.py:470-471
void cft_combine(const void* dec, int dec_dt, const void* scale, const void* shift, int ss_dt, float cond,
                 void* out, int o_dt, size_t numel, cudaStream_t s);
// GEGLU: in (rows, 2*inner) -> out (rows, inner) = in[:, :inner] * gelu(in[:, inner:])
void geglu(const float* in, float* out, int rows, int inner, cudaStream_t s);
// softmax over rows of length L (in place), optional swin region mask: add -100 where region[q] != region[k]
//   rows are organised as (batch, Lq); region ids indexed [(batch % n_win)][token]
void softmax_rows(float* s, long long rows, int L, const int* region, int n_win, int Lq, cudaStream_t s_);
// out[row][0:2] = sum_k softmax(s[row])[k] * v[(row / Lq)][k][0:2]   (- sub[row % Lq][0:2] if sub)
//   v is indexed [(row / Lq) * v_bstride + 2*k]  (v_bstride = 0 shares one value table across batches)
void softmax_expect2(const float* s, long long rows, int L, int Lq, const float* v, long long v_bstride, const float* sub,
                     float* out, cudaStream_t s_);
void kalman_update(const float* z, const float* zp, const float* gain, float* out, int pixels, int c, cudaStream_t s, int* status = nullptr);
// logits (tokens, ncodes) -> idx (tokens) int32, quant (tokens, cdim) = codebook[idx]; forced idx optional
void argmax_gather(const float* logits, int tokens, int ncodes, const float* codebook, int cdim,
                   const int* forced_idx, int* idx_out, void* quant, int q_dt, cudaStream_t s, int* status = nullptr);
// VectorQuantizer.forward (vqgan_arch.py:37-76): z (tokens, cdim) -> idx = argmin_j ||z - e_j||^2 (int32, ties -> lowest j),
// zq (tokens, cdim) = e[idx] (straight_through: the forward value z + (e[idx] - z)), dmin (tokens) the winning distance;
// zq / dmin optional
void vq_nearest(const float* z, int tokens, int cdim, const float* codebook, int ncodes, int straight_through, int* idx_out,
                float* zq, float* dmin_out, cudaStream_t s);
// sparse-causal K/V gather: out[(b,f)][0:L] = kv[(b,0)], out[(b,f)][L:2L] = kv[(b,max(f-1,0))]  keep_arch.py:704-716
void sparse_causal_gather(const float* kv, float* out, int b, int T, int L, int c, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// layout / geometry
// ------------------------------------------------------------------------------------------
// x NCHW fp32 [-1,1] -> NHWC (dt): mode 0 copy ; mode 1 GMFlow normalisation ((x+1)/2 - mean)/std
void nchw_to_nhwc(const float* x, void* out, int o_dt, int n, int c, int h, int w, int mode, cudaStream_t s);
void nhwc_to_nchw(const void* x, int dt, void* out, int out_dt, int n, int c, int h, int w, cudaStream_t s, int* status = nullptr);
// caller-side conversions of keep_processor.py folded in: uint8 BGR HWC crops -> fp32 RGB NCHW in [-1, 1]
// (img2tensor(crop / 255., bgr2rgb=True) + normalize(0.5, 0.5)), and fp32 RGB NHWC -> uint8 BGR HWC (tensor2img, min_max (-1, 1))
void u8bgr_to_nchw_norm(const unsigned char* x, float* out, int n, int h, int w, cudaStream_t s);
void nhwc_to_u8bgr(const void* x, int dt, unsigned char* out, int n, int h, int w, cudaStream_t s, int* status = nullptr);
// bilinear warp (grid_sample bilinear / zeros / align_corners=True), img NHWC c channels, flow (n,h,w,2) px
void flow_warp(const void* img, int dt, const float* flow, void* out, int o_dt, int n, int h, int w, int c, cudaStream_t s, int* status = nullptr);
// GMFlow: add windowed sine position embedding in place on (n, h, w, c) fp32  (utils.py:66-86)
void add_window_sine_pos(float* x, int n, int h, int w, int c, int splits, cudaStream_t s);
// GMFlow window partition with optional cyclic shift: (n, h, w, c) -> (n*k*k, h/k*w/k, c) and back
void window_partition(const float* x, float* out, int n, int h, int w, int c, int k, int shift_h, int shift_w,
                      int ldx, cudaStream_t s);
void window_merge(const float* x, float* out, int n, int h, int w, int c, int k, int shift_h, int shift_w, cudaStream_t s);
// convex x8 upsampling (gmflow.py:74-88): mask (n,h,w,576), flow (n,h,w,2) -> (n,8h,8w,2)
void convex_upsample8(const float* mask, const float* flow, float* out, int n, int h, int w, cudaStream_t s);
// weight layout transform at engine creation: OIHW / (O, I) on the device -> [(tap * I + i)][o]
void oihw_to_kc(const float* w_dev, float* out_dev, int O, int I, int taps, cudaStream_t s);
// multi-head attention output (heads, rows, dh) -> token-major (rows, heads * dh)
void heads_to_tokens(const float* x, float* out, int heads, int rows, int dh, cudaStream_t s);
// concat along channels of two (rows, c) fp32 matrices / generic strided copy
void concat2(const float* a, int ca, const float* b, int cb, float* out, long long rows, cudaStream_t s);

}  // namespace keep
