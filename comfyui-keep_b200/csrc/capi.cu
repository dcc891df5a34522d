// keep_b200 — the C-ABI boundary (include/keep_b200.h): exceptions -> error codes, no torch types.
#include <mutex>
#include <vector>

#include "engine.h"

using namespace keep;

struct keep_engine_s {
    Engine* e;
    std::mutex mu;
};

static thread_local std::string g_err;

#define KEEP_API_BEGIN try {
#define KEEP_API_END                                   \
    }                                                  \
    catch (const std::exception& ex) {                 \
        g_err = ex.what();                             \
        cudaGetLastError();                            \
        return -1;                                     \
    }                                                  \
    catch (...) {                                      \
        g_err = "keep_b200: unknown error";            \
        return -2;                                     \
    }

int keepop_conv2d_tc(const ConvArgs& a, const float* w_oihw_host, int passes, cudaStream_t s, const float* gn_gamma = nullptr,
                     const float* gn_beta = nullptr, float* gn_scale = nullptr, float* gn_shift = nullptr);   // conv_tcgen05.cu

namespace keep {
void stamp_set_conv_simt(unsigned long long*); void stamp_set_conv_small(unsigned long long*); void stamp_set_conv_tc(unsigned long long*);
void stamp_set_gemm(unsigned long long*); void stamp_set_misc(unsigned long long*); void stamp_set_norm(unsigned long long*);
void stamp_set_attn(unsigned long long*);
extern long long* g_attn_trace;   // attn_tcgen05.cu
void launch_log_enable(bool on);
int launch_log_dump(const char* path);
}

extern "C" {

const char* keep_last_error(void) { return g_err.c_str(); }

int keep_create(keep_handle* out, int device, const keep_weight_desc* weights, int n_weights, int flags) {
    KEEP_API_BEGIN
    KEEP_CHECK(out && weights && n_weights > 0, "keep_create: null argument");
    *out = nullptr;
    keep_engine_s* h = new keep_engine_s();
    try {
        h->e = new Engine(device, weights, n_weights, flags);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    return 0;
    KEEP_API_END
}

size_t keep_workspace_bytes(keep_handle h, int b, int T) {
    try {
        if (!h) throw Error("keep_workspace_bytes: null handle");
        std::lock_guard<std::mutex> lk(h->mu);
        return h->e->workspace_bytes(b, T);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return 0;
    }
}

int keep_forward(keep_handle h, const float* x_dev, int b, int T, void* out_dev, int out_dtype, void* workspace,
                 size_t workspace_bytes, void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(h, "keep_forward: null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->e->forward(x_dev, b, T, out_dev, out_dtype, workspace, workspace_bytes, (cudaStream_t)stream);
    return 0;
    KEEP_API_END
}

int keep_set_batch_clips(keep_handle h, int max_clips) {
    KEEP_API_BEGIN
    KEEP_CHECK(h, "keep_set_batch_clips: null handle");
    KEEP_CHECK(max_clips >= 1 && max_clips <= 8, "keep_set_batch_clips: 1 <= max_clips <= 8 (got %d)", max_clips);
    std::lock_guard<std::mutex> lk(h->mu);
    h->e->set_batch_clips(max_clips);
    return 0;
    KEEP_API_END
}

int keep_forward_u8(keep_handle h, const unsigned char* x_u8_dev, int b, int T, unsigned char* out_u8_dev, void* workspace,
                    size_t workspace_bytes, void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(h, "keep_forward_u8: null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->e->forward_u8(x_u8_dev, b, T, out_u8_dev, workspace, workspace_bytes, (cudaStream_t)stream);
    return 0;
    KEEP_API_END
}

int keep_destroy(keep_handle h) {
    KEEP_API_BEGIN
    if (h) {
        delete h->e;
        delete h;
    }
    return 0;
    KEEP_API_END
}

long long keep_launch_count(keep_handle h) { return h ? h->e->launches() : 0; }

int keep_status(keep_handle h, int clear, int* status_out) {
    KEEP_API_BEGIN
    KEEP_CHECK(h && status_out, "keep_status: null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *status_out = h->e->status(clear != 0);
    return 0;
    KEEP_API_END
}

int keep_profile_enable(keep_handle h, int enable) {
    KEEP_API_BEGIN
    KEEP_CHECK(h, "null handle");
    h->e->set_profile(enable != 0);
    return 0;
    KEEP_API_END
}

int keep_profile_read(keep_handle h, double* out8) {
    KEEP_API_BEGIN
    KEEP_CHECK(h && out8, "null argument");
    h->e->profile_read(out8);
    return 0;
    KEEP_API_END
}

int keep_profile_dump(keep_handle h, const char* path) {
    KEEP_API_BEGIN
    KEEP_CHECK(h && path, "null argument");
    h->e->profile_dump(path);
    return 0;
    KEEP_API_END
}

int keep_plan_dump(keep_handle h, int clips, int T, const char* path) {
    KEEP_API_BEGIN
    KEEP_CHECK(h && path, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    h->e->plan_dump(clips, T, path);
    return 0;
    KEEP_API_END
}

int keep_debug_capture(keep_handle h, int enable) {
    KEEP_API_BEGIN
    KEEP_CHECK(h, "null handle");
    h->e->set_capture(enable != 0);
    return 0;
    KEEP_API_END
}

int keep_debug_force(keep_handle h, const char* what, const void* host_data, size_t bytes) {
    KEEP_API_BEGIN
    KEEP_CHECK(h && what, "null argument");
    h->e->force(what, host_data, bytes);
    return 0;
    KEEP_API_END
}

long long keep_debug_read(keep_handle h, const char* what, void* host_data, size_t bytes) {
    try {
        if (!h || !what) throw Error("null argument");
        return (long long)h->e->read(what, host_data, bytes);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
}

// ---------------------------------------------------------------------------------------------
// op-level hooks: a tiny weight-less Engine-free harness around the launchers
// ---------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 256)); }
    ~DevBuf() { cudaFree(p); }
};
}  // namespace

static int conv2d_op(int use_tc, const float* x_dev, int n, int h, int w, int cin, const float* weight_host, const float* bias_host,
                     int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r, int up,
                     const float* pre_scale_dev, const float* pre_shift_dev, int pre_act, int act, const float* res_dev,
                     float* out_dev, void* stream, const float* gn_gamma, const float* gn_beta, float* gn_scale, float* gn_shift) {
    KEEP_API_BEGIN
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<float> packed((size_t)cout * cin * kh * kw);
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i)
            for (int y = 0; y < kh; ++y)
                for (int x = 0; x < kw; ++x)
                    packed[((size_t)(y * kw + x) * cin + i) * cout + o] = weight_host[(((size_t)o * cin + i) * kh + y) * kw + x];
    DevBuf wd(packed.size() * 4), bd((size_t)cout * 4);
    CUDA_CHECK(cudaMemcpy(wd.p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    if (bias_host) CUDA_CHECK(cudaMemcpy(bd.p, bias_host, (size_t)cout * 4, cudaMemcpyHostToDevice));
    ConvArgs a;
    a.in0 = x_dev; a.c0 = cin; a.n = n; a.h = h; a.w = w; a.up = up;
    a.pre_scale = pre_scale_dev; a.pre_shift = pre_shift_dev; a.pre_act = pre_act;
    a.wt = (const float*)wd.p; a.bias = bias_host ? (const float*)bd.p : nullptr;
    a.kh = kh; a.kw = kw; a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l; a.cout = cout;
    a.ho = (h * up + pad_t + pad_b - kh) / stride + 1;
    a.wo = (w * up + pad_l + pad_r - kw) / stride + 1;
    a.act = act; a.res = res_dev; a.out = out_dev;
    if (use_tc == 4) {   // bandwidth-bound stem / head kernels
        conv2d_small(a, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
        return 0;
    }
    if (use_tc) {   // 1: fp16 operands; 3: split precision; 19 (= 3 | 16): split precision with bf16 activation pairs (KEEP_FLAG_TC_WIDE)
        a.a_wide = (use_tc & 16) ? 1 : 0;
        int rc = keepop_conv2d_tc(a, weight_host, (use_tc & 3) == 3 ? 3 : 1, s, gn_gamma, gn_beta, gn_scale, gn_shift);
        CUDA_CHECK(cudaStreamSynchronize(s));
        return rc;
    }
    KEEP_CHECK(!gn_scale, "keepop_conv2d_gn: statistics are emitted by the tcgen05 path only (use_tc = 1 or 3)");
    a.splitk = conv_pick_splitk(a);
    DevBuf part(a.splitk > 1 ? (size_t)a.splitk * n * a.ho * a.wo * cout * 4 : 0);
    a.partial = (float*)part.p;
    conv2d_simt(a, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
    KEEP_API_END
}

int keepop_conv2d(int use_tc, const float* x_dev, int n, int h, int w, int cin, const float* weight_host, const float* bias_host,
                  int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r, int up,
                  const float* pre_scale_dev, const float* pre_shift_dev, int pre_act, int act, const float* res_dev,
                  float* out_dev, void* stream) {
    return conv2d_op(use_tc, x_dev, n, h, w, cin, weight_host, bias_host, cout, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, up, pre_scale_dev,
                     pre_shift_dev, pre_act, act, res_dev, out_dev, stream, nullptr, nullptr, nullptr, nullptr);
}

int keepop_conv2d_gn(int use_tc, const float* x_dev, int n, int h, int w, int cin, const float* weight_host, const float* bias_host,
                     int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r, int up,
                     const float* pre_scale_dev, const float* pre_shift_dev, int pre_act, int act, const float* res_dev,
                     float* out_dev, const float* gn_gamma_dev, const float* gn_beta_dev, float* gn_scale_dev, float* gn_shift_dev,
                     void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(gn_gamma_dev && gn_beta_dev && gn_scale_dev && gn_shift_dev, "keepop_conv2d_gn: null GroupNorm argument");
    KEEP_API_END
    return conv2d_op(use_tc, x_dev, n, h, w, cin, weight_host, bias_host, cout, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, up, pre_scale_dev,
                     pre_shift_dev, pre_act, act, res_dev, out_dev, stream, gn_gamma_dev, gn_beta_dev, gn_scale_dev, gn_shift_dev);
}

int keepop_linear_ln(const float* x_dev, int rows, int cin, const float* weight_host, const float* bias_host, int cout,
                     const float* res_dev, const float* ln_g_dev, const float* ln_b_dev, float eps, const float* add2_dev, int add2_rows,
                     float* out_dev, float* ln_out_dev, float* ln_out2_dev, void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(ln_g_dev && ln_b_dev && ln_out_dev && (!ln_out2_dev || add2_dev), "keepop_linear_ln: null LayerNorm argument");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<float> packed((size_t)cout * cin);
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i) packed[(size_t)i * cout + o] = weight_host[(size_t)o * cin + i];
    DevBuf wd(packed.size() * 4), bd((size_t)cout * 4);
    CUDA_CHECK(cudaMemcpy(wd.p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    if (bias_host) CUDA_CHECK(cudaMemcpy(bd.p, bias_host, (size_t)cout * 4, cudaMemcpyHostToDevice));
    ConvArgs a;
    a.in0 = x_dev; a.c0 = cin; a.n = 1; a.h = rows; a.w = 1; a.up = 1;
    a.wt = (const float*)wd.p; a.bias = bias_host ? (const float*)bd.p : nullptr;
    a.kh = 1; a.kw = 1; a.stride = 1; a.cout = cout; a.ho = rows; a.wo = 1;
    a.res = res_dev; a.out = out_dev;
    a.ln_g = ln_g_dev; a.ln_b = ln_b_dev; a.ln_eps = eps; a.ln_out = ln_out_dev;
    a.ln_add2 = add2_dev; a.ln_add2_rows = add2_rows; a.ln_out2 = ln_out2_dev;
    int rc = keepop_conv2d_tc(a, weight_host, 3, s, nullptr, nullptr, nullptr, nullptr);
    CUDA_CHECK(cudaStreamSynchronize(s));
    return rc;
    KEEP_API_END
}

// debug: kernel-start timeline.  stamps = device buffer of 1 + 65536 uint64 (null = off); launch log = host-side names
int keepop_kernel_stamps(unsigned long long* dev_buf) {
    keep::stamp_set_conv_simt(dev_buf); keep::stamp_set_conv_small(dev_buf); keep::stamp_set_conv_tc(dev_buf);
    keep::stamp_set_gemm(dev_buf); keep::stamp_set_misc(dev_buf); keep::stamp_set_norm(dev_buf); keep::stamp_set_attn(dev_buf);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}
int keepop_launch_log(int enable) { keep::launch_log_enable(enable != 0); return 0; }
int keepop_launch_log_dump(const char* path) { return keep::launch_log_dump(path); }

// debug: point the tcgen05 kernel's role timeline at a device buffer of 160 int64 (null = off)
int keepop_tc_trace(long long* dev_buf) {
    keep::g_tc_trace = dev_buf;
    return 0;
}

int keepop_groupnorm_affine(const float* x_dev, int n, int hw, int c, int groups, float eps, const float* gamma_dev,
                            const float* beta_dev, float* scale_dev, float* shift_dev, void* stream) {
    KEEP_API_BEGIN
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf scratch(gn_scratch_doubles(n, hw, c) * 8);
    groupnorm_affine(x_dev, F32, n, hw, c, c / groups, eps, gamma_dev, beta_dev, scale_dev, shift_dev, c, 0, (double*)scratch.p, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
    KEEP_API_END
}

int keepop_layernorm(const float* x_dev, int rows, int c, const float* g_dev, const float* b_dev, float eps, float* out_dev,
                     void* stream) {
    KEEP_API_BEGIN
    layernorm(x_dev, rows, c, g_dev, b_dev, eps, nullptr, out_dev, nullptr, 0, nullptr, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

// op hooks: K / V pack scratch for attention_tc (mode -1: KEEP_ATTN_PACK from the environment, default on; 0 / 1: forced by the tests)
static int g_attn_pack_mode = -1;
int keepop_attention_pack_mode(int mode) { g_attn_pack_mode = mode; return 0; }
namespace {
struct AttnPack {
    void* p = nullptr;
    AttnPack(int nb, int Lk, int dh) {
        const bool on = g_attn_pack_mode >= 0 ? g_attn_pack_mode != 0 : !(getenv("KEEP_ATTN_PACK") && getenv("KEEP_ATTN_PACK")[0] == '0');
        if (on) CUDA_CHECK(cudaMalloc(&p, attention_tc_pack_bytes(nb, Lk, dh)));
    }
    ~AttnPack() { cudaFree(p); }
};
}  // namespace

int keepop_attention_fused(const float* q, const float* k, const float* v, int nb, int Lq, int Lk, int dh, float scale,
                           const unsigned char* region_dev, int n_win, float* out_dev, void* stream) {
    KEEP_API_BEGIN
    AttnPack pk(nb, Lk, dh);
    attention_tc(q, dh, (long long)Lq * dh, k, dh, (long long)Lk * dh, v, dh, (long long)Lk * dh, out_dev, dh, (long long)Lq * dh, nb, Lq, Lk, dh,
                 scale, region_dev, n_win, (cudaStream_t)stream, 0, 0, 0, 0, 1, pk.p);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_attn_trace(long long* dev_buf_320_i64) { keep::g_attn_trace = dev_buf_320_i64; return 0; }

int keepop_attention_fused_heads(const float* q, const float* k, const float* v, int nb, int Lq, int Lk, int heads, int dh, float scale,
                                 float* out_dev, void* stream) {
    KEEP_API_BEGIN
    const int D = heads * dh;
    AttnPack pk(nb * heads, Lk, dh);
    attention_tc(q, D, (long long)Lq * D, k, D, (long long)Lk * D, v, D, (long long)Lk * D, out_dev, D, (long long)Lq * D, nb * heads, Lq, Lk, dh, scale,
                 nullptr, 1, (cudaStream_t)stream, 0, 0, 0, 0, heads, pk.p);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_attention_window(const float* q, const float* k, const float* v, int nimg, int map_w, int wsz, int shift, int dh, float scale,
                            const unsigned char* region_dev, float* out_dev, void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(wsz > 0 && map_w % wsz == 0, "keepop_attention_window: bad geometry");
    const int side = map_w / wsz, L = wsz * wsz;
    const long long img = (long long)map_w * map_w * dh;
    AttnPack pk(nimg * side * side, L, dh);
    attention_tc(q, dh, img, k, dh, img, v, dh, img, out_dev, dh, img, nimg * side * side, L, L, dh, scale, region_dev, side * side,
                 (cudaStream_t)stream, side, wsz, map_w, shift, 1, pk.p);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_attention(const float* q, const float* k, const float* v, int nb, int Lq, int Lk, int heads, int dh, float scale,
                     float* out_dev, void* stream) {
    KEEP_API_BEGIN
    cudaStream_t s = (cudaStream_t)stream;
    const int D = heads * dh;
    DevBuf S((size_t)nb * heads * Lq * Lk * 4);
    BGemmArgs g;
    g.A = q; g.B = k; g.C = (float*)S.p;
    g.M = Lq; g.N = Lk; g.K = dh; g.lda = D; g.ldb = D; g.ldc = Lk; g.transB = 1; g.alpha = scale;
    g.nz0 = nb; g.nz1 = heads;
    g.sA[0] = (long long)Lq * D; g.sA[1] = dh; g.sB[0] = (long long)Lk * D; g.sB[1] = dh;
    g.sC[0] = (long long)heads * Lq * Lk; g.sC[1] = (long long)Lq * Lk;
    bgemm_simt(g, s);
    softmax_rows((float*)S.p, (long long)nb * heads * Lq, Lk, nullptr, 1, Lq, s);
    BGemmArgs m;
    m.A = (float*)S.p; m.B = v; m.C = out_dev;
    m.M = Lq; m.N = dh; m.K = Lk; m.lda = Lk; m.ldb = D; m.ldc = D; m.transB = 0;
    m.nz0 = nb; m.nz1 = heads;
    m.sA[0] = (long long)heads * Lq * Lk; m.sA[1] = (long long)Lq * Lk;
    m.sB[0] = (long long)Lk * D; m.sB[1] = dh; m.sC[0] = (long long)Lq * D; m.sC[1] = dh;
    bgemm_simt(m, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
    KEEP_API_END
}

int keepop_flow_warp(const float* img, const float* flow, float* out, int n, int h, int w, int c, void* stream) {
    KEEP_API_BEGIN
    flow_warp(img, F32, flow, out, F32, n, h, w, c, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_convex_upsample8(const float* mask, const float* flow, float* out, int n, int h, int w, void* stream) {
    KEEP_API_BEGIN
    convex_upsample8(mask, flow, out, n, h, w, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_window_sine_pos(float* x, int n, int h, int w, int c, int splits, void* stream) {
    KEEP_API_BEGIN
    add_window_sine_pos(x, n, h, w, c, splits, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_argmax_gather(const float* logits, int tokens, int ncodes, const float* codebook, int cdim, int* idx, float* quant,
                         void* stream) {
    KEEP_API_BEGIN
    argmax_gather(logits, tokens, ncodes, codebook, cdim, nullptr, idx, quant, F32, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
    KEEP_API_END
}

int keepop_vq_nearest(const float* z, int tokens, int cdim, const float* codebook, int ncodes, int straight_through, int* idx,
                      float* zq, float* dmin, void* stream) {
    KEEP_API_BEGIN
    KEEP_CHECK(z && codebook && idx, "keepop_vq_nearest: null tensor");
    vq_nearest(z, tokens, cdim, codebook, ncodes, straight_through, idx, zq, dmin, (cudaStream_t)stream);
    return 0;
    KEEP_API_END
}

}  // extern "C"
