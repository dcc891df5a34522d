// keep_b200 — the KEEP inference engine (host orchestration of the sm_100a kernels).
#pragma once
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/keep_b200.h"
#include "ops.h"
#include "tc.h"

namespace keep {

// first-fit free-list allocator over one contiguous device workspace.  All work is enqueued on a
// single stream, so a block may be reused as soon as its last consumer has been *enqueued*.
class Arena {
   public:
    void reset(char* base, size_t cap, bool dry);
    void* alloc(size_t bytes);
    void free(void* p);
    size_t peak() const { return peak_; }
    bool dry() const { return dry_; }

   private:
    struct Block { size_t off, size; bool used; };
    std::vector<Block> blocks_;
    char* base_ = nullptr;
    size_t cap_ = 0, peak_ = 0;
    bool dry_ = false;
};

struct DevArr { float* p = nullptr; long long numel = 0; int d[4] = {0, 0, 0, 0}; };
struct ConvW { const float* w = nullptr; const float* b = nullptr; int cin = 0, cout = 0, kh = 1, kw = 1; };
struct Aff { float* scale = nullptr; float* shift = nullptr; };

// LayerNorm of a split-K layer's output rows folded into its reduce kernel (ConvArgs::ln_out): filled in by Engine::conv() when
// the layer qualifies (done = true: out / out2 hold LN(y) and LN(y) + add2), else the caller runs Engine::ln() itself
struct LnFuse {
    std::string prefix;
    const float* add2 = nullptr; int add2_rows = 0;
    Tensor out, out2;
    bool done = false;
};

struct ConvOpt {
    int stride = 1, pad_t = 0, pad_l = 0, pad_b = 0, pad_r = 0, up = 1;
    const Aff* pre = nullptr; int pre_act = ACT_NONE;
    int act = ACT_NONE;
    const Tensor* res = nullptr;
    const Tensor* in1 = nullptr;
    int out_dt = -1;  // -1: engine feature-map dtype
    bool exact = false;  // force the exact-fp32 CUDA-core kernel even when the tcgen05 path is enabled
    bool want_stats = false;   // the output's next consumer is a GroupNorm(32): let the producing kernel emit its statistics
    LnFuse* ln = nullptr;      // the output's next consumer is this LayerNorm (fp32 token matrices)
    std::string stats_norm;    // ... and that norm's parameter prefix, when known: split-K layers finalize inside the reduce kernel
    ConvOpt& pad(int p) { pad_t = pad_l = pad_b = pad_r = p; return *this; }
};

class Engine {
   public:
    Engine(int device, const keep_weight_desc* w, int n_w, int flags);
    ~Engine();
    size_t workspace_bytes(int b, int T);
    void forward(const float* x_dev, int b, int T, void* out_dev, int out_dtype, void* ws, size_t ws_bytes, cudaStream_t s);
    void forward_u8(const unsigned char* x_u8_dev, int b, int T, unsigned char* out_u8_dev, void* ws, size_t ws_bytes, cudaStream_t s);
    // test hooks (stage-wise teacher forcing / intermediate capture)
    void force(const std::string& what, const void* host, size_t bytes);
    size_t read(const std::string& what, void* host, size_t bytes);
    void set_capture(bool on) { capture_ = on; }
    int status(bool clear);   // sticky non-finite bits (KEEP_STATUS_*), synchronises
    void set_batch_clips(int n) { batch_max_ = n < 1 ? 1 : (n > 8 ? 8 : n); }
    // host-side plan of one call (dry run, no device): one line per conv / linear / GroupNorm / LayerNorm / attention op with
    // its shape and the kernel choice (CPU test tier: control flow of both configs and of the lockstep path)
    void plan_dump(int nb, int T, const char* path);
    // per-launch CUDA-event timing of the conv/GEMM kernel family (bench.py roofline)
    void set_profile(bool on);
    void profile_read(double* out8);
    void profile_dump(const char* path);
    int device() const { return device_; }
    long long launches() const { return launches_; }

    // building blocks are public so the op-level C-ABI test hooks can drive them
    Tensor talloc(int n, int h, int w, int c, int dt);
    void tfree(Tensor& t);
    void afree(Aff& a);
    Tensor conv(const Tensor& x, const ConvW& cw, const ConvOpt& o);
    Tensor conv(const Tensor& x, const std::string& prefix, const ConvOpt& o) { return conv(x, convw(prefix), o); }
    Tensor linear(const Tensor& x, const std::string& prefix, int act = ACT_NONE, const Tensor* res = nullptr, LnFuse* lnf = nullptr);
    Aff gn(const Tensor& x, const std::string& prefix, const Tensor* x2 = nullptr);
    Aff inorm(const Tensor& x);
    Tensor ln(const Tensor& x, const std::string& prefix, const Tensor* res = nullptr, const float* add2 = nullptr,
              int add2_rows = 0, Tensor* out2 = nullptr);
    // out_stats: the block's output is normalised next (GroupNorm) -> its last conv emits the statistics
    Tensor res_block(const Tensor& x, const std::string& p, const Tensor* x2 = nullptr, bool out_stats = false, const std::string& out_norm = "");
    Tensor attn_block(const Tensor& x, const std::string& p, bool out_stats = false, const std::string& out_norm = "");
    Tensor encoder(const Tensor& img, const std::string& p, const std::function<void(int, const Tensor&)>& tap);
    // multi-head attention on (rows, ld) matrices; writes (nb*Lq, heads*dh)
    Tensor mha(const float* q, int ldq, long long sq, const float* k, int ldk, long long sk, const float* v, int ldv, long long sv,
               int nb, int Lq, int Lk, int heads, int dh, float scale);
    Tensor gemm_nt_tc(const float* A, int nb, int M, int K, const float* B, long long b_bstride, int ld_n, int ld_k, int N, float alpha, int lda = 0);
    // multi-head attention on token-major (L, heads * dh) fp32 matrices, all three contractions on the tcgen05 kernel
    Tensor mha_tc(const float* q, const float* k, const float* v, int Lq, int Lk, int heads, int dh, float scale);
    ConvW convw(const std::string& prefix) const;
    const float* warr(const std::string& key) const;
    bool has(const std::string& key) const { return W_.count(key) != 0; }
    void begin(void* ws, size_t ws_bytes, cudaStream_t s, bool dry);
    cudaStream_t stream() const { return s_; }
    Arena& arena() { return *ar_; }

   private:
    void forward_clip(const float* x_dev, int T, void* out_dev, int out_dtype);
    // KEEP_FLAG_BATCH_CLIPS: nb clips advance through the per-frame recurrence in lockstep (one hq_encoder / code transformer /
    // generator pass over a batch of nb per frame index)
    void forward_clips(const float* x_dev, int nb, int T, void* out_dev, int out_dtype);
    void forward_batched(const float* x_dev, int b, int T, void* out_dev, int out_dtype, cudaStream_t s);
    size_t plan_clips(int nb, int T);
    void gmflow(const float* x_nchw, int T, float* flows, int p_lo = 0, int p_hi = -1);
    void gm_resblock(Tensor& x, Aff* x_aff, const std::string& p, int stride);
    void gm_layer(Tensor& src, const Tensor& tgt, const std::string& p, int nimg, bool shift, bool ffn);
    Tensor kalman_gains(const Tensor& z_codes, int T);
    Tensor code_transformer(const Tensor& z_hat, int frame);
    Tensor cft(const Tensor& enc, const Tensor& dec, const std::string& p);
    Tensor cfa(const Tensor& cur, const Tensor& prev, const std::string& p);
    Tensor generator(const Tensor& quant, int frame, Tensor taps[6], Tensor cfa_prev[6]);
    void pack_weights(const keep_weight_desc* w, int n_w);
    void add_arr(const std::string& key, const std::vector<float>& host, int d0, int d1 = 0, int d2 = 0, int d3 = 0);

    int device_ = 0, flags_ = 0;
    // feature sizes (16, 32, 64, 128, 256, 512) that own cft.<s>.* / cfa.<s>.* tensors: 'KEEP' fuses CFT at 16/32/64,
    // 'Asian' at 32/64/128/256 (modules/utils.py:46,62); CFA at 16/32 in both
    bool cft_on_[6] = {false, false, false, false, false, false};
    bool cfa_on_[6] = {false, false, false, false, false, false};
    bool dry_only_ = false;  // KEEP_FLAG_PLAN_ONLY: host-side planning only (workspace sizing, key checks), no device
    int adt_ = F32;  // feature-map storage dtype
    std::unordered_map<std::string, DevArr> W_;
    std::vector<std::pair<std::string, std::vector<float>>> staging_;
    // bulk of the state dict (every plain conv / linear weight): uploaded raw and transposed on the device at creation
    struct PackJob { std::string key; const float* src; int O, I, taps; };
    std::vector<PackJob> jobs_;
    void add_job(const std::string& key, const float* src, int O, int I, int kh, int kw);
    float* wpool_ = nullptr;
    float* u8_stage_ = nullptr; size_t u8_stage_bytes_ = 0;   // fp32 copy of a uint8 input clip (forward_u8)
    int* status_ = nullptr;      // sticky non-finite status word (device)
    cudaEvent_t ev_last_ = nullptr;   // tail of the previous forward call (cross-stream ordering of engine-owned buffers)
    int* gn_tickets_ = nullptr;  // GroupNorm fused-finalize arrival counters [2 streams][gn_ticket_count()]
    int* region_ = nullptr;      // GMFlow shifted-window region ids [4][1024]
    unsigned char* region8_ = nullptr;   // the same table as bytes (fused attention kernel)
    float* grid64_ = nullptr;    // GMFlow coordinate grid (4096, 2)
    Arena arena_, arena2_;          // main / side-branch (GMFlow) workspaces
    Arena* ar_ = &arena_;           // arena of the branch currently being enqueued
    cudaStream_t s_main_ = nullptr, side_ = nullptr;
    cudaEvent_t ev_fork_ = nullptr;
    std::vector<cudaEvent_t> ev_flow_;   // one per GMFlow chunk of 4 pairs (per clip in the lockstep mode)
    int ev_flow_base_ = 0;               // first event of the clip whose GMFlow is being enqueued
    int batch_max_ = 2;                  // KEEP_FLAG_BATCH_CLIPS: clips per lockstep group (KEEP_BATCH_MAX)
    size_t side_bytes_ = 0;
    int main_cap_ = 148;                 // grid cap of main-stream persistent kernels (lowered while GMFlow overlaps)
    int side_sms_ = 64;                   // grid cap of persistent kernels on the side branch (measured: 64 -> 164.4, 100 -> 161.9, 148 -> 159.6 frames/s)
    std::unordered_map<int, size_t> side_cache_;
    cudaStream_t s_ = nullptr;
    long long launches_ = 0;
    // engine-owned workspace (used when the caller passes none)
    void* own_ws_ = nullptr; size_t own_ws_bytes_ = 0;
    std::unordered_map<int, size_t> ws_cache_;
    // tcgen05 path: fp16 weight panels, packed on first use, keyed by the fp32 weight pointer; variants are kept (a layer can be
    // asked for a second N tile when the per-clip and the lockstep path alternate), nothing is freed before the engine dies
    struct TcW { __half* p = nullptr; int bn = 0, passes = 0, wide = 0; bool pooled = false; };
    // panels of every layer a forward will run, packed at creation into one pool (no cudaMalloc / repack inside the first call,
    // so the first call can already capture the CUDA graph); the dry run of a 2-frame clip lists the variants
    struct TcReq { ConvW cw; int bn, passes, s2d_pad, wide; };
    std::vector<TcReq>* tc_collect_ = nullptr;
    __half* tcw_pool_ = nullptr;
    bool prepacked_ = false;
    void prepack_tc_weights();
    std::unordered_map<const float*, std::vector<TcW>> tcw_;   // every (N tile, passes, wide) variant a layer has been run with
    const __half* tc_weights(const ConvW& cw, int bn, int passes, int s2d_pad, int wide = 0, bool* pooled = nullptr);
    bool wide_scope_ = false; // KEEP_FLAG_TC_WIDE and inside generator(): raw-input feature-map layers use bf16 activation pairs
    int pass_override_ = 0;   // != 0: operand passes for the layers being enqueued (generator tail experiment)
    int tc_passes_ = 1;   // 1: fp16 operands; 3: split-precision (fp32-grade) tensor-core mode
    int num_sms_ = 148;
    // CUDA graph of one clip forward (KEEP_FLAG_CUDA_GRAPH): ~1800 launches per frame collapse into one graph launch
    struct ClipGraph { cudaGraphExec_t exec = nullptr; void* ws = nullptr; int out_dtype = 0; };
    std::unordered_map<int, ClipGraph> graphs_;   // keyed by T
    float* gx_ = nullptr; void* gout_ = nullptr; size_t gx_bytes_ = 0, gout_bytes_ = 0;   // static in/out staging for replays
    cudaStream_t gs_ = nullptr; cudaEvent_t ev_in_ = nullptr, ev_out_ = nullptr;
    std::unordered_map<int, long long> launches_per_clip_;
    std::unordered_map<int, int> eager_runs_;     // per T: eager forwards done (weights packed, allocations warmed)
    // debug capture / forcing
    struct Cap { void* p = nullptr; size_t bytes = 0; };
    std::unordered_map<std::string, Cap> cap_;      // device buffers holding last forward's intermediates
    std::unordered_map<std::string, Cap> forced_;   // device buffers with forced values
    bool capture_ = false;
    std::vector<std::string>* plan_ = nullptr;   // plan_dump(): op trace of the dry run in progress
    bool profile_ = false;
    struct Prof { cudaEvent_t a, b; double flops, bytes; int tag; int m, k, n, kh, splitk, bn; };
    std::vector<Prof> prof_;
    std::vector<cudaEvent_t> ev_pool_;
    cudaEvent_t get_event();
    void capture(const std::string& name, const void* dev, size_t bytes);
};

}  // namespace keep
