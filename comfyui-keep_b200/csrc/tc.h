// keep_b200 — tcgen05 convolution kernel: shared declarations
#pragma once
#include "ops.h"

namespace keep {

struct TcConvArgs {
    const void* in0; const void* in1; int in0_dt, in1_dt; int c0, c1;
    int ld0;            // row stride of in0 in elements (= c0 unless the A operand is a column slice of a wider matrix)
    int n, h, w, up;
    const float* pre_scale; const float* pre_shift; int pre_act; int pre_exact;
    const __half* wt; const float* bias;
    long long wt_img_stride;   // halfs between per-image weight panel sets (attention GEMMs); 0 for convolutions
    int taps, cout, bn, ho, wo, ncb;
    int win, s2d, ncbr, pad_t, pad_l;   // window mode (3 / 2 / 1); stride-2 virtual space-to-depth parameters
    int tiles_y, tiles_x, ntile_n;
    FastDiv fd_ntile, fd_mtiles, fd_mtimg, fd_tilesx, fd_splitk;   // work-item decoding without hardware-less 64-bit divisions
    int cb_per;         // channel blocks per K split
    int gn_cpg_shift;   // log2(gn_cpg)
    int wt_static;      // 1: the weight panels were packed before this launch was enqueued behind its predecessor (engine
                        // creation): the loader may fetch them before griddepcontrol.wait
    int splitk; float* partial;
    int act; const void* res; int res_dt; void* out; int out_dt;
    // GroupNorm(32) statistics of the OUTPUT, emitted by the epilogue (no split-K): per (image, m tile, epilogue warp) one slot of
    // 32 x (sum, sum of squares) fp32 partials over the warp's 32 pixels -> gn_part[img][32][gn_P][2]; a fixed-order finalize
    // kernel turns them into the consumer's per-channel affine (norm.cu: gn_finalize_parts)
    float* gn_part; int gn_cpg; int gn_P;
    long long M;
    int tmem_cols;
    int sa_stages, sb_stages;
    int cluster_k;      // 1: split-K across the CTAs of a thread-block cluster, reduced over distributed shared memory
    int a_stat;         // 1: A-stationary walk over the N tiles of each (m tile, k split) item (1x1 layers)
    int w_resident;     // 1: the whole panel set of the layer stays in shared memory (loaded once per CTA)
    long long* trace;   // debug timeline (null = off)
    int swap_lbo_sbo;   // debug: KEEP_TC_SWAP_LBO_SBO=1
};

extern long long* g_tc_trace;
void tc_configure_device();      // per-device kernel attributes (idempotent)
bool tc_eligible(const ConvArgs& a);
int tc_pick_bn(int cout, long long m_tiles, int passes);
int tc_pick_splitk(long long m_tiles, int ntile_n, int ncb);
size_t tc_packed_weight_halfs(int cin, int cout, int taps, int bn, int passes);
bool tc_is_s2d(const ConvArgs& a);
int tc_cb(int passes);                        // input channels per A stage / weight panel: 64, or 32 in the split-precision mode
int tc_virtual_cin(const ConvArgs& a, int passes);   // 4 * ceil(cin/cb) * cb for the stride-2 mode, else cin
void tc_pack_weights(const float* w_oihw, int cout, int cin, int kh, int kw, int bn, int passes, __half* out, int wide = 0);
// s2d_pad < 0: plain repack; else stride-2 mode with pad_t = pad_l = s2d_pad (weights indexed [9*cin][cout])
void tc_repack_device(const float* w_kc, int cin, int cout, int taps, int bn, int passes, int s2d_pad, __half* out, cudaStream_t s, int wide = 0);
// pack B[z] (N x K, strided) into per-batch panel sets; returns halfs per batch (out == null: size query only)
size_t tc_pack_matrix(const float* src, long long bstride, int ld_n, int ld_k, int nbatch, int N, int K, int bn, int passes,
                      float alpha, __half* out, cudaStream_t s);
// partial: splitk * M * cout floats when splitk > 1
// returns the number of kernel launches made (1, or 2 with the separate split-K reduce)
int conv2d_tc(const ConvArgs& a, const __half* packed, int bn, int passes, int splitk, float* partial, int num_sms,
              cudaStream_t s);
// fixed-order split-K reduce + epilogue (conv_simt.cu)
// gn_part != null: also emit GroupNorm(32) partial statistics of the final values, one slot of 32 x (sum, sumsq) per
// 1024-element block -> gn_part[img][32][gn_P][2] (hw = output pixels per image; needs conv_gn_slots(...) > 0)
void splitk_reduce(const float* partial, int splitk, long long MN, int cout, const float* bias, int act, const void* res,
                   int res_dt, void* out, int out_dt, cudaStream_t s, float* gn_part = nullptr, int gn_P = 0, long long hw = 0,
                   const float* fin_gamma = nullptr, const float* fin_beta = nullptr, float* fin_scale = nullptr, float* fin_shift = nullptr,
                   int* fin_tickets = nullptr);
// same reduce + LayerNorm over each output row (see ConvArgs::ln_out); fp32 out, MN % 1024 == 0, cout in {128, 256, 512, 1024}
void splitk_reduce_ln(const float* partial, int splitk, long long MN, int cout, const float* bias, int act, const void* res, int res_dt,
                      float* out, const float* ln_g, const float* ln_b, float eps, float* ln_out, const float* add2, int add2_rows,
                      float* ln_out2, cudaStream_t s);
bool splitk_reduce_ln_eligible(long long MN, int cout);
// largest slot count per (image, group) the reduce kernel finalizes itself (one block walks 32 x P slots)
constexpr int kGnReduceFinalMaxP = 256;

}  // namespace keep
