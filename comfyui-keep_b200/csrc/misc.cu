// keep_b200 — memory-bound kernels of the KEEP path: elementwise merges, softmax, warp, Kalman
// update, logits->argmax->codebook gather, GMFlow geometry (windows, sine position, convex x8).
#include "ops.h"

namespace keep {
namespace {

__device__ __forceinline__ float4 ld4_any(const void* p, int dt, size_t i) {
    return dt == F32 ? ld4(reinterpret_cast<const float*>(p), i) : ld4(reinterpret_cast<const __half*>(p), i);
}
__device__ __forceinline__ void st4_any(void* p, int dt, size_t i, float4 v) {
    if (dt == F32) st4(reinterpret_cast<float*>(p), i, v);
    else st4(reinterpret_cast<__half*>(p), i, v);
}
__device__ __forceinline__ float ld1_any(const void* p, int dt, size_t i) {
    return dt == F32 ? reinterpret_cast<const float*>(p)[i] : __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st1_any(void* p, int dt, size_t i, float v) {
    if (dt == F32) reinterpret_cast<float*>(p)[i] = v;
    else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) elementwise_kernel(const EwArgs a, size_t n4) {
    pdl_prologue();
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const size_t i = i4 * 4;
    const int c = (int)(i % a.c);
    const size_t nidx = i / ((size_t)a.hw * a.c);
    float4 va = ld4_any(a.A, a.a_dt, i);
    float x[4] = {va.x, va.y, va.z, va.w};
    if (a.sa) {
        const float4 s = *reinterpret_cast<const float4*>(a.sa + nidx * a.c + c);
        const float4 b = *reinterpret_cast<const float4*>(a.ba + nidx * a.c + c);
        x[0] = fmaf(x[0], s.x, b.x); x[1] = fmaf(x[1], s.y, b.y); x[2] = fmaf(x[2], s.z, b.z); x[3] = fmaf(x[3], s.w, b.w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = apply_act(x[j], a.act_a);
    if (a.B) {
        float4 vb = ld4_any(a.B, a.b_dt, i);
        float y[4] = {vb.x, vb.y, vb.z, vb.w};
        if (a.sb) {
            const float4 s = *reinterpret_cast<const float4*>(a.sb + nidx * a.c + c);
            const float4 b = *reinterpret_cast<const float4*>(a.bb + nidx * a.c + c);
            y[0] = fmaf(y[0], s.x, b.x); y[1] = fmaf(y[1], s.y, b.y); y[2] = fmaf(y[2], s.z, b.z); y[3] = fmaf(y[3], s.w, b.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] += apply_act(y[j], a.act_b);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = apply_act(x[j], a.act_o);
    st4_any(a.out, a.o_dt, i, make_float4(x[0], x[1], x[2], x[3]));
}

__global__ void __launch_bounds__(256) cft_combine_kernel(const void* dec, int dec_dt, const void* scale, const void* shift,
                                                          int ss_dt, float cond, void* out, int o_dt, size_t n4) {
    pdl_prologue_tiny();
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const size_t i = i4 * 4;
    const float4 d = ld4_any(dec, dec_dt, i), sc = ld4_any(scale, ss_dt, i), sh = ld4_any(shift, ss_dt, i);
    float4 o;
    o.x = d.x + cond * (d.x * sc.x + sh.x);
    o.y = d.y + cond * (d.y * sc.y + sh.y);
    o.z = d.z + cond * (d.z * sc.z + sh.z);
    o.w = d.w + cond * (d.w * sc.w + sh.w);
    st4_any(out, o_dt, i, o);
}

__global__ void __launch_bounds__(256) geglu_kernel(const float* __restrict__ in, float* __restrict__ out, size_t total4,
                                                    int inner) {
    pdl_prologue_tiny();
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const size_t i = i4 * 4;
    const size_t row = i / inner;
    const int col = (int)(i % inner);
    const float4 h = *reinterpret_cast<const float4*>(in + row * 2 * inner + col);
    const float4 g = *reinterpret_cast<const float4*>(in + row * 2 * inner + inner + col);
    float4 o;
    o.x = h.x * apply_act(g.x, ACT_GELU);
    o.y = h.y * apply_act(g.y, ACT_GELU);
    o.z = h.z * apply_act(g.z, ACT_GELU);
    o.w = h.w * apply_act(g.w, ACT_GELU);
    *reinterpret_cast<float4*>(out + i) = o;
}

// one warp per row, L <= 1024
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, long long rows, int L,
                                                           const int* __restrict__ region, int n_win, int Lq) {
    pdl_prologue_tiny();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* r = s + (size_t)row * L;
    float v[32];
    float mx = -INFINITY;
    const int* reg = nullptr;
    int myreg = 0;
    if (region) {
        const long long batch = row / Lq;
        reg = region + (size_t)(batch % n_win) * L;   // self-attention windows: Lq == L
        myreg = reg[row % Lq];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int k = lane + i * 32;
        float x = -INFINITY;
        if (k < L) {
            x = r[k];
            if (reg && reg[k] != myreg) x += -100.0f;
        }
        v[i] = x;
        mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int k = lane + i * 32;
        v[i] = k < L ? expf(v[i] - mx) : 0.0f;
        sum += v[i];
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int k = lane + i * 32;
        if (k < L) r[k] = v[i] * inv;
    }
}

// vectorised variant: L % 4 == 0, NV4 float4 per lane (L <= 128 * NV4), one warp per row
template <int NV4>
__global__ void __launch_bounds__(256) softmax_rows_v4_kernel(float* __restrict__ s, long long rows, int L,
                                                              const int* __restrict__ region, int n_win, int Lq) {
    pdl_prologue_tiny();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4* r = reinterpret_cast<float4*>(s + (size_t)row * L);
    const int L4 = L >> 2;
    const int4* reg = nullptr;
    int myreg = 0;
    if (region) {
        const long long batch = row / Lq;
        const int* rg = region + (size_t)(batch % n_win) * L;   // self-attention windows: Lq == L
        myreg = rg[row % Lq];
        reg = reinterpret_cast<const int4*>(rg);
    }
    float4 v[NV4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int q = lane + i * 32;
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (q < L4) {
            x = r[q];
            if (reg) {
                const int4 g = reg[q];
                if (g.x != myreg) x.x += -100.0f;
                if (g.y != myreg) x.y += -100.0f;
                if (g.z != myreg) x.z += -100.0f;
                if (g.w != myreg) x.w += -100.0f;
            }
        }
        v[i] = x;
        mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        if (lane + i * 32 < L4) {
            v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int q = lane + i * 32;
        if (q < L4) r[q] = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    }
}

// one block (256 threads) per row; V has two columns
__global__ void __launch_bounds__(256) softmax_expect2_kernel(const float* __restrict__ s, int L, int Lq,
                                                              const float* __restrict__ v, long long v_bstride,
                                                              const float* __restrict__ sub, float* __restrict__ out) {
    pdl_prologue();
    __shared__ float red[3][8];
    const long long row = blockIdx.x;
    const float* r = s + (size_t)row * L;
    const float* vv = v + (row / Lq) * v_bstride;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float mx = -INFINITY;
    for (int k = tid; k < L; k += 256) mx = fmaxf(mx, r[k]);
    mx = warp_max(mx);
    if (lane == 0) red[0][wid] = mx;
    __syncthreads();
    mx = red[0][0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[0][i]);
    __syncthreads();
    float se = 0.0f, sx = 0.0f, sy = 0.0f;
    for (int k = tid; k < L; k += 256) {
        const float e = expf(r[k] - mx);
        const float2 val = *reinterpret_cast<const float2*>(vv + 2 * (size_t)k);
        se += e;
        sx = fmaf(e, val.x, sx);
        sy = fmaf(e, val.y, sy);
    }
    se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) { red[0][wid] = se; red[1][wid] = sx; red[2][wid] = sy; }
    __syncthreads();
    if (tid == 0) {
        float a = 0, b = 0, c = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
        float ox = b / a, oy = c / a;
        if (sub) { ox -= sub[2 * (row % Lq)]; oy -= sub[2 * (row % Lq) + 1]; }
        out[2 * row] = ox;
        out[2 * row + 1] = oy;
    }
}

__global__ void __launch_bounds__(256) kalman_update_kernel(const float* __restrict__ z, const float* __restrict__ zp,
                                                            const float* __restrict__ gain, float* __restrict__ out,
                                                            size_t total, int c, int* status) {
    pdl_prologue_tiny();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float g = gain[i / c];
    const float v = (1.0f - g) * z[i] + g * zp[i];   // keep_arch.py:798
    out[i] = v;
    if (status && !isfinite(v)) atomicOr(status, KEEP_STATUS_BAD_LATENT);
}

// one warp per token; ties resolve to the lowest index
__global__ void __launch_bounds__(256) argmax_gather_kernel(const float* __restrict__ logits, int tokens, int ncodes,
                                                            const float* __restrict__ codebook, int cdim,
                                                            const int* __restrict__ forced, int* __restrict__ idx_out,
                                                            void* quant, int q_dt, int* status) {
    pdl_prologue_tiny();
    const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tok >= tokens) return;
    const float* r = logits + (size_t)tok * ncodes;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    bool bad = false;
    for (int k = lane; k < ncodes; k += 32) {
        const float x = r[k];
        bad = bad || !isfinite(x);
        if (x > best) { best = x; bi = k; }
    }
    if (status && bad) atomicOr(status, KEEP_STATUS_BAD_LOGITS);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;   // a row of NaNs compares false everywhere: stay inside the codebook (torch returns a NaN's index)
    if (lane == 0) idx_out[tok] = bi;
    const int use = forced ? forced[tok] : bi;
    for (int ch = lane; ch < cdim; ch += 32) st1_any(quant, q_dt, (size_t)tok * cdim + ch, codebook[(size_t)use * cdim + ch]);
}

// ---------------------------------------------------------------------------------------------
// VectorQuantizer.forward (vqgan_arch.py:37-76): nearest codebook entry per token,
//   d[t][j] = (||z_t||^2 + ||e_j||^2) - 2 z_t . e_j ,  idx[t] = argmin_j d[t][j]  (ties -> lowest j),  z_q[t] = e[idx[t]]
// One CTA = 16 tokens (8 warps x 2), staged once in shared memory.  The codebook streams through a 32-code x 128-dim
// shared tile (row pitch 132 floats: conflict-free 128-bit reads, one code per lane); a lane keeps the running dot
// products of its code with the warp's two tokens plus ||e||^2 in registers, so no shuffle is needed until the final
// warp-wide argmin (5 xor-shuffles of (distance, index) per token).  z is read once, the 1 MB codebook stays in L2.
// ---------------------------------------------------------------------------------------------
constexpr int VQ_TOK = 16, VQ_DC = 128, VQ_LD = VQ_DC + 4;
__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, int tokens, int cdim,
                                                         const float* __restrict__ codebook, int ncodes, int straight_through,
                                                         int* __restrict__ idx_out, float* __restrict__ zq,
                                                         float* __restrict__ dmin_out) {
    pdl_prologue();
    __shared__ __align__(16) float tile[32 * VQ_LD];
    extern __shared__ __align__(16) float vq_zs[];   // [VQ_TOK][cdim]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tok0 = blockIdx.x * VQ_TOK, c4n = cdim >> 2;
    for (int i = tid; i < VQ_TOK * c4n; i += 256) {
        const int t = i / c4n, tok = tok0 + t;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok < tokens) v = __ldg(reinterpret_cast<const float4*>(z) + (size_t)tok * c4n + (i - t * c4n));
        reinterpret_cast<float4*>(vq_zs)[i] = v;
    }
    __syncthreads();
    const float* za = vq_zs + (size_t)(warp * 2) * cdim;
    const float* zb = za + cdim;
    float zz0 = 0.f, zz1 = 0.f;
    for (int k = lane; k < cdim; k += 32) { zz0 = fmaf(za[k], za[k], zz0); zz1 = fmaf(zb[k], zb[k], zz1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        zz0 += __shfl_xor_sync(0xffffffffu, zz0, o);
        zz1 += __shfl_xor_sync(0xffffffffu, zz1, o);
    }
    float best0 = INFINITY, best1 = INFINITY;
    int bi0 = 0x7fffffff, bi1 = 0x7fffffff;
    for (int c0 = 0; c0 < ncodes; c0 += 32) {
        float dot0 = 0.f, dot1 = 0.f, ee = 0.f;
        for (int d0 = 0; d0 < cdim; d0 += VQ_DC) {
            __syncthreads();   // the previous tile has been consumed by every warp
            for (int i = tid; i < 32 * (VQ_DC / 4); i += 256) {
                const int r = i / (VQ_DC / 4), q = i - r * (VQ_DC / 4), code = c0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (code < ncodes) v = __ldg(reinterpret_cast<const float4*>(codebook + (size_t)code * cdim + d0) + q);
                *reinterpret_cast<float4*>(tile + r * VQ_LD + q * 4) = v;
            }
            __syncthreads();
            const float* row = tile + lane * VQ_LD;
#pragma unroll 8
            for (int k = 0; k < VQ_DC; k += 4) {
                const float4 e = *reinterpret_cast<const float4*>(row + k);
                const float4 a = *reinterpret_cast<const float4*>(za + d0 + k);
                const float4 b = *reinterpret_cast<const float4*>(zb + d0 + k);
                dot0 = fmaf(e.x, a.x, dot0); dot0 = fmaf(e.y, a.y, dot0); dot0 = fmaf(e.z, a.z, dot0); dot0 = fmaf(e.w, a.w, dot0);
                dot1 = fmaf(e.x, b.x, dot1); dot1 = fmaf(e.y, b.y, dot1); dot1 = fmaf(e.z, b.z, dot1); dot1 = fmaf(e.w, b.w, dot1);
                ee = fmaf(e.x, e.x, ee); ee = fmaf(e.y, e.y, ee); ee = fmaf(e.z, e.z, ee); ee = fmaf(e.w, e.w, ee);
            }
        }
        const int code = c0 + lane;
        if (code < ncodes) {   // codes arrive in increasing order: strict < keeps the lowest index on ties
            const float d0v = __fsub_rn(__fadd_rn(zz0, ee), __fmul_rn(2.f, dot0));
            const float d1v = __fsub_rn(__fadd_rn(zz1, ee), __fmul_rn(2.f, dot1));
            if (d0v < best0) { best0 = d0v; bi0 = code; }
            if (d1v < best1) { best1 = d1v; bi1 = code; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob0 = __shfl_xor_sync(0xffffffffu, best0, o), ob1 = __shfl_xor_sync(0xffffffffu, best1, o);
        const int oi0 = __shfl_xor_sync(0xffffffffu, bi0, o), oi1 = __shfl_xor_sync(0xffffffffu, bi1, o);
        if (ob0 < best0 || (ob0 == best0 && oi0 < bi0)) { best0 = ob0; bi0 = oi0; }
        if (ob1 < best1 || (ob1 == best1 && oi1 < bi1)) { best1 = ob1; bi1 = oi1; }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int tok = tok0 + warp * 2 + u;
        if (tok >= tokens) continue;
        const int bi = u ? bi1 : bi0;
        if (lane == 0) {
            idx_out[tok] = bi;
            if (dmin_out) dmin_out[tok] = u ? best1 : best0;
        }
        if (!zq || bi < 0 || bi >= ncodes) continue;   // all-NaN rows leave no valid index: nothing to gather
        const float* zt = u ? zb : za;
        for (int q = lane; q < c4n; q += 32) {
            float4 e = __ldg(reinterpret_cast<const float4*>(codebook + (size_t)bi * cdim) + q);
            if (straight_through) {   // forward value of z + (z_q - z).detach()   (vqgan_arch.py:61), same two roundings
                const float4 a = *reinterpret_cast<const float4*>(zt + q * 4);
                e.x = __fadd_rn(a.x, __fsub_rn(e.x, a.x)); e.y = __fadd_rn(a.y, __fsub_rn(e.y, a.y));
                e.z = __fadd_rn(a.z, __fsub_rn(e.z, a.z)); e.w = __fadd_rn(a.w, __fsub_rn(e.w, a.w));
            }
            reinterpret_cast<float4*>(zq + (size_t)tok * cdim)[q] = e;
        }
    }
}

__global__ void __launch_bounds__(256) sparse_causal_gather_kernel(const float* __restrict__ kv, float* __restrict__ out,
                                                                   int T, int L, int c4, size_t total4) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int cc = (int)(i % c4);
    size_t r = i / c4;
    const int j = (int)(r % (2 * L)); r /= (2 * L);
    const int f = (int)(r % T);
    const size_t b = r / T;
    const int srcf = j < L ? 0 : (f > 0 ? f - 1 : 0);
    const int tok = j < L ? j : j - L;
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(kv)[((b * T + srcf) * L + tok) * c4 + cc];
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, void* out, int o_dt, int c, int hw,
                                                           size_t total, int mode) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw
    if (i >= total) return;
    const size_t n = i / hw, p = i % hw;
    for (int ch = 0; ch < c; ++ch) {
        float v = x[(n * c + ch) * hw + p];
        if (mode == 1) {  // gmflow_arch.py:53-54 then gmflow/utils.py:55-63
            const float mean = ch == 0 ? 0.485f : (ch == 1 ? 0.456f : 0.406f);
            const float sd = ch == 0 ? 0.229f : (ch == 1 ? 0.224f : 0.225f);
            v = (v + 1.0f) / 2.0f * 255.0f;
            v = (v / 255.0f - mean) / sd;
        }
        st1_any(out, o_dt, i * c + ch, v);
    }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const void* x, int dt, void* out, int o_dt, int c, int hw,
                                                           size_t total, int* status) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw
    if (i >= total) return;
    const size_t n = i / hw, p = i % hw;
    bool bad = false;
    for (int ch = 0; ch < c; ++ch) {
        const float v = ld1_any(x, dt, i * c + ch);
        bad = bad || !isfinite(v);
        st1_any(out, o_dt, (n * c + ch) * hw + p, v);
    }
    if (status && bad) atomicOr(status, KEEP_STATUS_BAD_PIXELS);
}

// keep_processor.py:258-260: float32(crop_u8 / 255.) (the division is done in float64), BGR -> RGB, (v - 0.5) / 0.5
__global__ void __launch_bounds__(256) u8bgr_to_nchw_norm_kernel(const unsigned char* __restrict__ x, float* __restrict__ out, int hw,
                                                                 size_t total) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw
    if (i >= total) return;
    const size_t n = i / hw, p = i % hw;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {   // output channel ch (R, G, B) <- input byte 2 - ch
        float v = (float)((double)x[i * 3 + (2 - ch)] / 255.0);
        v = (v - 0.5f) / 0.5f;
        out[(n * 3 + ch) * hw + p] = v;
    }
}

// B/utils/img_util.py:38-94 (tensor2img, rgb2bgr=True, min_max=(-1, 1), uint8): clamp, (x - min) / (max - min), * 255, round half even
__global__ void __launch_bounds__(256) nhwc_to_u8bgr_kernel(const void* x, int dt, unsigned char* __restrict__ out, size_t total,
                                                            int* status) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw
    if (i >= total) return;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float v = ld1_any(x, dt, i * 3 + ch);
        if (status && !isfinite(v)) atomicOr(status, KEEP_STATUS_BAD_PIXELS);
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        v = (v - (-1.0f)) / (1.0f - (-1.0f));
        v = v * 255.0f;
        out[i * 3 + (2 - ch)] = (unsigned char)__float2int_rn(v);
    }
}

// arch_util.py:113-144 -> F.grid_sample(bilinear, zeros, align_corners=True)
__global__ void __launch_bounds__(256) flow_warp_kernel(const void* img, int dt, const float* __restrict__ flow, void* out,
                                                        int o_dt, int h, int w, int c, size_t total, int* status) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*h*w
    if (i >= total) return;
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const size_t n = i / ((size_t)w * h);
    const float vx = (float)x + flow[2 * i], vy = (float)y + flow[2 * i + 1];
    if (status && !(isfinite(vx) && isfinite(vy))) atomicOr(status, KEEP_STATUS_BAD_FLOW);
    const float wm = (float)max(w - 1, 1), hm = (float)max(h - 1, 1);
    const float gx = 2.0f * vx / wm - 1.0f, gy = 2.0f * vy / hm - 1.0f;
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(w - 1), iy = ((gy + 1.0f) / 2.0f) * (float)(h - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.0f) - ix, wy0 = (fy + 1.0f) - iy;
    const float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
    // guard the float->int conversion against huge flows
    const bool finite_ok = (ix > -4.0f) && (ix < (float)w + 4.0f) && (iy > -4.0f) && (iy < (float)h + 4.0f);
    const int x0 = finite_ok ? (int)fx : -8, y0 = finite_ok ? (int)fy : -8;
    const size_t base = n * (size_t)h * w;
    for (int ch = 0; ch < c; ++ch) {
        float acc = 0.0f;
        if (y0 >= 0 && y0 < h) {
            if (x0 >= 0 && x0 < w) acc += ld1_any(img, dt, (base + (size_t)y0 * w + x0) * c + ch) * w_nw;
            if (x0 + 1 >= 0 && x0 + 1 < w) acc += ld1_any(img, dt, (base + (size_t)y0 * w + x0 + 1) * c + ch) * w_ne;
        }
        if (y0 + 1 >= 0 && y0 + 1 < h) {
            if (x0 >= 0 && x0 < w) acc += ld1_any(img, dt, (base + (size_t)(y0 + 1) * w + x0) * c + ch) * w_sw;
            if (x0 + 1 >= 0 && x0 + 1 < w) acc += ld1_any(img, dt, (base + (size_t)(y0 + 1) * w + x0 + 1) * c + ch) * w_se;
        }
        st1_any(out, o_dt, i * c + ch, acc);
    }
}

// gmflow/position.py:26-46 on (h/splits, w/splits) windows, tiled over the map (gmflow/utils.py:66-86)
__global__ void __launch_bounds__(256) add_window_sine_pos_kernel(float* __restrict__ x, int h, int w, int c, int splits,
                                                                  size_t total) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % c);
    const size_t p = i / c;
    const int xx = (int)(p % w), yy = (int)((p / w) % h);
    const int wh = h / splits, ww = w / splits;
    const int half = c / 2;
    const bool is_y = ch < half;
    const int k = is_y ? ch : ch - half;
    const float pos = is_y ? (float)(yy % wh + 1) : (float)(xx % ww + 1);
    const float den = (is_y ? (float)wh : (float)ww) + 1e-6f;
    const float e = pos / den * 6.283185307179586f;
    const float dim_t = powf(10000.0f, (float)(2 * (k / 2)) / (float)half);
    const float arg = e / dim_t;
    x[i] += (k & 1) ? cosf(arg) : sinf(arg);
}

__global__ void __launch_bounds__(256) window_partition_kernel(const float* __restrict__ x, float* __restrict__ out, int h, int w,
                                                               int c4, int k, int sh, int sw, int ldx4, size_t total4,
                                                               int merge) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over windowed layout (n*k*k, wh*ww, c4)
    if (i >= total4) return;
    const int wh = h / k, ww = w / k;
    const int cc = (int)(i % c4);
    size_t r = i / c4;
    const int t = (int)(r % (wh * ww)); r /= (wh * ww);
    const int win = (int)(r % (k * k));
    const size_t n = r / (k * k);
    const int yy = t / ww, xx = t % ww;
    const int wy = win / k, wx = win % k;
    const int y = (wy * wh + yy + sh) % h, xcol = (wx * ww + xx + sw) % w;   // roll by -shift == read at +shift
    const size_t full = ((n * h + y) * w + xcol);
    if (!merge) reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(x)[full * ldx4 + cc];
    else reinterpret_cast<float4*>(out)[full * c4 + cc] = reinterpret_cast<const float4*>(x)[i];
}

__global__ void __launch_bounds__(256) convex_upsample8_kernel(const float* __restrict__ mask, const float* __restrict__ flow,
                                                               float* __restrict__ out, int h, int w, size_t total) {
    pdl_prologue();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n * 8h * 8w
    if (i >= total) return;
    const int W8 = 8 * w, H8 = 8 * h;
    const int X = (int)(i % W8), Y = (int)((i / W8) % H8);
    const size_t n = i / ((size_t)W8 * H8);
    const int x = X >> 3, j = X & 7, y = Y >> 3, ii = Y & 7;
    const float* m = mask + ((n * h + y) * w + x) * 576 + ii * 8 + j;
    float lg[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) { lg[k] = m[k * 64]; mx = fmaxf(mx, lg[k]); }
    float se = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float e = expf(lg[k] - mx);
        se += e;
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            const float2 f = *reinterpret_cast<const float2*>(flow + 2 * ((n * h + yy) * w + xx));
            ax = fmaf(e, 8.0f * f.x, ax);
            ay = fmaf(e, 8.0f * f.y, ay);
        }
    }
    out[2 * i] = ax / se;
    out[2 * i + 1] = ay / se;
}

// conv OIHW (or linear (O, I): taps = 1) -> [(tap * I + i)][o]: the layout of every kernel's weight operand; one-time, at keep_create
__global__ void __launch_bounds__(256) oihw_to_kc_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I, int taps,
                                                         size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int o = (int)(idx % O);
    const size_t r = idx / O;
    const int i = (int)(r % I), t = (int)(r / I);
    out[idx] = w[((size_t)o * I + i) * taps + t];
}

__global__ void __launch_bounds__(256) heads_to_tokens_kernel(const float4* __restrict__ x, float4* __restrict__ out, int heads, int rows,
                                                              int dh4, size_t total4) {
    pdl_prologue_tiny();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over the output: (row, head, d4)
    if (i >= total4) return;
    const int d = (int)(i % dh4);
    const size_t r = i / dh4;
    const int h = (int)(r % heads);
    const size_t row = r / heads;
    out[i] = x[((size_t)h * rows + row) * dh4 + d];
}

__global__ void __launch_bounds__(256) concat2_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb,
                                                      float* __restrict__ out, size_t total) {
    pdl_prologue_light();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ct = ca + cb;
    const size_t row = i / ct;
    const int col = (int)(i % ct);
    out[i] = col < ca ? a[row * ca + col] : (b ? b[row * cb + (col - ca)] : 0.0f);   // b == null: zero padding
}

static inline unsigned blocks_for(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }
}  // namespace

// ---------------------------------------------------------------------------------------------
void elementwise(const EwArgs& a, cudaStream_t s) {
    KEEP_CHECK(a.c % 4 == 0, "elementwise: c %% 4 != 0");
    const size_t n4 = (size_t)a.n * a.hw * a.c / 4;
    launch_k(elementwise_kernel, dim3(blocks_for(n4)), dim3(256), 0, s, a, n4);
    CUDA_CHECK(cudaGetLastError());
}

void cft_combine(const void* dec, int dec_dt, const void* scale, const void* shift, int ss_dt, float cond, void* out, int o_dt,
                 size_t numel, cudaStream_t s) {
    KEEP_CHECK(numel % 4 == 0, "cft_combine: numel %% 4 != 0");
    launch_k(cft_combine_kernel, dim3(blocks_for(numel / 4)), dim3(256), 0, s, dec, dec_dt, scale, shift, ss_dt, cond, out, o_dt, numel / 4);
    CUDA_CHECK(cudaGetLastError());
}

void geglu(const float* in, float* out, int rows, int inner, cudaStream_t s) {
    KEEP_CHECK(inner % 4 == 0, "geglu: inner %% 4 != 0");
    const size_t t4 = (size_t)rows * inner / 4;
    launch_k(geglu_kernel, dim3(blocks_for(t4)), dim3(256), 0, s, in, out, t4, inner);
    CUDA_CHECK(cudaGetLastError());
}

void softmax_rows(float* sc, long long rows, int L, const int* region, int n_win, int Lq, cudaStream_t s) {
    KEEP_CHECK(L <= 1024, "softmax_rows: L=%d > 1024", L);
    KEEP_CHECK(!region || Lq == L, "softmax_rows: region mask needs square windows");
    const int nw = n_win > 0 ? n_win : 1, lq = Lq > 0 ? Lq : 1;
    const bool v4 = L % 4 == 0 && ((reinterpret_cast<uintptr_t>(sc) | reinterpret_cast<uintptr_t>(region)) & 15) == 0;
    // few rows (the per-frame chain's 256-token attentions): 2 rows per block spreads them over the SMs
    const int wpb = rows <= 4096 ? 2 : 8;
    const dim3 grid(blocks_for((size_t)rows, wpb)), block(wpb * 32);
    if (v4 && L <= 128) launch_k(softmax_rows_v4_kernel<1>, grid, block, 0, s, sc, rows, L, region, nw, lq);
    else if (v4 && L <= 256) launch_k(softmax_rows_v4_kernel<2>, grid, block, 0, s, sc, rows, L, region, nw, lq);
    else if (v4 && L <= 512) launch_k(softmax_rows_v4_kernel<4>, grid, block, 0, s, sc, rows, L, region, nw, lq);
    else if (v4 && L <= 1024) launch_k(softmax_rows_v4_kernel<8>, grid, block, 0, s, sc, rows, L, region, nw, lq);
    else
    launch_k(softmax_rows_kernel, dim3(blocks_for((size_t)rows, 8)), dim3(256), 0, s, sc, rows, L, region, nw, lq);
    CUDA_CHECK(cudaGetLastError());
}

void softmax_expect2(const float* sc, long long rows, int L, int Lq, const float* v, long long v_bstride, const float* sub,
                     float* out, cudaStream_t s) {
    launch_k(softmax_expect2_kernel, dim3((unsigned)rows), dim3(256), 0, s, sc, L, Lq, v, v_bstride, sub, out);
    CUDA_CHECK(cudaGetLastError());
}

void kalman_update(const float* z, const float* zp, const float* gain, float* out, int pixels, int c, cudaStream_t s, int* status) {
    const size_t total = (size_t)pixels * c;
    launch_k(kalman_update_kernel, dim3(blocks_for(total)), dim3(256), 0, s, z, zp, gain, out, total, c, status);
    CUDA_CHECK(cudaGetLastError());
}

void argmax_gather(const float* logits, int tokens, int ncodes, const float* codebook, int cdim, const int* forced_idx,
                   int* idx_out, void* quant, int q_dt, cudaStream_t s, int* status) {
    launch_k(argmax_gather_kernel, dim3(blocks_for((size_t)tokens, 8)), dim3(256), 0, s, logits, tokens, ncodes, codebook, cdim, forced_idx,
                                                                      idx_out, quant, q_dt, status);
    CUDA_CHECK(cudaGetLastError());
}

void vq_nearest(const float* z, int tokens, int cdim, const float* codebook, int ncodes, int straight_through, int* idx_out,
                float* zq, float* dmin_out, cudaStream_t s) {
    KEEP_CHECK(tokens > 0 && ncodes > 0, "vq_nearest: empty input (tokens %d, codes %d)", tokens, ncodes);
    KEEP_CHECK(cdim % VQ_DC == 0 && cdim <= 384, "vq_nearest: embedding dim %d (need a multiple of %d, <= 384)", cdim, VQ_DC);
    launch_k(vq_nearest_kernel, dim3((unsigned)((tokens + VQ_TOK - 1) / VQ_TOK)), dim3(256), (size_t)VQ_TOK * cdim * sizeof(float), s,
             z, tokens, cdim, codebook, ncodes, straight_through, idx_out, zq, dmin_out);
    CUDA_CHECK(cudaGetLastError());
}

void sparse_causal_gather(const float* kv, float* out, int b, int T, int L, int c, cudaStream_t s) {
    KEEP_CHECK(c % 4 == 0, "sparse_causal_gather: c %% 4");
    const size_t t4 = (size_t)b * T * 2 * L * (c / 4);
    launch_k(sparse_causal_gather_kernel, dim3(blocks_for(t4)), dim3(256), 0, s, kv, out, T, L, c / 4, t4);
    CUDA_CHECK(cudaGetLastError());
}

void nchw_to_nhwc(const float* x, void* out, int o_dt, int n, int c, int h, int w, int mode, cudaStream_t s) {
    const size_t total = (size_t)n * h * w;
    launch_k(nchw_to_nhwc_kernel, dim3(blocks_for(total)), dim3(256), 0, s, x, out, o_dt, c, h * w, total, mode);
    CUDA_CHECK(cudaGetLastError());
}

void nhwc_to_nchw(const void* x, int dt, void* out, int out_dt, int n, int c, int h, int w, cudaStream_t s, int* status) {
    const size_t total = (size_t)n * h * w;
    launch_k(nhwc_to_nchw_kernel, dim3(blocks_for(total)), dim3(256), 0, s, x, dt, out, out_dt, c, h * w, total, status);
    CUDA_CHECK(cudaGetLastError());
}

void u8bgr_to_nchw_norm(const unsigned char* x, float* out, int n, int h, int w, cudaStream_t s) {
    const size_t total = (size_t)n * h * w;
    launch_k(u8bgr_to_nchw_norm_kernel, dim3(blocks_for(total)), dim3(256), 0, s, x, out, h * w, total);
    CUDA_CHECK(cudaGetLastError());
}

void nhwc_to_u8bgr(const void* x, int dt, unsigned char* out, int n, int h, int w, cudaStream_t s, int* status) {
    const size_t total = (size_t)n * h * w;
    launch_k(nhwc_to_u8bgr_kernel, dim3(blocks_for(total)), dim3(256), 0, s, x, dt, out, total, status);
    CUDA_CHECK(cudaGetLastError());
}

void flow_warp(const void* img, int dt, const float* flow, void* out, int o_dt, int n, int h, int w, int c, cudaStream_t s, int* status) {
    const size_t total = (size_t)n * h * w;
    launch_k(flow_warp_kernel, dim3(blocks_for(total)), dim3(256), 0, s, img, dt, flow, out, o_dt, h, w, c, total, status);
    CUDA_CHECK(cudaGetLastError());
}

void add_window_sine_pos(float* x, int n, int h, int w, int c, int splits, cudaStream_t s) {
    const size_t total = (size_t)n * h * w * c;
    launch_k(add_window_sine_pos_kernel, dim3(blocks_for(total)), dim3(256), 0, s, x, h, w, c, splits, total);
    CUDA_CHECK(cudaGetLastError());
}

void window_partition(const float* x, float* out, int n, int h, int w, int c, int k, int shift_h, int shift_w, int ldx,
                      cudaStream_t s) {
    KEEP_CHECK(c % 4 == 0 && ldx % 4 == 0 && h % k == 0 && w % k == 0, "window_partition: bad shape");
    const size_t t4 = (size_t)n * h * w * (c / 4);
    launch_k(window_partition_kernel, dim3(blocks_for(t4)), dim3(256), 0, s, x, out, h, w, c / 4, k, shift_h, shift_w, ldx / 4, t4, 0);
    CUDA_CHECK(cudaGetLastError());
}

void window_merge(const float* x, float* out, int n, int h, int w, int c, int k, int shift_h, int shift_w, cudaStream_t s) {
    KEEP_CHECK(c % 4 == 0 && h % k == 0 && w % k == 0, "window_merge: bad shape");
    const size_t t4 = (size_t)n * h * w * (c / 4);
    launch_k(window_partition_kernel, dim3(blocks_for(t4)), dim3(256), 0, s, x, out, h, w, c / 4, k, shift_h, shift_w, c / 4, t4, 1);
    CUDA_CHECK(cudaGetLastError());
}

void convex_upsample8(const float* mask, const float* flow, float* out, int n, int h, int w, cudaStream_t s) {
    const size_t total = (size_t)n * 64 * h * w;
    launch_k(convex_upsample8_kernel, dim3(blocks_for(total)), dim3(256), 0, s, mask, flow, out, h, w, total);
    CUDA_CHECK(cudaGetLastError());
}

void oihw_to_kc(const float* w_dev, float* out_dev, int O, int I, int taps, cudaStream_t s) {
    const size_t total = (size_t)O * I * taps;
    cudaLaunchConfig_t cfg;   // plain launch (no PDL attribute): runs once at engine creation
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks_for(total)); cfg.blockDim = dim3(256); cfg.stream = s;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, oihw_to_kc_kernel, w_dev, out_dev, O, I, taps, total));
}

void heads_to_tokens(const float* x, float* out, int heads, int rows, int dh, cudaStream_t s) {
    KEEP_CHECK(dh % 4 == 0, "heads_to_tokens: dh %% 4");
    const size_t total4 = (size_t)heads * rows * (dh / 4);
    launch_k(heads_to_tokens_kernel, dim3(blocks_for(total4)), dim3(256), 0, s, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out),
             heads, rows, dh / 4, total4);
    CUDA_CHECK(cudaGetLastError());
}

void concat2(const float* a, int ca, const float* b, int cb, float* out, long long rows, cudaStream_t s) {
    const size_t total = (size_t)rows * (ca + cb);
    launch_k(concat2_kernel, dim3(blocks_for(total)), dim3(256), 0, s, a, ca, b, cb, out, total);
    CUDA_CHECK(cudaGetLastError());
}

KEEP_STAMP_SETTER(stamp_set_misc)

}  // namespace keep
