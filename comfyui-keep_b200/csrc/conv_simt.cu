// keep_b200 — exact-fp32 implicit-GEMM convolution / linear on CUDA cores (NHWC).
//
// This is the precision-reference compute path of the engine ("fp32" mode) and the fallback for
// shapes the tcgen05 kernel does not take (Cin = 3 stems, Cin = 130 upsampler, Cout = 3 / 1 heads).
// Replaces every nn.Conv2d / nn.Linear the reference dispatches to cuDNN / cuBLAS on this path
// (vqgan_arch.py:161-181,260-286,311-335; keep_arch.py:445-455,766-772,78-87,391-393,929,936-938;
//  gmflow/backbone.py:11-14,50,64; gmflow/gmflow.py:46-48; gmflow/transformer.py:128-143,336-337).
//
// Fusions: nearest-x2 upsample and channel concat in the im2col gather, GroupNorm/InstanceNorm
// apply + activation in the A-operand prologue, bias + activation + residual in the epilogue,
// deterministic split-K for the small-M (16^2 / 32^2) layers of the serial per-frame chain.
#include "ops.h"
#include "tc.h"

namespace keep {

namespace {
constexpr int BM = 128, BN = 64, BK = 16, NTHREADS = 256;

__device__ __forceinline__ void load8(const void* src, int dt, size_t off, float* v) {
    if (dt == F32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
        float4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(src) + off);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            v[2 * j] = f.x; v[2 * j + 1] = f.y;
        }
    }
}
__device__ __forceinline__ float load1(const void* src, int dt, size_t off) {
    return dt == F32 ? reinterpret_cast<const float*>(src)[off] : __half2float(reinterpret_cast<const __half*>(src)[off]);
}
__device__ __forceinline__ void store1(void* dst, int dt, size_t off, float v) {
    if (dt == F32) reinterpret_cast<float*>(dst)[off] = v;
    else reinterpret_cast<__half*>(dst)[off] = __float2half_rn(v);
}

template <bool FAST>
__global__ void __launch_bounds__(NTHREADS) conv_igemm_simt_kernel(const ConvArgs a) {
    pdl_prologue();
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int cin = a.c0 + a.c1;
    const int HoWo = a.ho * a.wo;
    const int M = a.n * HoWo;
    const int K = a.kh * a.kw * cin;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nchunks = (K + BK - 1) / BK;
    const int per = (nchunks + a.splitk - 1) / a.splitk;
    const int ck0 = blockIdx.z * per;
    const int ck1 = min(nchunks, ck0 + per);

    // ---- loader coordinates: one output pixel per thread, 8 consecutive k per thread
    const int lp = tid & (BM - 1);
    const int lk = (tid >> 7) * 8;
    const int m = m0 + lp;
    const bool mvalid = m < M;
    int pn = 0, oy = 0, ox = 0;
    if (mvalid) {
        pn = m / HoWo;
        int r = m - pn * HoWo;
        oy = r / a.wo;
        ox = r - oy * a.wo;
    }
    const int iy0 = oy * a.stride - a.pad_t, ix0 = ox * a.stride - a.pad_l;
    const int Hl = a.h * a.up, Wl = a.w * a.up;
    const int bk = tid >> 4, bn4 = (tid & 15) * 4;  // B loader: row k, 4 consecutive n

    float av[8];
    float bv[4];

    auto load_chunk = [&](int ck) {
        const int k0 = ck * BK;
        // ---------------- A (im2col gather + prologue) ----------------
        if (FAST) {
            const int kk = k0 + lk;
            const int tap = kk / cin;
            const int ci = kk - tap * cin;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            const int iy = iy0 + ky, ix = ix0 + kx;
            const bool ok = mvalid && iy >= 0 && iy < Hl && ix >= 0 && ix < Wl;
            if (ok) {
                const size_t pix = ((size_t)pn * a.h + iy / a.up) * a.w + ix / a.up;
                if (ci < a.c0) load8(a.in0, a.in0_dt, pix * a.c0 + ci, av);
                else load8(a.in1, a.in1_dt, pix * a.c1 + (ci - a.c0), av);
                if (a.pre_scale) {
                    const float* sc = a.pre_scale + (size_t)pn * cin + ci;
                    const float* sh = a.pre_shift + (size_t)pn * cin + ci;
#pragma unroll
                    for (int j = 0; j < 8; ++j) av[j] = apply_act(fmaf(av[j], sc[j], sh[j]), a.pre_act);
                } else if (a.pre_act != ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) av[j] = apply_act(av[j], a.pre_act);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) av[j] = 0.0f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kk = k0 + lk + j;
                float v = 0.0f;
                if (mvalid && kk < K) {
                    const int tap = kk / cin;
                    const int ci = kk - tap * cin;
                    const int ky = tap / a.kw, kx = tap - ky * a.kw;
                    const int iy = iy0 + ky, ix = ix0 + kx;
                    if (iy >= 0 && iy < Hl && ix >= 0 && ix < Wl) {
                        const size_t pix = ((size_t)pn * a.h + iy / a.up) * a.w + ix / a.up;
                        v = ci < a.c0 ? load1(a.in0, a.in0_dt, pix * a.c0 + ci)
                                      : load1(a.in1, a.in1_dt, pix * a.c1 + (ci - a.c0));
                        if (a.pre_scale) v = fmaf(v, a.pre_scale[(size_t)pn * cin + ci], a.pre_shift[(size_t)pn * cin + ci]);
                        v = apply_act(v, a.pre_act);
                    }
                }
                av[j] = v;
            }
        }
        // ---------------- B (weights [K][cout]) ----------------
        const int kb = k0 + bk;
        if (FAST) {
            if (n0 + bn4 < a.cout) {  // kb < K always (K % 16 == 0), cout % 4 == 0
                float4 t = *reinterpret_cast<const float4*>(a.wt + (size_t)kb * a.cout + n0 + bn4);
                bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
            } else {
                bv[0] = bv[1] = bv[2] = bv[3] = 0.0f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nn = n0 + bn4 + j;
                bv[j] = (kb < K && nn < a.cout) ? a.wt[(size_t)kb * a.cout + nn] : 0.0f;
            }
        }
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const int ty = tid >> 4, tx = tid & 15;
    if (ck0 < ck1) load_chunk(ck0);
    for (int ck = ck0; ck < ck1; ++ck) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) As[lk + j][lp] = av[j];
        *reinterpret_cast<float4*>(&Bs[bk][bn4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
        if (ck + 1 < ck1) load_chunk(ck + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

    // ---------------- epilogue ----------------
    const int nb = n0 + tx * 4;
    if (a.splitk > 1) {
        float* part = a.partial + (size_t)blockIdx.z * M * a.cout;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int mm = m0 + ty * 8 + i;
            if (mm >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nb + j < a.cout) part[(size_t)mm * a.cout + nb + j] = acc[i][j];
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int mm = m0 + ty * 8 + i;
        if (mm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = nb + j;
            if (nn >= a.cout) continue;
            float v = acc[i][j] + (a.bias ? a.bias[nn] : 0.0f);
            v = apply_act(v, a.act);
            const size_t o = (size_t)mm * a.cout + nn;
            if (a.res) v += load1(a.res, a.res_dt, o);
            store1(a.out, a.out_dt, o, v);
        }
    }
}

// four consecutive outputs per thread (MN and cout are multiples of 4), all K-split loads of a thread in flight at once;
// partials are summed in split order -> deterministic.
// GN (gn_part != null; MN % 1024 == 0, cout in {128, 256, 512}): the block also emits GroupNorm(32) partial statistics of
// the final values it writes -- its 1024 consecutive elements are 1024 / cout whole pixels of one image, every group has
// exactly 8 of the block's threads; 32 threads add them up in a fixed order -> one slot per group, gn_part[img][32][slot][2]
template <bool GN>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splitk, long long MN, int cout,
                                                            const float* __restrict__ bias, int act, const void* res,
                                                            int res_dt, void* out, int out_dt, float* __restrict__ gn_part, int gn_P,
                                                            long long img_elems, const float* __restrict__ fin_gamma,
                                                            const float* __restrict__ fin_beta, float* __restrict__ fin_scale,
                                                            float* __restrict__ fin_shift, int* __restrict__ fin_tickets) {
    pdl_prologue_tiny();
    __shared__ float2 s_gn[GN ? 256 : 1];
    __shared__ int s_last;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (!GN && i >= MN) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 4 <= splitk; z += 4) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)z * MN + i));
        const float4 p1 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 1) * MN + i));
        const float4 p2 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 2) * MN + i));
        const float4 p3 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 3) * MN + i));
        v.x = (((v.x + p0.x) + p1.x) + p2.x) + p3.x; v.y = (((v.y + p0.y) + p1.y) + p2.y) + p3.y;
        v.z = (((v.z + p0.z) + p1.z) + p2.z) + p3.z; v.w = (((v.w + p0.w) + p1.w) + p2.w) + p3.w;
    }
    for (; z < splitk; ++z) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)z * MN + i));
        v.x += p0.x; v.y += p0.y; v.z += p0.z; v.w += p0.w;
    }
    if (bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias + (i % cout));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (act != ACT_NONE) { v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act); }
    if (res) {
        const float4 r = res_dt == F32 ? ld4(reinterpret_cast<const float*>(res), (size_t)i) : ld4(reinterpret_cast<const __half*>(res), (size_t)i);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (out_dt == F32) st4(reinterpret_cast<float*>(out), (size_t)i, v);
    else st4(reinterpret_cast<__half*>(out), (size_t)i, v);
    if (GN) {
        s_gn[threadIdx.x] = make_float2((v.x + v.y) + (v.z + v.w), fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w));
        __syncthreads();
        if (threadIdx.x < 32) {
            const int g = threadIdx.x, tpp = cout >> 2, tpg = cout >> 7;   // threads per pixel / per group (cpg / 4)
            float ss = 0.0f, qq = 0.0f;
            for (int p = 0; p < 256 / tpp; ++p)
                for (int k = 0; k < tpg; ++k) {
                    const float2 e = s_gn[p * tpp + g * tpg + k];
                    ss += e.x; qq += e.y;
                }
            const long long e0 = (long long)blockIdx.x * 1024;
            const long long img = e0 / img_elems;
            const long long slot = (e0 - img * img_elems) >> 10;
            *reinterpret_cast<float2*>(gn_part + ((size_t)(img * 32 + g) * gn_P + slot) * 2) = make_float2(ss, qq);
            if (fin_scale) __threadfence();
        }
        if (fin_scale) {
            // the image's last block to arrive turns the 32 x gn_P slots into the consumer's per-channel affine (the gn_P blocks
            // of an image are this kernel's only writers of its slots); 8 threads per group, double accumulation, fixed order
            const long long img = ((long long)blockIdx.x * 1024) / img_elems;
            __syncthreads();
            if (threadIdx.x == 0) {
                const int t = atomicAdd(fin_tickets + img, 1);
                s_last = (t == gn_P - 1);
                if (s_last) fin_tickets[img] = 0;
            }
            __syncthreads();
            if (!s_last) return;
            __threadfence();
            const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
            const float2* base = reinterpret_cast<const float2*>(gn_part) + (size_t)(img * 32 + g) * gn_P;
            double sm = 0.0, sq = 0.0;
            int k = l;
            for (; k + 24 < gn_P; k += 32) {   // four loads in flight per lane (a rolled loop paid one L2 round trip per slot); same order of additions
                const float2 e0 = __ldcg(base + k), e1 = __ldcg(base + k + 8), e2 = __ldcg(base + k + 16), e3 = __ldcg(base + k + 24);
                sm += (double)e0.x; sq += (double)e0.y;
                sm += (double)e1.x; sq += (double)e1.y;
                sm += (double)e2.x; sq += (double)e2.y;
                sm += (double)e3.x; sq += (double)e3.y;
            }
            for (; k < gn_P; k += 8) {
                const float2 e = __ldcg(base + k);
                sm += (double)e.x; sq += (double)e.y;
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) {
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            const int cpg = cout >> 5;
            const double cnt = (double)(img_elems / cout) * cpg;
            const double mean = sm / cnt;
            double var = sq / cnt - mean * mean;
            if (var < 0) var = 0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-6));
            for (int i = l; i < cpg; i += 8) {
                const int ch = g * cpg + i;
                const float sc = fin_gamma[ch] * rstd;
                fin_scale[(size_t)img * cout + ch] = sc;
                fin_shift[(size_t)img * cout + ch] = fin_beta[ch] - (float)mean * sc;
            }
        }
    }
}
// Split-K reduce whose rows feed a LayerNorm (transformer linears): the block's 1024 consecutive elements are 1024 / cout whole
// rows, cout / 128 warps per row; two-pass statistics on the values held in registers (as layernorm_v4_kernel)
__global__ void __launch_bounds__(256) splitk_reduce_ln_kernel(const float* __restrict__ part, int splitk, long long MN, int cout,
                                                               const float* __restrict__ bias, int act, const void* res, int res_dt,
                                                               float* __restrict__ out, const float* __restrict__ ln_g,
                                                               const float* __restrict__ ln_b, float eps, float* __restrict__ ln_out,
                                                               const float* __restrict__ add2, int add2_rows, float* __restrict__ ln_out2) {
    pdl_prologue_tiny();
    __shared__ float s_red[2][8];
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 4 <= splitk; z += 4) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)z * MN + i));
        const float4 p1 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 1) * MN + i));
        const float4 p2 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 2) * MN + i));
        const float4 p3 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)(z + 3) * MN + i));
        v.x = (((v.x + p0.x) + p1.x) + p2.x) + p3.x; v.y = (((v.y + p0.y) + p1.y) + p2.y) + p3.y;
        v.z = (((v.z + p0.z) + p1.z) + p2.z) + p3.z; v.w = (((v.w + p0.w) + p1.w) + p2.w) + p3.w;
    }
    for (; z < splitk; ++z) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + (size_t)z * MN + i));
        v.x += p0.x; v.y += p0.y; v.z += p0.z; v.w += p0.w;
    }
    const int col = (int)(i % cout);
    if (bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias + col);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (act != ACT_NONE) { v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act); }
    if (res) {
        const float4 r = res_dt == F32 ? ld4(reinterpret_cast<const float*>(res), (size_t)i) : ld4(reinterpret_cast<const __half*>(res), (size_t)i);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    st4(out, (size_t)i, v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpr = cout >> 7, w0 = (warp / wpr) * wpr;
    float sm = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    if (lane == 0) s_red[0][warp] = sm;
    __syncthreads();
    float tot = 0.0f;
    for (int k = 0; k < wpr; ++k) tot += s_red[0][w0 + k];
    const float mean = tot / (float)cout;
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    float qq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    if (lane == 0) s_red[1][warp] = qq;
    __syncthreads();
    tot = 0.0f;
    for (int k = 0; k < wpr; ++k) tot += s_red[1][w0 + k];
    const float rstd = rsqrtf(tot / (float)cout + eps);
    const float4 gg = *reinterpret_cast<const float4*>(ln_g + col), bb = *reinterpret_cast<const float4*>(ln_b + col);
    float4 y;
    y.x = dx * rstd * gg.x + bb.x; y.y = dy * rstd * gg.y + bb.y; y.z = dz * rstd * gg.z + bb.z; y.w = dw * rstd * gg.w + bb.w;
    st4(ln_out, (size_t)i, y);
    if (ln_out2) {
        const long long row = i / cout;
        const float4 a2 = *reinterpret_cast<const float4*>(add2 + (size_t)(row % add2_rows) * cout + col);
        st4(ln_out2, (size_t)i, make_float4(y.x + a2.x, y.y + a2.y, y.z + a2.z, y.w + a2.w));
    }
}
__global__ void __launch_bounds__(256) splitk_reduce1_kernel(const float* __restrict__ part, int splitk, long long MN, int cout,
                                                             const float* __restrict__ bias, int act, const void* res,
                                                             int res_dt, void* out, int out_dt) {
    pdl_prologue_tiny();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float v = 0.0f;
    for (int z = 0; z < splitk; ++z) v += part[(size_t)z * MN + i];
    if (bias) v += bias[i % cout];
    v = apply_act(v, act);
    if (res) v += load1(res, res_dt, i);
    store1(out, out_dt, i, v);
}
}  // namespace

int conv_pick_splitk(const ConvArgs& a) {
    const long long M = (long long)a.n * a.ho * a.wo;
    const int K = a.kh * a.kw * (a.c0 + a.c1);
    const long long ctas = (long long)cdiv(M, BM) * cdiv(a.cout, BN);
    const int nchunks = cdiv(K, BK);
    if (ctas >= 120 || nchunks < 16) return 1;
    long long s = (296 + ctas - 1) / ctas;
    if (s > nchunks / 4) s = nchunks / 4;
    if (s > 32) s = 32;
    return s < 1 ? 1 : (int)s;
}

void splitk_reduce(const float* partial, int splitk, long long MN, int cout, const float* bias, int act, const void* res,
                   int res_dt, void* out, int out_dt, cudaStream_t s, float* gn_part, int gn_P, long long hw, const float* fin_gamma,
                   const float* fin_beta, float* fin_scale, float* fin_shift, int* fin_tickets) {
    KEEP_CHECK(!fin_scale || (gn_part && fin_shift && fin_gamma && fin_beta && fin_tickets && gn_P <= kGnReduceFinalMaxP),
               "splitk_reduce: finalize-in-reduce needs statistics slots (<= %d per group), gamma / beta and tickets", kGnReduceFinalMaxP);
    if (gn_part) {
        const long long img_elems = hw * cout;
        KEEP_CHECK(MN % 1024 == 0 && img_elems % 1024 == 0 && (cout == 128 || cout == 256 || cout == 512) && gn_P == img_elems / 1024,
                   "splitk_reduce: layer cannot emit GroupNorm statistics (MN %lld, cout %d)", MN, cout);
        launch_k(splitk_reduce_kernel<true>, dim3((unsigned)(MN / 1024)), dim3(256), 0, s, partial, splitk, MN, cout, bias, act, res, res_dt, out,
                 out_dt, gn_part, gn_P, img_elems, fin_gamma, fin_beta, fin_scale, fin_shift, fin_tickets);
    } else if (MN % 4 != 0 || cout % 4 != 0)
        launch_k(splitk_reduce1_kernel, dim3(cdiv(MN, 256)), dim3(256), 0, s, partial, splitk, MN, cout, bias, act, res, res_dt, out, out_dt);
    else
        launch_k(splitk_reduce_kernel<false>, dim3(cdiv(MN, 1024)), dim3(256), 0, s, partial, splitk, MN, cout, bias, act, res, res_dt, out, out_dt,
                 (float*)nullptr, 0, 0LL, (const float*)nullptr, (const float*)nullptr, (float*)nullptr, (float*)nullptr, (int*)nullptr);
    CUDA_CHECK(cudaGetLastError());
}

bool splitk_reduce_ln_eligible(long long MN, int cout) {
    return MN > 0 && MN % 1024 == 0 && (cout == 128 || cout == 256 || cout == 512 || cout == 1024);
}

void splitk_reduce_ln(const float* partial, int splitk, long long MN, int cout, const float* bias, int act, const void* res, int res_dt,
                      float* out, const float* ln_g, const float* ln_b, float eps, float* ln_out, const float* add2, int add2_rows,
                      float* ln_out2, cudaStream_t s) {
    KEEP_CHECK(splitk_reduce_ln_eligible(MN, cout) && ln_g && ln_b && ln_out && (!ln_out2 || (add2 && add2_rows > 0)),
               "splitk_reduce_ln: unsupported shape (MN %lld, cout %d) or null LayerNorm argument", MN, cout);
    launch_k(splitk_reduce_ln_kernel, dim3((unsigned)(MN / 1024)), dim3(256), 0, s, partial, splitk, MN, cout, bias, act, res, res_dt, out, ln_g,
             ln_b, eps, ln_out, add2, add2_rows > 0 ? add2_rows : 1, ln_out2);
    CUDA_CHECK(cudaGetLastError());
}

void conv2d_simt(const ConvArgs& a, cudaStream_t s) {
    const int cin = a.c0 + a.c1;
    const long long M = (long long)a.n * a.ho * a.wo;
    KEEP_CHECK(M > 0 && a.cout > 0 && cin > 0, "conv2d_simt: empty problem");
    KEEP_CHECK(a.up == 1 || a.up == 2, "conv2d_simt: up must be 1 or 2");
    KEEP_CHECK(a.splitk == 1 || a.partial != nullptr, "conv2d_simt: split-K needs a partial buffer");
    const bool fast = (a.c0 % 16 == 0) && (a.c1 % 16 == 0) && (a.cout % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.in0) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.in1) & 15) == 0);
    dim3 grid(cdiv(M, BM), cdiv(a.cout, BN), a.splitk);
    if (fast) launch_k(conv_igemm_simt_kernel<true>, dim3(grid), dim3(NTHREADS), 0, s, a);
    else launch_k(conv_igemm_simt_kernel<false>, dim3(grid), dim3(NTHREADS), 0, s, a);
    CUDA_CHECK(cudaGetLastError());
    if (a.splitk > 1) {
        const long long MN = M * a.cout;
        splitk_reduce(a.partial, a.splitk, MN, a.cout, a.bias, a.act, a.res, a.res_dt, a.out, a.out_dt, s);   // (no statistics on this path)
    }
}

KEEP_STAMP_SETTER(stamp_set_conv_simt)

}  // namespace keep
