// keep_b200 — the two memory-bound ends of the VQGAN / GMFlow stacks, which have no tensor-core shape:
//   * stems:  Cin = 3  -> 64  (3x3 s1 p1: vqgan_arch.py:260-261 ; 7x7 s2 p3: gmflow/backbone.py:50)
//   * heads:  Cin = 64 -> 3   (3x3 s1 p1 after GroupNorm: vqgan_arch.py:333-335)
// Both are bandwidth problems (K = 27 / 147, or N = 3): one pass over the big 512x512 tensor, fp32 math.
#include "ops.h"

namespace keep {
namespace {

__device__ __forceinline__ float ld1(const void* p, int dt, size_t i) {
    return dt == F32 ? reinterpret_cast<const float*>(p)[i] : __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// ---- stem: thread = (output pixel, group of 16 output channels); weights [KS*KS*3][64] in shared memory
template <int KS, int STRIDE>
__global__ void __launch_bounds__(256) conv_cin3_kernel(const void* __restrict__ in, int in_dt, int n, int h, int w, int pad,
                                                        const float* __restrict__ wt, const float* __restrict__ bias, int ho, int wo,
                                                        void* __restrict__ out, int out_dt) {
    pdl_prologue();
    constexpr int K = KS * KS * 3;
    __shared__ __align__(16) float sw[K * 64];
    // weights in shared memory as [k][q][grp][4]: the four channel groups' q-th float4 are 64 contiguous bytes, so the
    // LDS.128 of a quarter-warp (2 pixel groups x 4 channel groups) hits 16 distinct banks (the plain [k][64] layout put groups
    // 0 / 2 and 1 / 3 on the same banks: a 2-way conflict on every weight load -- ncu: 3.7M conflicts per frame)
    for (int i = threadIdx.x; i < K * 64; i += 256) {
        const int k = i >> 6, ch = i & 63;
        sw[k * 64 + ((ch & 15) >> 2) * 16 + (ch >> 4) * 4 + (ch & 3)] = wt[i];
    }
    __syncthreads();
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long pix = gid >> 2;
    const int grp = (int)(gid & 3);
    if (pix >= (long long)n * ho * wo) return;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), img = (int)(pix / ((long long)wo * ho));
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = bias ? bias[grp * 16 + j] : 0.0f;
    const int iy0 = oy * STRIDE - pad, ix0 = ox * STRIDE - pad;
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
        const int iy = iy0 + ky;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
            const int ix = ix0 + kx;
            if (ix < 0 || ix >= w) continue;
            const size_t base = (((size_t)img * h + iy) * w + ix) * 3;
            const float x0 = ld1(in, in_dt, base), x1 = ld1(in, in_dt, base + 1), x2 = ld1(in, in_dt, base + 2);
            const float4* w0 = reinterpret_cast<const float4*>(sw + ((ky * KS + kx) * 3 + 0) * 64 + grp * 4);
            const float4* w1 = reinterpret_cast<const float4*>(sw + ((ky * KS + kx) * 3 + 1) * 64 + grp * 4);
            const float4* w2 = reinterpret_cast<const float4*>(sw + ((ky * KS + kx) * 3 + 2) * 64 + grp * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a = w0[q * 4], b = w1[q * 4], c = w2[q * 4];
                acc[4 * q + 0] = fmaf(x0, a.x, fmaf(x1, b.x, fmaf(x2, c.x, acc[4 * q + 0])));
                acc[4 * q + 1] = fmaf(x0, a.y, fmaf(x1, b.y, fmaf(x2, c.y, acc[4 * q + 1])));
                acc[4 * q + 2] = fmaf(x0, a.z, fmaf(x1, b.z, fmaf(x2, c.z, acc[4 * q + 2])));
                acc[4 * q + 3] = fmaf(x0, a.w, fmaf(x1, b.w, fmaf(x2, c.w, acc[4 * q + 3])));
            }
        }
    }
    const size_t o = (size_t)pix * 64 + grp * 16;
    if (out_dt == F32) {
        float4* po = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o);
#pragma unroll
        for (int q = 0; q < 4; ++q) po[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st4(reinterpret_cast<__half*>(out), o + 4 * q, make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
    }
}

// ---- stem, register-tiled: thread = (4 consecutive output pixels of a row, group of 16 output channels).  Every weight
// float4 fetched from shared memory feeds 4 pixels (16 FMA per LDS.128: FMA-bound instead of LDS-bound), and the input row
// segment the 4 pixels share ((4-1)*STRIDE + KS columns x 3 channels, contiguous in NHWC) is loaded once per filter row.
template <int KS, int STRIDE>
__global__ void __launch_bounds__(256) conv_cin3_px4_kernel(const void* __restrict__ in, int in_dt, int n, int h, int w, int pad,
                                                            const float* __restrict__ wt, const float* __restrict__ bias, int ho, int wo,
                                                            void* __restrict__ out, int out_dt) {
    pdl_prologue();
    constexpr int K = KS * KS * 3, PX = 4, SPAN = (PX - 1) * STRIDE + KS;
    __shared__ __align__(16) float sw[K * 64];
    // weights in shared memory as [k][q][grp][4]: the four channel groups' q-th float4 are 64 contiguous bytes, so the
    // LDS.128 of a quarter-warp (2 pixel groups x 4 channel groups) hits 16 distinct banks (the plain [k][64] layout put groups
    // 0 / 2 and 1 / 3 on the same banks: a 2-way conflict on every weight load -- ncu: 3.7M conflicts per frame)
    for (int i = threadIdx.x; i < K * 64; i += 256) {
        const int k = i >> 6, ch = i & 63;
        sw[k * 64 + ((ch & 15) >> 2) * 16 + (ch >> 4) * 4 + (ch & 3)] = wt[i];
    }
    __syncthreads();
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long pg = gid >> 2;                      // group of 4 output pixels
    const int grp = (int)(gid & 3);
    const int wg = wo / PX;
    if (pg >= (long long)n * ho * wg) return;
    const int ox0 = (int)(pg % wg) * PX, oy = (int)((pg / wg) % ho), img = (int)(pg / ((long long)wg * ho));
    float acc[PX][16];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[p][j] = bias ? bias[grp * 16 + j] : 0.0f;
    const int ix0 = ox0 * STRIDE - pad;
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
        const int iy = oy * STRIDE - pad + ky;
        if (iy < 0 || iy >= h) continue;
        float xin[SPAN * 3];
        const size_t rowbase = ((size_t)img * h + iy) * w;
#pragma unroll
        for (int j = 0; j < SPAN; ++j) {
            const int ix = ix0 + j;
            const bool ok = ix >= 0 && ix < w;
#pragma unroll
            for (int c = 0; c < 3; ++c) xin[j * 3 + c] = ok ? ld1(in, in_dt, (rowbase + ix) * 3 + c) : 0.0f;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * KS + kx) * 3 + c) * 64 + grp * 4);
                const float4 w0 = wp[0], w1 = wp[4], w2 = wp[8], w3 = wp[12];
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const float x = xin[(p * STRIDE + kx) * 3 + c];
                    acc[p][0] = fmaf(x, w0.x, acc[p][0]); acc[p][1] = fmaf(x, w0.y, acc[p][1]);
                    acc[p][2] = fmaf(x, w0.z, acc[p][2]); acc[p][3] = fmaf(x, w0.w, acc[p][3]);
                    acc[p][4] = fmaf(x, w1.x, acc[p][4]); acc[p][5] = fmaf(x, w1.y, acc[p][5]);
                    acc[p][6] = fmaf(x, w1.z, acc[p][6]); acc[p][7] = fmaf(x, w1.w, acc[p][7]);
                    acc[p][8] = fmaf(x, w2.x, acc[p][8]); acc[p][9] = fmaf(x, w2.y, acc[p][9]);
                    acc[p][10] = fmaf(x, w2.z, acc[p][10]); acc[p][11] = fmaf(x, w2.w, acc[p][11]);
                    acc[p][12] = fmaf(x, w3.x, acc[p][12]); acc[p][13] = fmaf(x, w3.y, acc[p][13]);
                    acc[p][14] = fmaf(x, w3.z, acc[p][14]); acc[p][15] = fmaf(x, w3.w, acc[p][15]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const size_t o = ((((size_t)img * ho + oy) * wo) + ox0 + p) * 64 + grp * 16;
        if (out_dt == F32) {
            float4* po = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) po[q] = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st4(reinterpret_cast<__half*>(out), o + 4 * q, make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]));
        }
    }
}

// ---- head: 3x3 s1 p1, Cin % 16 == 0, Cout <= 4.  Block = 32x32 output pixels, thread = 4 pixels of one column (rows ly, ly + 8,
// ly + 16, ly + 24): every weight float4 fetched from shared memory (warp-uniform broadcast) feeds 4 pixels -- the first version
// (one pixel per thread: 5 LDS.128 per 16 FMA) was bound by the shared-memory pipe (ncu: mio_throttle 5.2 stalls per issue,
// 81 us per 512^2 frame).  The input halo is staged 8 channels at a time with the GroupNorm affine applied; pixel pitch 12
// floats: the float4 reads of 8 consecutive pixels hit 8 distinct bank quads.  Weights [9][cin][4] in shared memory.
constexpr int HT_W = 32, HT_H = 32, HROWS = 8, HPX = HT_H / HROWS, HC = 8, HPITCH = 12;
__global__ void __launch_bounds__(256, 2) conv_cout4_kernel(const void* __restrict__ in, int in_dt, int n, int h, int w, int cin,
                                                         const float* __restrict__ pre_scale, const float* __restrict__ pre_shift,
                                                         const float* __restrict__ wt /*[9*cin][cout]*/, const float* __restrict__ bias,
                                                         int cout, float* __restrict__ out) {
    pdl_prologue();
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                                   // [9][cin][4]
    float* sx = sm + 9 * cin * 4;                     // [(HT_H+2)*(HT_W+2)][HPITCH]
    for (int i = threadIdx.x; i < 9 * cin * 4; i += 256) {
        const int c = i & 3, k = i >> 2;
        sw[i] = c < cout ? wt[(size_t)k * cout + c] : 0.0f;
    }
    const int tiles_x = (w + HT_W - 1) / HT_W, tiles_y = (h + HT_H - 1) / HT_H;
    const int img = blockIdx.x / (tiles_x * tiles_y);
    const int t = blockIdx.x - img * tiles_x * tiles_y;
    const int ty0 = (t / tiles_x) * HT_H, tx0 = (t % tiles_x) * HT_W;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    float acc[HPX][4];
#pragma unroll
    for (int j = 0; j < HPX; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
    constexpr int HALO = (HT_H + 2) * (HT_W + 2);
    for (int c0 = 0; c0 < cin; c0 += HC) {
        __syncthreads();
        // every global load of the stage in flight before the first is used (a rolled loop paid one L2 round trip per iteration:
        // ~7 of the 8.8 us a stage took)
        constexpr int NU = (HALO * (HC / 4) + 255) / 256;
        float4 v[NU];
        unsigned inside = 0;                // bit i: unit i lies inside the image (zero padding is applied AFTER the normalisation)
        const int q = threadIdx.x & 1;      // (256 is even: a thread's units all have the same channel quad)
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const int u = threadIdx.x + i * 256, p = u >> 1;
            const int hy = p / (HT_W + 2), hx = p - hy * (HT_W + 2);
            const int iy = ty0 + hy - 1, ix = tx0 + hx - 1;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < HALO * (HC / 4) && iy >= 0 && iy < h && ix >= 0 && ix < w) {
                const size_t off = (((size_t)img * h + iy) * w + ix) * cin + c0 + q * 4;
                v[i] = in_dt == F32 ? ld4(reinterpret_cast<const float*>(in), off) : ld4(reinterpret_cast<const __half*>(in), off);
                inside |= 1u << i;
            }
        }
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pre_scale) {
            sc = *reinterpret_cast<const float4*>(pre_scale + (size_t)img * cin + c0 + q * 4);
            sh = *reinterpret_cast<const float4*>(pre_shift + (size_t)img * cin + c0 + q * 4);
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const int u = threadIdx.x + i * 256;
            if (u >= HALO * (HC / 4)) continue;
            float4 o = v[i];
            if (pre_scale && ((inside >> i) & 1u)) {
                o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y); o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
            }
            *reinterpret_cast<float4*>(sx + (u >> 1) * HPITCH + q * 4) = o;
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float4* wp = reinterpret_cast<const float4*>(sw + (tap * cin + c0) * 4);
#pragma unroll
            for (int q = 0; q < HC / 4; ++q) {
                const float4 w0 = wp[q * 4 + 0], w1 = wp[q * 4 + 1], w2 = wp[q * 4 + 2], w3 = wp[q * 4 + 3];
#pragma unroll
                for (int j = 0; j < HPX; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(sx + ((ly + j * HROWS + tap / 3) * (HT_W + 2) + lx + tap % 3) * HPITCH + q * 4);
                    acc[j][0] = fmaf(x.x, w0.x, fmaf(x.y, w1.x, fmaf(x.z, w2.x, fmaf(x.w, w3.x, acc[j][0]))));
                    acc[j][1] = fmaf(x.x, w0.y, fmaf(x.y, w1.y, fmaf(x.z, w2.y, fmaf(x.w, w3.y, acc[j][1]))));
                    acc[j][2] = fmaf(x.x, w0.z, fmaf(x.y, w1.z, fmaf(x.z, w2.z, fmaf(x.w, w3.z, acc[j][2]))));
                    acc[j][3] = fmaf(x.x, w0.w, fmaf(x.y, w1.w, fmaf(x.z, w2.w, fmaf(x.w, w3.w, acc[j][3]))));
                }
            }
        }
    }
    const int ox = tx0 + lx;
#pragma unroll
    for (int j = 0; j < HPX; ++j) {
        const int oy = ty0 + ly + j * HROWS;
        if (oy < h && ox < w) {
            const size_t o = (((size_t)img * h + oy) * w + ox) * cout;
#pragma unroll
            for (int c = 0; c < 4; ++c)      // (compile-time indices: a run-time `c < cout` loop index would push acc[][] into local memory)
                if (c < cout) out[o + c] = acc[j][c] + (bias ? bias[c] : 0.0f);
        }
    }
}
}  // namespace

bool conv_small_eligible(const ConvArgs& a) {
    const int cin = a.c0 + a.c1;
    if (a.c1 != 0 || a.up != 1 || a.res != nullptr || a.act != ACT_NONE) return false;
    if (cin == 3 && a.cout == 64 && a.pre_scale == nullptr && a.pre_act == ACT_NONE) {
        if (a.kh == 3 && a.kw == 3 && a.stride == 1 && a.pad_t == 1 && a.pad_l == 1) return true;
        if (a.kh == 7 && a.kw == 7 && a.stride == 2 && a.pad_t == 3 && a.pad_l == 3) return true;
        return false;
    }
    if (a.cout <= 4 && cin % 16 == 0 && cin <= 128 && a.kh == 3 && a.kw == 3 && a.stride == 1 && a.pad_t == 1 && a.pad_l == 1 &&
        a.pre_act == ACT_NONE && a.out_dt == F32 && a.ho == a.h && a.wo == a.w)
        return true;
    return false;
}

void conv_small_configure_device() {
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured))
        CUDA_CHECK(cudaFuncSetAttribute(conv_cout4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
}

void conv2d_small(const ConvArgs& a, cudaStream_t s) {
    KEEP_CHECK(conv_small_eligible(a), "conv2d_small: not eligible");
    const int cin = a.c0 + a.c1;
    if (cin == 3) {
        if (a.wo % 4 == 0) {   // register-tiled variant: 4 output pixels per thread
            const long long threads4 = (long long)a.n * a.ho * (a.wo / 4) * 4;
            const unsigned grid4 = (unsigned)((threads4 + 255) / 256);
            if (a.kh == 3) launch_k(conv_cin3_px4_kernel<3, 1>, dim3(grid4), dim3(256), 0, s, a.in0, a.in0_dt, a.n, a.h, a.w, 1, a.wt, a.bias, a.ho, a.wo, a.out, a.out_dt);
            else launch_k(conv_cin3_px4_kernel<7, 2>, dim3(grid4), dim3(256), 0, s, a.in0, a.in0_dt, a.n, a.h, a.w, 3, a.wt, a.bias, a.ho, a.wo, a.out, a.out_dt);
            CUDA_CHECK(cudaGetLastError());
            return;
        }
        const long long threads = (long long)a.n * a.ho * a.wo * 4;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        if (a.kh == 3) launch_k(conv_cin3_kernel<3, 1>, dim3(grid), dim3(256), 0, s, a.in0, a.in0_dt, a.n, a.h, a.w, 1, a.wt, a.bias, a.ho, a.wo, a.out, a.out_dt);
        else launch_k(conv_cin3_kernel<7, 2>, dim3(grid), dim3(256), 0, s, a.in0, a.in0_dt, a.n, a.h, a.w, 3, a.wt, a.bias, a.ho, a.wo, a.out, a.out_dt);
    } else {
        const int tiles = ((a.w + HT_W - 1) / HT_W) * ((a.h + HT_H - 1) / HT_H);
        const size_t smem = (size_t)(9 * cin * 4 + (HT_H + 2) * (HT_W + 2) * HPITCH) * sizeof(float);
        conv_small_configure_device();
        launch_k(conv_cout4_kernel, dim3(a.n * tiles), dim3(256), smem, s, a.in0, a.in0_dt, a.n, a.h, a.w, cin, a.pre_scale, a.pre_shift, a.wt, a.bias,
                                                         a.cout, (float*)a.out);
    }
    CUDA_CHECK(cudaGetLastError());
}

KEEP_STAMP_SETTER(stamp_set_conv_small)

}  // namespace keep
