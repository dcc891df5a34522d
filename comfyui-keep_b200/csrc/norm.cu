// keep_b200 — GroupNorm / InstanceNorm statistics -> per-(n, channel) affine, and LayerNorm.
//
// GroupNorm(32, eps 1e-6) (vqgan_arch.py:16-17) and GMFlow's affine-free InstanceNorm2d (eps 1e-5,
// gmflow/backbone.py:17-36) are both "statistics over H*W*cpg, then y = x*scale[n][c] + shift[n][c]".
// Statistics are reduced here (double accumulation, two deterministic stages); the *apply* (and the
// swish / ReLU that follows) is fused into the consuming convolution's A-operand prologue or into
// the residual-merge elementwise kernel, so normalised activations never round-trip HBM.
#include "ops.h"

namespace keep {
namespace {

constexpr int GN_MAX_CHUNKS = 296;   // two waves of 148 SMs per image; also bounds the fused finalize (one block reads them all)
constexpr int GN_TICKETS = 256;       // arrival counters of the fused finalize: one per image, self-resetting

// ~16K elements per block (16 float4 per thread): 512^2 x 64 -> 1024 blocks per image = ~7 resident blocks per SM
static inline int gn_num_chunks(int hw, int c) {
    int s = cdiv((long long)hw * c, 16384);
    return s < 1 ? 1 : (s > GN_MAX_CHUNKS ? GN_MAX_CHUNKS : s);
}

// Stage 1: per-(image, chunk, channel) partial sums -> partial[n][chunk][c][2] (double).  Pure streaming read.
// Fused finalize (tickets != null): the block that arrives LAST for its image reduces all of that image's partials in
// chunk order (deterministic) and writes the per-channel affine -- the second kernel of the pair disappears (~4 us of a
// ~12 us pair on the serial per-frame chain).
template <typename T>
__global__ void __launch_bounds__(256) gn_partial_kernel(const T* __restrict__ x, int hw, int c, int nchunks,
                                                         double* __restrict__ partial, int* __restrict__ tickets, int cpg, float eps,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ scale, float* __restrict__ shift, int c_total, int c_off) {
    pdl_prologue_light();
    extern __shared__ double sm[];  // [lanes][c][2]
    __shared__ int s_last;
    const int c4 = c >> 2;
    const int lanes = blockDim.x / c4;
    const int cv = threadIdx.x % c4, lane = threadIdx.x / c4;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int per = (hw + nchunks - 1) / nchunks;
    const int p0 = chunk * per, p1 = min(hw, p0 + per);
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const T* base = x + (size_t)n * hw * c;
    if (lane < lanes) {
        int p = p0 + lane;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {      // four independent 16-byte loads in flight per thread
            const float4 v0 = ld4(base, (size_t)p * c + cv * 4);
            const float4 v1 = ld4(base, (size_t)(p + lanes) * c + cv * 4);
            const float4 v2 = ld4(base, (size_t)(p + 2 * lanes) * c + cv * 4);
            const float4 v3 = ld4(base, (size_t)(p + 3 * lanes) * c + cv * 4);
            // fp32 pre-sum of four values, then one double accumulate (sums of squares of O(1..100) values: safe in fp32x4)
            s[0] += (double)((v0.x + v1.x) + (v2.x + v3.x)); q[0] += (double)((v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x));
            s[1] += (double)((v0.y + v1.y) + (v2.y + v3.y)); q[1] += (double)((v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y));
            s[2] += (double)((v0.z + v1.z) + (v2.z + v3.z)); q[2] += (double)((v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z));
            s[3] += (double)((v0.w + v1.w) + (v2.w + v3.w)); q[3] += (double)((v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w));
        }
        for (; p < p1; p += lanes) {
            const float4 v = ld4(base, (size_t)p * c + cv * 4);
            s[0] += v.x; q[0] += (double)v.x * v.x;
            s[1] += v.y; q[1] += (double)v.y * v.y;
            s[2] += v.z; q[2] += (double)v.z * v.z;
            s[3] += v.w; q[3] += (double)v.w * v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sm[((size_t)lane * c + cv * 4 + j) * 2 + 0] = s[j];
            sm[((size_t)lane * c + cv * 4 + j) * 2 + 1] = q[j];
        }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        double ss = 0, qq = 0;
        for (int l = 0; l < lanes; ++l) {
            ss += sm[((size_t)l * c + ch) * 2 + 0];
            qq += sm[((size_t)l * c + ch) * 2 + 1];
        }
        double* o = partial + (((size_t)n * nchunks + chunk) * c + ch) * 2;
        o[0] = ss;
        o[1] = qq;
    }
    if (!tickets) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(&tickets[n], 1);
        s_last = (t == nchunks - 1) ? 1 : 0;
        if (s_last) tickets[n] = 0;          // ready for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // per-channel totals over the chunks (fixed order), into shared memory
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        double ss = 0, qq = 0;
        const double* o = partial + ((size_t)n * nchunks * c + ch) * 2;
        for (int k = 0; k < nchunks; ++k) {
            ss += __ldcg(o + (size_t)k * c * 2);
            qq += __ldcg(o + (size_t)k * c * 2 + 1);
        }
        sm[ch * 2] = ss;
        sm[ch * 2 + 1] = qq;
    }
    __syncthreads();
    const int groups = c / cpg;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double ss = 0, qq = 0;
        for (int i = 0; i < cpg; ++i) { ss += sm[(g * cpg + i) * 2]; qq += sm[(g * cpg + i) * 2 + 1]; }
        const double cnt = (double)hw * cpg;
        const double mean = ss / cnt;
        double var = qq / cnt - mean * mean;
        if (var < 0) var = 0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        for (int i = 0; i < cpg; ++i) {
            const int ch = g * cpg + i;
            const float ga = gamma ? gamma[c_off + ch] : 1.0f;
            const float be = beta ? beta[c_off + ch] : 0.0f;
            const float sc = ga * rstd;
            scale[(size_t)n * c_total + c_off + ch] = sc;
            shift[(size_t)n * c_total + c_off + ch] = be - (float)mean * sc;
        }
    }
}

// Stage 2: one 128-thread block per (image, group): fixed-order reduction of the partials -> per-(n, channel) affine
__global__ void __launch_bounds__(128) gn_finalize_kernel(const double* __restrict__ partial, int hw, int c, int cpg, int nchunks,
                                                          float eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ scale,
                                                          float* __restrict__ shift, int c_total, int c_off) {
    pdl_prologue_tiny();
    __shared__ double rs[4], rq[4];
    const int groups = c / cpg;
    const int in = blockIdx.x / groups, g = blockIdx.x % groups;
    double s = 0, q = 0;
    const int total = nchunks * cpg;
    for (int i = threadIdx.x; i < total; i += 128) {
        const int chunk = i / cpg, ch = g * cpg + i % cpg;
        const double* o = partial + (((size_t)in * nchunks + chunk) * c + ch) * 2;
        s += o[0];
        q += o[1];
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    s = (rs[0] + rs[1]) + (rs[2] + rs[3]);
    q = (rq[0] + rq[1]) + (rq[2] + rq[3]);
    const double cnt = (double)hw * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    for (int i = threadIdx.x; i < cpg; i += 128) {
        const int ch = g * cpg + i;
        const float ga = gamma ? gamma[c_off + ch] : 1.0f;
        const float be = beta ? beta[c_off + ch] : 0.0f;
        const float sc = ga * rstd;
        scale[(size_t)in * c_total + c_off + ch] = sc;
        shift[(size_t)in * c_total + c_off + ch] = be - (float)mean * sc;
    }
}

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows, int c, const float* __restrict__ g,
                                                        const float* __restrict__ b, float eps, const float* __restrict__ res,
                                                        float* __restrict__ out, const float* __restrict__ add2, int add2_rows,
                                                        float* __restrict__ out2) {
    pdl_prologue_tiny();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * c;
    float v[32];  // c <= 1024
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int ch = lane + i * 32;
        v[i] = ch < c ? xr[ch] : 0.0f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float)c;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int ch = lane + i * 32;
        const float d = ch < c ? v[i] - mean : 0.0f;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int ch = lane + i * 32;
        if (ch < c) {
            float y = (v[i] - mean) * rstd * g[ch] + b[ch];
            if (res) y += res[(size_t)row * c + ch];
            out[(size_t)row * c + ch] = y;
            if (out2) out2[(size_t)row * c + ch] = y + add2[(size_t)(row % add2_rows) * c + ch];
        }
    }
}
// vectorised variant: c % 4 == 0, NV4 float4 per lane (c <= 128 * NV4); one warp per row, two-pass statistics in registers
template <int NV4>
__global__ void __launch_bounds__(256) layernorm_v4_kernel(const float* __restrict__ x, int rows, int c, const float* __restrict__ g,
                                                           const float* __restrict__ b, float eps, const float* __restrict__ res,
                                                           float* __restrict__ out, const float* __restrict__ add2, int add2_rows,
                                                           float* __restrict__ out2) {
    pdl_prologue_tiny();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * c);
    const int c4 = c >> 2;
    float4 v[NV4];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int q = lane + i * 32;
        v[i] = q < c4 ? xr[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)c;
    float qq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        if (lane + i * 32 < c4) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            qq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
    const float rstd = rsqrtf(warp_sum(qq) / (float)c + eps);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int q = lane + i * 32;
        if (q < c4) {
            const float4 gg = reinterpret_cast<const float4*>(g)[q], bb = reinterpret_cast<const float4*>(b)[q];
            float4 y;
            y.x = (v[i].x - mean) * rstd * gg.x + bb.x; y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
            y.z = (v[i].z - mean) * rstd * gg.z + bb.z; y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
            if (res) {
                const float4 r = reinterpret_cast<const float4*>(res + (size_t)row * c)[q];
                y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
            }
            reinterpret_cast<float4*>(out + (size_t)row * c)[q] = y;
            if (out2) {
                const float4 a2 = reinterpret_cast<const float4*>(add2 + (size_t)(row % add2_rows) * c)[q];
                reinterpret_cast<float4*>(out2 + (size_t)row * c)[q] = make_float4(y.x + a2.x, y.y + a2.y, y.z + a2.z, y.w + a2.w);
            }
        }
    }
}
}  // namespace

namespace {
// Small tensors (the 16^2..64^2 maps of the serial per-frame chain): one launch, one block per (image, group) reads the
// group's hw x cpg slab and writes its channels' affine directly -- no partial buffer, no second kernel.
template <typename T>
__global__ void __launch_bounds__(256) gn_small_kernel(const T* __restrict__ x, int hw, int c, int cpg, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ scale, float* __restrict__ shift, int c_total, int c_off) {
    pdl_prologue_tiny();
    __shared__ double rs[8], rq[8];
    const int groups = c / cpg;
    const int in = blockIdx.x / groups, g = blockIdx.x % groups;
    const T* base = x + (size_t)in * hw * c + (size_t)g * cpg;
    double s = 0, q = 0;
    if ((cpg & 3) == 0) {
        const int v4 = cpg >> 2;
        const int total = hw * v4;
        for (int i = threadIdx.x; i < total; i += 256) {
            const int p = i / v4, j = i - p * v4;
            const float4 v = ld4(base, (size_t)p * c + j * 4);
            s += (double)((v.x + v.y) + (v.z + v.w));
            q += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
        }
    } else {
        const int total = hw * cpg;
        for (int i = threadIdx.x; i < total; i += 256) {
            const int p = i / cpg, j = i - p * cpg;
            const float v = ldf(base, (size_t)p * c + j);
            s += v;
            q += (double)v * v;
        }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    s = ((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + (rs[6] + rs[7]));
    q = ((rq[0] + rq[1]) + (rq[2] + rq[3])) + ((rq[4] + rq[5]) + (rq[6] + rq[7]));
    const double cnt = (double)hw * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    for (int i = threadIdx.x; i < cpg; i += 256) {
        const int ch = g * cpg + i;
        const float ga = gamma ? gamma[c_off + ch] : 1.0f;
        const float be = beta ? beta[c_off + ch] : 0.0f;
        const float sc = ga * rstd;
        scale[(size_t)in * c_total + c_off + ch] = sc;
        shift[(size_t)in * c_total + c_off + ch] = be - (float)mean * sc;
    }
}
}  // namespace

namespace {
// Statistics emitted by the producing kernel (conv epilogue / split-K reduce): part[img][32 groups][P][2] fp32 (sum, sum of
// squares per slot).  One block per (image, group) streams the group's P contiguous slots (float4 = two slots per load, four
// loads in flight per thread), accumulates in double and reduces in a fixed order -> the consumer's per-(n, channel) affine.
// Replaces the read pass over the whole tensor (gn_partial / gn_small) and its second kernel.
__global__ void __launch_bounds__(128) gn_finalize_parts_kernel(const float* __restrict__ part, int P, int c, int cpg, double cnt, float eps,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ scale, float* __restrict__ shift) {
    pdl_prologue_tiny();
    __shared__ double rs[4], rq[4];
    const int img = blockIdx.x >> 5, g = blockIdx.x & 31;
    const float4* base = reinterpret_cast<const float4*>(part + ((size_t)img * 32 + g) * P * 2);   // P is even: P / 2 float4
    const int n4 = P >> 1;
    double s = 0, q = 0;
    int k = threadIdx.x;
    for (; k + 3 * 128 < n4; k += 4 * 128) {
        const float4 a = __ldcg(base + k), b = __ldcg(base + k + 128), c4 = __ldcg(base + k + 256), d = __ldcg(base + k + 384);
        s += (double)((a.x + a.z) + (b.x + b.z)) + (double)((c4.x + c4.z) + (d.x + d.z));
        q += (double)((a.y + a.w) + (b.y + b.w)) + (double)((c4.y + c4.w) + (d.y + d.w));
    }
    for (; k < n4; k += 128) {
        const float4 a = __ldcg(base + k);
        s += (double)(a.x + a.z);
        q += (double)(a.y + a.w);
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    s = (rs[0] + rs[1]) + (rs[2] + rs[3]);
    q = (rq[0] + rq[1]) + (rq[2] + rq[3]);
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    for (int i = threadIdx.x; i < cpg; i += 128) {
        const int ch = g * cpg + i;
        const float sc = gamma[ch] * rstd;
        scale[(size_t)img * c + ch] = sc;
        shift[(size_t)img * c + ch] = beta[ch] - (float)mean * sc;
    }
}
}  // namespace

void gn_finalize_parts(const float* part, int n, int P, int hw, int c, float eps, const float* gamma, const float* beta, float* scale,
                       float* shift, cudaStream_t s) {
    KEEP_CHECK(c % 32 == 0 && P > 0 && P % 2 == 0 && gamma && beta, "gn_finalize_parts: bad arguments (c=%d, P=%d)", c, P);
    const int cpg = c / 32;
    launch_k(gn_finalize_parts_kernel, dim3(n * 32), dim3(128), 0, s, part, P, c, cpg, (double)hw * cpg, eps, gamma, beta, scale, shift);
    CUDA_CHECK(cudaGetLastError());
}

size_t gn_scratch_doubles(int n, int hw, int c) { return (size_t)n * gn_num_chunks(hw, c) * c * 2; }

void gn_warmup() {}
int gn_ticket_count() { return GN_TICKETS; }

void groupnorm_affine(const void* x, int dt, int n, int hw, int c, int cpg, float eps, const float* gamma, const float* beta,
                      float* scale, float* shift, int c_total, int c_off, double* scratch, cudaStream_t s, int* ticket_buf) {
    KEEP_CHECK(c % 4 == 0 && c / 4 <= 256 && c % cpg == 0, "groupnorm: unsupported shape c=%d cpg=%d", c, cpg);
    if ((long long)hw * cpg <= 8192 && (long long)n * (c / cpg) >= 16) {   // small slab per group and enough groups to fill SMs
        const int blocks = n * (c / cpg);
        if (dt == F32) launch_k(gn_small_kernel<float>, dim3(blocks), dim3(256), 0, s, (const float*)x, hw, c, cpg, eps, gamma, beta, scale, shift, c_total, c_off);
        else launch_k(gn_small_kernel<__half>, dim3(blocks), dim3(256), 0, s, (const __half*)x, hw, c, cpg, eps, gamma, beta, scale, shift, c_total, c_off);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    const int nchunks = gn_num_chunks(hw, c);
    const int c4 = c / 4;
    const int lanes = 256 / c4 > 0 ? 256 / c4 : 1;
    const int threads = ((lanes * c4 + 31) / 32) * 32;
    const size_t smem = (size_t)lanes * c * 2 * sizeof(double);
    dim3 grid(nchunks, n);
    // opt-in (KEEP_GN_FUSED=1): measured 135.2 -> 129.0 frames/s -- one block reducing nchunks x c doubles serially costs more
    // than the ~4 us slot of a second, 32-block kernel (profiles/r1_experiments.md)
    static const bool fuse_en = getenv("KEEP_GN_FUSED") && getenv("KEEP_GN_FUSED")[0] == '1';
    int* tickets = (fuse_en && ticket_buf && n <= GN_TICKETS) ? ticket_buf : nullptr;
    if (dt == F32) launch_k(gn_partial_kernel<float>, dim3(grid), dim3(threads), smem, s, (const float*)x, hw, c, nchunks, scratch, tickets, cpg, eps, gamma, beta, scale, shift, c_total, c_off);
    else launch_k(gn_partial_kernel<__half>, dim3(grid), dim3(threads), smem, s, (const __half*)x, hw, c, nchunks, scratch, tickets, cpg, eps, gamma, beta, scale, shift, c_total, c_off);
    CUDA_CHECK(cudaGetLastError());
    if (tickets) return;
    launch_k(gn_finalize_kernel, dim3(n * (c / cpg)), dim3(128), 0, s, scratch, hw, c, cpg, nchunks, eps, gamma, beta, scale, shift, c_total, c_off);
    CUDA_CHECK(cudaGetLastError());
}

void layernorm(const float* x, int rows, int c, const float* g, const float* b, float eps, const float* res, float* out,
               const float* add2, int add2_rows, float* out2, cudaStream_t s) {
    KEEP_CHECK(c <= 1024, "layernorm: c=%d > 1024", c);
    const int ar = add2_rows > 0 ? add2_rows : 1;
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b) |
                      reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(add2) | reinterpret_cast<uintptr_t>(out2)) & 15) == 0;
    // few rows (the 256-token matrices of the per-frame chain): 2 rows per block spreads them over the SMs
    const int wpb = rows <= 2048 ? 2 : 8;
    const dim3 grid(cdiv(rows, wpb)), block(wpb * 32);
    if (c % 4 == 0 && al && c <= 128) launch_k(layernorm_v4_kernel<1>, grid, block, 0, s, x, rows, c, g, b, eps, res, out, add2, ar, out2);
    else if (c % 4 == 0 && al && c <= 256) launch_k(layernorm_v4_kernel<2>, grid, block, 0, s, x, rows, c, g, b, eps, res, out, add2, ar, out2);
    else if (c % 4 == 0 && al && c <= 512) launch_k(layernorm_v4_kernel<4>, grid, block, 0, s, x, rows, c, g, b, eps, res, out, add2, ar, out2);
    else if (c % 4 == 0 && al) launch_k(layernorm_v4_kernel<8>, grid, block, 0, s, x, rows, c, g, b, eps, res, out, add2, ar, out2);
    else
    launch_k(layernorm_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, s, x, rows, c, g, b, eps, res, out, add2, ar, out2);
    CUDA_CHECK(cudaGetLastError());
}

KEEP_STAMP_SETTER(stamp_set_norm)

}  // namespace keep
