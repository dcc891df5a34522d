// keep_b200 — batched strided fp32 GEMM on CUDA cores for the attention contractions
// (QK^T and PV of every attention flavour on the path: AttnBlock vqgan_arch.py:231-238,
//  nn.MultiheadAttention keep_arch.py:431, CrossAttention._attention keep_arch.py:205-234,
//  GMFlow window / global attention gmflow/transformer.py:12-14,88-96, matching.py:16,31).
// Head split / merge, the temporal "(b f) d c -> (b d) f c" regrouping (keep_arch.py:673-680)
// and window batching are expressed as 3-level batch strides, so no permute copies are made.
#include "ops.h"

namespace keep {
namespace {
constexpr int TM = 64, TN = 64, TK = 16;

template <bool TRANSB>
__global__ void __launch_bounds__(256) bgemm_kernel(const BGemmArgs a, int vecA, int vecB) {
    pdl_prologue();
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    int z = blockIdx.z;
    const int z2 = z % a.nz2; z /= a.nz2;
    const int z1 = z % a.nz1; z /= a.nz1;
    const int z0 = z;
    const float* A = a.A + z0 * a.sA[0] + z1 * a.sA[1] + z2 * a.sA[2];
    const float* B = a.B + z0 * a.sB[0] + z1 * a.sB[1] + z2 * a.sB[2];
    float* C = a.C + z0 * a.sC[0] + z1 * a.sC[1] + z2 * a.sC[2];
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;

    const int lr = tid >> 2, lq = (tid & 3) * 4;    // row-major [row][k] loader: 64 rows x 16 k
    const int kr = tid >> 4, nq = (tid & 15) * 4;   // [k][n] loader: 16 k x 64 n
    float ra[4], rb[4];

    auto load = [&](int k0) {
        {   // A[m][k]
            const int m = m0 + lr, k = k0 + lq;
            if (m < a.M && vecA && k + 3 < a.K) {
                float4 t = *reinterpret_cast<const float4*>(A + (size_t)m * a.lda + k);
                ra[0] = t.x; ra[1] = t.y; ra[2] = t.z; ra[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) ra[j] = (m < a.M && k + j < a.K) ? A[(size_t)m * a.lda + k + j] : 0.0f;
            }
        }
        if (TRANSB) {  // B[n][k]
            const int n = n0 + lr, k = k0 + lq;
            if (n < a.N && vecB && k + 3 < a.K) {
                float4 t = *reinterpret_cast<const float4*>(B + (size_t)n * a.ldb + k);
                rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) rb[j] = (n < a.N && k + j < a.K) ? B[(size_t)n * a.ldb + k + j] : 0.0f;
            }
        } else {       // B[k][n]
            const int k = k0 + kr, n = n0 + nq;
            if (k < a.K && vecB && n + 3 < a.N) {
                float4 t = *reinterpret_cast<const float4*>(B + (size_t)k * a.ldb + n);
                rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) rb[j] = (k < a.K && n + j < a.N) ? B[(size_t)k * a.ldb + n + j] : 0.0f;
            }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const int ty = tid >> 4, tx = tid & 15;

    load(0);
    for (int k0 = 0; k0 < a.K; k0 += TK) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lq + j][lr] = ra[j];
        if (TRANSB) {
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[lq + j][lr] = rb[j];
        } else {
            *reinterpret_cast<float4*>(&Bs[kr][nq]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        }
        __syncthreads();
        if (k0 + TK < a.K) load(k0 + TK);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float* c = C + (size_t)m * a.ldc + n;
            const float v = a.alpha * acc[i][j];
            *c = a.accumulate ? (*c + v) : v;
        }
    }
}
// 32 x 32 output tile, K step 32, 2 x 2 outputs per thread: for the 256-token attentions of the serial per-frame chain,
// whose 64-wide tiling fills only 16-32 SMs (26 us for a 256 x 256 x 512 product).  Same summation order over k.
template <bool TRANSB>
__global__ void __launch_bounds__(256) bgemm32_kernel(const BGemmArgs a, int vecA, int vecB) {
    pdl_prologue_light();
    constexpr int T = 32, KT = 32;
    __shared__ __align__(16) float As[KT][T + 4];
    __shared__ __align__(16) float Bs[KT][T + 4];
    const int tid = threadIdx.x;
    int z = blockIdx.z;
    const int z2 = z % a.nz2; z /= a.nz2;
    const int z1 = z % a.nz1; z /= a.nz1;
    const int z0 = z;
    const float* A = a.A + z0 * a.sA[0] + z1 * a.sA[1] + z2 * a.sA[2];
    const float* B = a.B + z0 * a.sB[0] + z1 * a.sB[1] + z2 * a.sB[2];
    float* C = a.C + z0 * a.sC[0] + z1 * a.sC[1] + z2 * a.sC[2];
    const int m0 = blockIdx.x * T, n0 = blockIdx.y * T;
    const int lr = tid >> 3, lq = (tid & 7) * 4;    // [row][k] loader: 32 rows x 32 k, one float4 per thread
    float ra[4], rb[4];
    auto load4 = [&](const float* P, int ld, int row, int rmax, int col, int cmax, int vec, float (&r)[4]) {
        if (row < rmax && vec && col + 3 < cmax) {
            const float4 t = *reinterpret_cast<const float4*>(P + (size_t)row * ld + col);
            r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = (row < rmax && col + j < cmax) ? P[(size_t)row * ld + col + j] : 0.0f;
        }
    };
    auto load = [&](int k0) {
        load4(A, a.lda, m0 + lr, a.M, k0 + lq, a.K, vecA, ra);                       // A[m][k]
        if (TRANSB) load4(B, a.ldb, n0 + lr, a.N, k0 + lq, a.K, vecB, rb);            // B[n][k]
        else load4(B, a.ldb, k0 + lr, a.K, n0 + lq, a.N, vecB, rb);                   // B[k][n]: 32 k x 32 n
    };
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const int ty = tid >> 4, tx = tid & 15;
    load(0);
    for (int k0 = 0; k0 < a.K; k0 += KT) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lq + j][lr] = ra[j];
        if (TRANSB) {
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[lq + j][lr] = rb[j];
        } else {
            *reinterpret_cast<float4*>(&Bs[lr][lq]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        }
        __syncthreads();
        if (k0 + KT < a.K) load(k0 + KT);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const float2 av = *reinterpret_cast<const float2*>(&As[k][ty * 2]);
            const float2 bv = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
            acc[0][0] = fmaf(av.x, bv.x, acc[0][0]); acc[0][1] = fmaf(av.x, bv.y, acc[0][1]);
            acc[1][0] = fmaf(av.y, bv.x, acc[1][0]); acc[1][1] = fmaf(av.y, bv.y, acc[1][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + tx * 2 + j;
            if (n >= a.N) continue;
            float* c = C + (size_t)m * a.ldc + n;
            const float v = a.alpha * acc[i][j];
            *c = a.accumulate ? (*c + v) : v;
        }
    }
}
}  // namespace

void bgemm_simt(const BGemmArgs& a, cudaStream_t s) {
    KEEP_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "bgemm: empty problem");
    const long long nz = (long long)a.nz0 * a.nz1 * a.nz2;
    KEEP_CHECK(nz > 0 && nz <= 65535, "bgemm: batch %lld out of range", nz);
    auto al4 = [](const void* p, int ld, const long long* st) {
        return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0) && (st[0] % 4 == 0) && (st[1] % 4 == 0) &&
               (st[2] % 4 == 0);
    };
    const int vecA = al4(a.A, a.lda, a.sA), vecB = al4(a.B, a.ldb, a.sB);
    const long long tiles64 = (long long)cdiv(a.M, TM) * cdiv(a.N, TN) * nz;
    if (tiles64 < 148) {   // too few 64 x 64 tiles to fill the SMs: quarter-size tiles
        dim3 grid(cdiv(a.M, 32), cdiv(a.N, 32), (unsigned)nz);
        if (a.transB) launch_k(bgemm32_kernel<true>, dim3(grid), dim3(256), 0, s, a, vecA, vecB);
        else launch_k(bgemm32_kernel<false>, dim3(grid), dim3(256), 0, s, a, vecA, vecB);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    dim3 grid(cdiv(a.M, TM), cdiv(a.N, TN), (unsigned)nz);
    if (a.transB) launch_k(bgemm_kernel<true>, dim3(grid), dim3(256), 0, s, a, vecA, vecB);
    else launch_k(bgemm_kernel<false>, dim3(grid), dim3(256), 0, s, a, vecA, vecB);
    CUDA_CHECK(cudaGetLastError());
}

KEEP_STAMP_SETTER(stamp_set_gemm)

}  // namespace keep
