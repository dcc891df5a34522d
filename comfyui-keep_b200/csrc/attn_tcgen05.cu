// keep_b200 — fused attention on tcgen05 for sm_100a: S = (Q K^T) * scale (+ shifted-window mask) -> softmax -> P V in ONE
// kernel; the score matrix lives in TMEM / shared memory only and never reaches HBM.
//
// Replaces, for GMFlow's swin window attention (gmflow/transformer.py:46-105: 1024 x 1024 tokens per window, d = 128, region
// mask in {0, -100} on shifted layers), the round-1 sequence  pack(K) -> QK^T GEMM (writes 4 MB of fp32 scores per window) ->
// softmax kernel (reads + writes them) -> pack(V^T) -> PV GEMM (reads them again): 536 MB of HBM traffic per attention call
// on a 4-pair chunk, 12 calls per chunk.
//
// One CTA = 128 queries of one batch element (window).  Split-precision operands like the convolution kernel
// (conv_tcgen05.cu): every fp32 value is carried as an fp16 (hi, lo) pair, three MMAs per K step (lo*hi, hi*lo, hi*hi) with
// fp32 accumulation in TMEM -> fp32-grade scores and outputs.
//
//   warps 0-3   softmax + epilogue: thread = query row = TMEM lane
//   warp  4     MMA issuer (one elected lane)
//   warps 5-12  producers: fp32 Q / K / V from global -> (hi, lo) fp16 in the UMMA SWIZZLE_128B K-major layout;
//               V is transposed on the way (the B operand of P V is V^T: rows = head dims, K = keys)
//
// Two passes over the keys instead of an online rescale of O: pass A computes the scores block by block (64 keys per
// block) and keeps only the running row maximum; pass B recomputes each score block (bitwise the same MMAs), forms
// P = exp(s - max) -- final, never rescaled --, stores it as the A operand of the second GEMM and accumulates O += P V_j in
// TMEM.  The extra Q K^T costs ~40 % more MMA work on a kernel that is bound by its softmax / producer warps, and removes
// the TMEM read-modify-write of O.  S is double-buffered in TMEM (2 x 64 columns) so the MMAs of block j+1 overlap the
// softmax of block j; K and V stages are double-buffered in shared memory.
#include "ops.h"

namespace keep {
namespace {

constexpr int AT_THREADS = 416;            // 13 warps
constexpr int AT_PROD = 256;               // producer threads (warps 5-12)
constexpr int AT_KB = 64;                  // keys per block

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void at_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D TMA: `bytes` (multiple of 16) global -> shared, completion counted on the mbarrier
__device__ __forceinline__ void at_tma_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void at_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void at_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void at_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void at_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void at_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void at_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool at_elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct AttnTcArgs {
    const float* q; const float* k; const float* v; float* out;
    long long q_bs, k_bs, v_bs, o_bs;   // batch strides (elements)
    int ldq, ldk, ldv, ldo;             // row strides (elements)
    int nb, Lq, Lk;
    float scale;
    const unsigned char* region;        // [n_win][Lk] region ids of a shifted-window layer (Lq == Lk), or null
    int n_win;
    // swin window mode (win_side > 0): q / k / v / out are whole token maps (map_w x map_w tokens per image, row stride ld*,
    // image stride *_bs) and batch z = image * win_side^2 + window; row r of a window is token
    //   ((wy * wsz + r / wsz + shift) mod map_w) * map_w + (wx * wsz + r mod wsz + shift) mod map_w
    // i.e. the window partition (with its cyclic shift, gmflow/transformer.py:78-98) and the merge are index math here
    int win_side, wsz_log2, map_w, shift;
    // multi-head mode (heads > 1, plain rows only): batch z = b * heads + h, head h = columns [h * dh, (h + 1) * dh) of the rows
    int heads;
    // K / V already converted by attn_pack_kv_kernel (null: the producers convert them, once per query tile): per batch z and
    // key block j the byte image of one K stage (NCB x 64 rows x 128 B) and one V^T stage (2 x DH rows x 128 B) -- a query tile's
    // stages then arrive by 1-D TMA instead of being re-derived from fp32 by every one of the window's Lq / 128 query tiles
    const uint8_t* kpack; const uint8_t* vpack;
    long long* trace;   // debug (keepop_attn_trace): 8 x 40 clock64 stamps of CTA 0, or null
};
#define AT_TRACE(slot, idx)                                                                       \
    do {                                                                                          \
        if (a.trace && blockIdx.x == 0 && (idx) < 40) a.trace[(slot) * 40 + (idx)] = clock64();   \
    } while (0)

// window row -> token index (plain mode: identity)
__device__ __forceinline__ int at_token(const AttnTcArgs& a, int win, int r) {
    if (a.win_side == 0) return r;
    const int wsz = 1 << a.wsz_log2;
    const int wy = win / a.win_side, wx = win - wy * a.win_side;
    const int y = (wy * wsz + (r >> a.wsz_log2) + a.shift) & (a.map_w - 1);
    const int x = (wx * wsz + (r & (wsz - 1)) + a.shift) & (a.map_w - 1);
    return y * a.map_w + x;
}

// 8 fp32 -> (hi, lo) fp16 pairs: hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void at_split8(const float* v, uint4& hi, uint4& lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 r = __half22float2(h[j]);
        l[j] = __floats2half2_rn(v[2 * j] - r.x, v[2 * j + 1] - r.y);
    }
    hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                    *reinterpret_cast<uint32_t*>(&h[3]));
    lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]),
                    *reinterpret_cast<uint32_t*>(&l[3]));
}

// shared-memory plan (bytes) for head dim DH: Q (DH/32 channel blocks x 128 rows x 128 B) | 2 K stages (DH/32 x 64 rows x 128 B)
// | 2 V^T stages (2 key sub-blocks x DH rows x 128 B) | P (2 key sub-blocks x 128 rows x 128 B) | region ids | barriers
template <int DH>
struct AttnSmem {
    static constexpr int NCB = DH / 32;
    static constexpr int Q_BYTES = NCB * 128 * 128;
    static constexpr int K_STAGE = NCB * AT_KB * 128;
    static constexpr int V_STAGE = 2 * DH * 128;
    static constexpr int P_BYTES = 2 * 128 * 128;
    static constexpr int OFF_K = Q_BYTES;
    static constexpr int OFF_V = OFF_K + 2 * K_STAGE;
    static constexpr int OFF_P = OFF_V + 2 * V_STAGE;
    static constexpr int OFF_REG = OFF_P + P_BYTES;
    static constexpr int OFF_BAR = OFF_REG + 1024;
    static constexpr int TOTAL = OFF_BAR + 256 + 1024;   // + alignment slack
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_tc_kernel(const AttnTcArgs a) {
    using SM = AttnSmem<DH>;
    constexpr int NCB = SM::NCB;
    extern __shared__ __align__(1024) uint8_t at_smem_raw[];
    uint8_t* smem = at_smem_raw + ((1024u - (at_smem_u32(at_smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + SM::OFF_K;
    uint8_t* sV = smem + SM::OFF_V;
    uint8_t* sP = smem + SM::OFF_P;
    uint8_t* sReg = smem + SM::OFF_REG;
    const uint32_t bar0 = at_smem_u32(smem + SM::OFF_BAR);
    // barrier map
    const uint32_t Q_FULL = bar0;
    auto K_FULL = [&](int s) { return bar0 + 8u * (1 + s); };
    auto K_EMPTY = [&](int s) { return bar0 + 8u * (3 + s); };
    auto V_FULL = [&](int s) { return bar0 + 8u * (5 + s); };
    auto V_EMPTY = [&](int s) { return bar0 + 8u * (7 + s); };
    auto S_FULL = [&](int s) { return bar0 + 8u * (9 + s); };
    auto S_EMPTY = [&](int s) { return bar0 + 8u * (11 + s); };
    const uint32_t P_FULL = bar0 + 8u * 13, P_EMPTY = bar0 + 8u * 14, O_FULL = bar0 + 8u * 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::OFF_BAR + 8 * 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) AT_TRACE(7, 0);
    const int qtiles = a.Lq >> 7;
    const int z = blockIdx.x / qtiles, qt = blockIdx.x - z * qtiles;
    const int NKB = a.Lk / AT_KB;
    const int nwin = a.win_side > 0 ? a.win_side * a.win_side : 1;
    const int img = a.win_side > 0 ? z / nwin : (a.heads > 1 ? z / a.heads : z);              // tensor batch index
    const int win = a.win_side > 0 ? z - img * nwin : 0;                                      // window within the image
    const int hoff = a.heads > 1 ? (z - img * a.heads) * DH : 0;                               // column offset of this head

    if (threadIdx.x == 0) {
        at_mbar_init(Q_FULL, AT_PROD);
        for (int s = 0; s < 2; ++s) {
            // K / V stages: one producer group each, or (packed operands) one TMA-issuing thread
            at_mbar_init(K_FULL(s), a.kpack ? 1 : AT_PROD / 2); at_mbar_init(K_EMPTY(s), 1);
            at_mbar_init(V_FULL(s), a.kpack ? 1 : AT_PROD / 2); at_mbar_init(V_EMPTY(s), 1);
            at_mbar_init(S_FULL(s), 1); at_mbar_init(S_EMPTY(s), 128);
        }
        at_mbar_init(P_FULL, 128); at_mbar_init(P_EMPTY, 1); at_mbar_init(O_FULL, 1);
        at_fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = 256;              // S0 [0,64) | S1 [64,128) | O [128, 128 + DH)
    if (warp == 4) at_tmem_alloc(at_smem_u32(tmem_slot), TMEM_COLS);
    at_tc_fence_before();
    __syncthreads();
    at_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 5) {
        // =========================== producers ===========================
        // Two groups of four warps.  A stage's critical path is global-load latency + conversion + the proxy fence (which also
        // waits for any load the thread has in flight, so a thread cannot prefetch across stages): in pass A the groups take
        // alternate K stages, in pass B group 0 produces the K blocks and group 1 the (transposed) V blocks, so one
        // stage's load latency hides behind another's conversion and behind the MMAs / softmax of the block before.
        const int pt = threadIdx.x - 5 * 32;          // 0 .. 255
        const int grp = pt >> 7, gt = pt & 127;       // producer group, thread within the group
        pdl_wait();                                   // q / k / v come from the previous kernels
        if (pt == 0) keep_stamp_here();
        // ---- Q tile (scaled): 128 rows x DH channels, unit = 8 channels (all 256 threads)
        {
            const float* qb = a.q + (size_t)img * a.q_bs + hoff;
            constexpr int UPR = NCB * 4;              // units per row
#pragma unroll 2
            for (int u = pt; u < 128 * UPR; u += AT_PROD) {
                const int row = u / UPR, cu = u - row * UPR, cb = cu >> 2, pl = cu & 3;
                float v[8];
                ldg256(qb + (size_t)at_token(a, win, qt * 128 + row) * a.ldq + cb * 32 + pl * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= a.scale;
                uint4 hi, lo;
                at_split8(v, hi, lo);
                uint8_t* dst = sQ + cb * (128 * 128) + row * 128 + ((pl ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>((uintptr_t)dst ^ 64)) = lo;
            }
            at_fence_proxy_async();
            at_mbar_arrive(Q_FULL);
            if (pt == 0) AT_TRACE(7, 1);                                // slot 7: [0] kernel entry, [1] Q produced, [2] done
        }
        const float* kb = a.k + (size_t)img * a.k_bs + hoff;
        const float* vb = a.v + (size_t)img * a.v_bs + hoff;
        // ---- K block j into stage buffer t & 1 (rows = keys), by the 128 threads of one group
        auto produce_K = [&](int t) {
            const int j = t < NKB ? t : t - NKB, s = t & 1;
            constexpr int UPR = NCB * 4;
            constexpr int NU = (AT_KB * UPR) / (AT_PROD / 2);          // units per thread (8 at DH = 128)
            float v[NU][8];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int u = gt + i * (AT_PROD / 2);
                const int row = u / UPR, cu = u - row * UPR;
                ldg256(kb + (size_t)at_token(a, win, j * AT_KB + row) * a.ldk + (cu >> 2) * 32 + (cu & 3) * 8, v[i]);
            }
            at_mbar_wait(K_EMPTY(s), (uint32_t)(((t >> 1) & 1) ^ 1));
            uint8_t* dstK = sK + s * SM::K_STAGE;
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                const int u = gt + i * (AT_PROD / 2);
                const int row = u / UPR, cu = u - row * UPR, cb = cu >> 2, pl = cu & 3;
                uint4 hi, lo;
                at_split8(v[i], hi, lo);
                uint8_t* dst = dstK + cb * (AT_KB * 128) + row * 128 + ((pl ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = hi;
                if (t >= NKB) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>((uintptr_t)dst ^ 64)) = lo;   // (pass A reads hi only)
            }
            at_fence_proxy_async();
            at_mbar_arrive(K_FULL(s));
            if (gt == 0) AT_TRACE(0, t);                                // slot 0: K stage t produced
        };
        // ---- V block j, transposed: B operand rows = head dims, K = the block's 64 keys (2 sub-blocks of 32).
        // lane = key within the sub-block: the 32 lanes of a warp write 32 consecutive halfs of one row (conflict-free)
        auto produce_V = [&](int j) {
            const int sv = j & 1;
            const int pw = gt >> 5;                                  // warp within the group: 0 .. 3
            constexpr int NCOMBO = 2 * (DH / 8);                     // (key sub-block, group of 8 head dims)
            constexpr int NC = NCOMBO / 4;                           // combos per warp (8 at DH = 128)
            float v[NC][8];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int c = pw + i * 4, kbk = c / (DH / 8), dg = c - kbk * (DH / 8);
                ldg256(vb + (size_t)at_token(a, win, j * AT_KB + kbk * 32 + lane) * a.ldv + dg * 8, v[i]);
            }
            at_mbar_wait(V_EMPTY(sv), (uint32_t)(((j >> 1) & 1) ^ 1));
            uint8_t* dstV = sV + sv * SM::V_STAGE;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int c = pw + i * 4, kbk = c / (DH / 8), dg = c - kbk * (DH / 8);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int row = dg * 8 + e;
                    const __half h = __float2half_rn(v[i][e]);
                    const __half l = __float2half_rn(v[i][e] - __half2float(h));
                    uint8_t* rowp = dstV + kbk * (DH * 128) + row * 128;
                    *reinterpret_cast<__half*>(rowp + ((((lane >> 3)) ^ (row & 7)) << 4) + (lane & 7) * 2) = h;
                    *reinterpret_cast<__half*>(rowp + ((((lane >> 3) + 4) ^ (row & 7)) << 4) + (lane & 7) * 2) = l;
                }
            }
            at_fence_proxy_async();
            at_mbar_arrive(V_FULL(sv));
            if (gt == 0) AT_TRACE(1, j);                                // slot 1: V stage j produced
        };
        if (a.kpack) {
            // packed operands: thread 0 of group 0 streams the K stages of both passes, thread 0 of group 1 the V stages
            if (gt == 0 && grp == 0) {
                const uint8_t* src = a.kpack + (size_t)z * NKB * SM::K_STAGE;
                for (int t = 0; t < 2 * NKB; ++t) {
                    const int j = t < NKB ? t : t - NKB, st = t & 1;
                    at_mbar_wait(K_EMPTY(st), (uint32_t)(((t >> 1) & 1) ^ 1));
                    at_mbar_arrive_expect_tx(K_FULL(st), (uint32_t)SM::K_STAGE);
                    at_tma_g2s(at_smem_u32(sK + st * SM::K_STAGE), src + (size_t)j * SM::K_STAGE, (uint32_t)SM::K_STAGE, K_FULL(st));
                    AT_TRACE(0, t);
                }
            } else if (gt == 0) {
                const uint8_t* src = a.vpack + (size_t)z * NKB * SM::V_STAGE;
                for (int j = 0; j < NKB; ++j) {
                    const int sv = j & 1;
                    at_mbar_wait(V_EMPTY(sv), (uint32_t)(((j >> 1) & 1) ^ 1));
                    at_mbar_arrive_expect_tx(V_FULL(sv), (uint32_t)SM::V_STAGE);
                    at_tma_g2s(at_smem_u32(sV + sv * SM::V_STAGE), src + (size_t)j * SM::V_STAGE, (uint32_t)SM::V_STAGE, V_FULL(sv));
                    AT_TRACE(1, j);
                }
            }
        } else {
        for (int t = grp; t < NKB; t += 2) produce_K(t);              // pass A: alternate stages (NKB is even or the tail is group 0's)
        if (grp == 0) {
            for (int t = NKB; t < 2 * NKB; ++t) produce_K(t);         // pass B: every K block
        } else {
            for (int j = 0; j < NKB; ++j) produce_V(j);               // pass B: every V block
        }
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        const bool leader = at_elect_one();
        constexpr uint32_t hi_desc = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
        constexpr uint32_t lo0 = 1u << 16;
        const uint32_t idesc_s = (1u << 4) | ((uint32_t)(AT_KB >> 3) << 17) | ((128u >> 4) << 24);   // M 128, N 64, f16 x f16 -> f32
        const uint32_t idesc_o = (1u << 4) | ((uint32_t)(DH >> 3) << 17) | ((128u >> 4) << 24);      // M 128, N DH
        auto issue = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, bool hi_only) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {                       // two K steps of 16 per 32-wide block; lo half 64 B further
                const uint64_t ad = ((uint64_t)hi_desc << 32) | (a_lo + k * 2u);
                const uint64_t bd = ((uint64_t)hi_desc << 32) | (b_lo + k * 2u);
                if (hi_only) {
                    at_umma(d_tmem, ad, bd, idesc, k == 0 ? acc : 1u);
                } else {
                    at_umma(d_tmem, ad + 4u, bd, idesc, k == 0 ? acc : 1u);
                    at_umma(d_tmem, ad, bd + 4u, idesc, 1u);
                    at_umma(d_tmem, ad, bd, idesc, 1u);
                }
            }
        };
        const uint32_t q_base = lo0 | (at_smem_u32(sQ) >> 4);
        auto issue_S = [&](int t) {      // t-th score block overall (pass A: 0 .. NKB-1, pass B: NKB .. 2 NKB - 1)
            const int s = t & 1;
            at_mbar_wait(K_FULL(s), (uint32_t)((t >> 1) & 1));
            at_mbar_wait(S_EMPTY(s), (uint32_t)(((t >> 1) & 1) ^ 1));
            at_tc_fence_after();
            if (leader) {
                const uint32_t k_base = lo0 | (at_smem_u32(sK + s * SM::K_STAGE) >> 4);
#pragma unroll
                // pass A only feeds the row maximum -- a stabiliser, any value near the true maximum gives the same softmax --, so
                // one MMA per K step (hi x hi, ~1e-3 relative) instead of three
                for (int cb = 0; cb < NCB; ++cb)
                    issue(tmem_base + (uint32_t)(s * AT_KB), q_base + (uint32_t)((cb * 128 * 128) >> 4), k_base + (uint32_t)((cb * AT_KB * 128) >> 4),
                          idesc_s, cb ? 1u : 0u, t < NKB);
                at_commit(K_EMPTY(s));
                at_commit(S_FULL(s));
            }
            __syncwarp();
            if (lane == 0) AT_TRACE(2, t);                              // slot 2: scores of block t issued
        };
        at_mbar_wait(Q_FULL, 0);
        at_tc_fence_after();
        for (int t = 0; t < NKB; ++t) issue_S(t);                  // pass A: scores only (row maxima)
        issue_S(NKB);                                              // pass B
        for (int j = 0; j < NKB; ++j) {
            if (j + 1 < NKB) issue_S(NKB + j + 1);                 // next block's scores overlap this block's softmax
            const int sv = j & 1;
            at_mbar_wait(P_FULL, (uint32_t)(j & 1));
            at_mbar_wait(V_FULL(sv), (uint32_t)((j >> 1) & 1));
            at_tc_fence_after();
            if (leader) {
                const uint32_t p_base = lo0 | (at_smem_u32(sP) >> 4);
                const uint32_t v_base = lo0 | (at_smem_u32(sV + sv * SM::V_STAGE) >> 4);
#pragma unroll
                for (int kbk = 0; kbk < 2; ++kbk)
                    issue(tmem_base + 128u, p_base + (uint32_t)((kbk * 128 * 128) >> 4), v_base + (uint32_t)((kbk * DH * 128) >> 4), idesc_o,
                          (j | kbk) ? 1u : 0u, false);
                at_commit(P_EMPTY);
                at_commit(V_EMPTY(sv));
                if (j == NKB - 1) at_commit(O_FULL);
            }
            __syncwarp();
            if (lane == 0) AT_TRACE(3, j);                              // slot 3: P V of block j issued
        }
    } else {
        // =========================== softmax + epilogue (warps 0-3) ===========================
        const int row = warp * 32 + lane;                          // query row of the tile = TMEM lane
        int myreg = 0;
        if (a.region) {
            const unsigned char* rg = a.region + (size_t)(z % a.n_win) * a.Lk;
            for (int i = threadIdx.x; i < a.Lk; i += 128) sReg[i] = rg[i];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            myreg = sReg[qt * 128 + row];
        }
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        float m = -INFINITY;
        // ---- pass A: running row maximum of the (scaled, masked) scores
        for (int t = 0; t < NKB; ++t) {
            const int s = t & 1;
            at_mbar_wait(S_FULL(s), (uint32_t)((t >> 1) & 1));
            at_tc_fence_after();
            // all four TMEM loads of the block in flight, ONE wait (each load waits ~1K cycles while the MMAs own the TMEM ports:
            // measured 5.8K cycles per block with a wait per 16 columns, profiles/r2_trace_attn_*.txt)
            uint32_t rr[AT_KB / 16][16];
#pragma unroll
            for (int c = 0; c < AT_KB / 16; ++c) at_tmem_ld16(t_lane + (uint32_t)(s * AT_KB + c * 16), rr[c]);
            at_tmem_ld_wait();
            at_tc_fence_before();
            at_mbar_arrive(S_EMPTY(s));                                 // (the scores are in registers now)
#pragma unroll
            for (int c = 0; c < AT_KB / 16; ++c) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float sc = __uint_as_float(rr[c][e]);
                    if (a.region && sReg[t * AT_KB + c * 16 + e] != myreg) sc += -100.0f;
                    m = fmaxf(m, sc);
                }
            }
            if (threadIdx.x == 0) AT_TRACE(4, t);                       // slot 4: pass A block t consumed
        }
        // ---- pass B: P = exp(s - m) as the (hi, lo) A operand of P V; row sum in fp32
        float l = 0.0f;
        const float m2 = m * 1.4426950408889634f;
        for (int j = 0; j < NKB; ++j) {
            const int t = NKB + j, s = t & 1;
            at_mbar_wait(S_FULL(s), (uint32_t)((t >> 1) & 1));
            at_tc_fence_after();
            // the whole block's probabilities are formed in registers first: reading S and the exponentials of block j overlap the
            // P V MMAs of block j-1, which still read the (single) P buffer
            uint4 ph[AT_KB / 8], pl[AT_KB / 8];
            {
                uint32_t rr[AT_KB / 16][16];
#pragma unroll
                for (int c = 0; c < AT_KB / 16; ++c) at_tmem_ld16(t_lane + (uint32_t)(s * AT_KB + c * 16), rr[c]);
                at_tmem_ld_wait();
                at_tc_fence_before();
                at_mbar_arrive(S_EMPTY(s));                        // S buffer free: the MMA warp may issue the scores of block j+2
#pragma unroll
                for (int c = 0; c < AT_KB / 16; ++c) {
                    float p[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float sc = __uint_as_float(rr[c][e]);
                        if (a.region && sReg[j * AT_KB + c * 16 + e] != myreg) sc += -100.0f;
                        float ex;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(sc, 1.4426950408889634f, -m2)));
                        p[e] = ex;
                        l += ex;
                    }
                    at_split8(p, ph[2 * c], pl[2 * c]);
                    at_split8(p + 8, ph[2 * c + 1], pl[2 * c + 1]);
                }
            }
            at_mbar_wait(P_EMPTY, (uint32_t)((j & 1) ^ 1));
#pragma unroll
            for (int c = 0; c < AT_KB / 16; ++c) {
                // keys c*16 .. c*16+15 of the block: sub-block c / 2, 16-byte chunks (c & 1) * 2 + {0, 1}; lo halves 4 chunks on
                uint8_t* rowp = sP + (c >> 1) * (128 * 128) + row * 128;
                const int ch = (c & 1) * 2;
                *reinterpret_cast<uint4*>(rowp + (((ch) ^ (row & 7)) << 4)) = ph[2 * c];
                *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (row & 7)) << 4)) = ph[2 * c + 1];
                *reinterpret_cast<uint4*>(rowp + (((ch + 4) ^ (row & 7)) << 4)) = pl[2 * c];
                *reinterpret_cast<uint4*>(rowp + (((ch + 5) ^ (row & 7)) << 4)) = pl[2 * c + 1];
            }
            at_fence_proxy_async();
            at_mbar_arrive(P_FULL);
            if (threadIdx.x == 0) AT_TRACE(5, j);                       // slot 5: P of block j handed over
        }
        // ---- epilogue: O / l -> global (token-major rows of DH floats)
        at_mbar_wait(O_FULL, 0);
        at_tc_fence_after();
        pdl_wait();                                                // (returns at once: the producers passed it long ago)
        const float inv = 1.0f / l;
        float* ob = a.out + (size_t)img * a.o_bs + (size_t)at_token(a, win, qt * 128 + row) * a.ldo + hoff;
#pragma unroll
        for (int c = 0; c < DH / 16; ++c) {
            uint32_t rr[16];
            at_tmem_ld16(t_lane + 128u + (uint32_t)(c * 16), rr);
            at_tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[e]) * inv;
            stg256(ob + c * 16, v);
            stg256(ob + c * 16 + 8, v + 8);
        }
    }
    at_tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) AT_TRACE(7, 2);
    if (warp == 4) {
        at_tc_fence_after();
        at_tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// K / V of every batch element -> the (hi, lo) fp16 stage images the attention kernel's MMAs read (same layout produce_K /
// produce_V build in shared memory), once per key block instead of once per (key block, query tile).  One block = one
// (batch z, key block j): the two images are formed in shared memory, then copied out with coalesced 16-byte stores.
template <int DH>
__global__ void __launch_bounds__(256) attn_pack_kv_kernel(const AttnTcArgs a, uint8_t* __restrict__ kpack, uint8_t* __restrict__ vpack) {
    using SM = AttnSmem<DH>;
    constexpr int NCB = SM::NCB;
    extern __shared__ __align__(16) uint8_t pk_smem[];
    uint8_t* pK = pk_smem;
    uint8_t* pV = pk_smem + SM::K_STAGE;
    pdl_wait();
    if (threadIdx.x == 0) keep_stamp_here();
    const int NKB = a.Lk / AT_KB;
    const int z = blockIdx.x / NKB, j = blockIdx.x - z * NKB;
    const int nwin = a.win_side > 0 ? a.win_side * a.win_side : 1;
    const int img = a.win_side > 0 ? z / nwin : (a.heads > 1 ? z / a.heads : z);
    const int win = a.win_side > 0 ? z - img * nwin : 0;
    const int hoff = a.heads > 1 ? (z - img * a.heads) * DH : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* kb = a.k + (size_t)img * a.k_bs + hoff;
    const float* vb = a.v + (size_t)img * a.v_bs + hoff;
    constexpr int UPR = NCB * 4;
    constexpr int NU = (AT_KB * UPR) / 256;                    // K units per thread (4 at DH = 128)
    constexpr int NCOMBO = 2 * (DH / 8), NC = NCOMBO / 8;      // V combos per warp (4 at DH = 128)
    float kv[NU][8], vv[NC][8];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int u = threadIdx.x + i * 256;
        const int row = u / UPR, cu = u - row * UPR;
        ldg256(kb + (size_t)at_token(a, win, j * AT_KB + row) * a.ldk + (cu >> 2) * 32 + (cu & 3) * 8, kv[i]);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = warp + i * 8, kbk = c / (DH / 8), dg = c - kbk * (DH / 8);
        ldg256(vb + (size_t)at_token(a, win, j * AT_KB + kbk * 32 + lane) * a.ldv + dg * 8, vv[i]);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        const int u = threadIdx.x + i * 256;
        const int row = u / UPR, cu = u - row * UPR, cb = cu >> 2, pl = cu & 3;
        uint4 hi, lo;
        at_split8(kv[i], hi, lo);
        const int off = cb * (AT_KB * 128) + row * 128 + ((pl ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(pK + off) = hi;
        *reinterpret_cast<uint4*>(pK + (off ^ 64)) = lo;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = warp + i * 8, kbk = c / (DH / 8), dg = c - kbk * (DH / 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int row = dg * 8 + e;
            const __half h = __float2half_rn(vv[i][e]);
            const __half l = __float2half_rn(vv[i][e] - __half2float(h));
            uint8_t* rowp = pV + kbk * (DH * 128) + row * 128;
            *reinterpret_cast<__half*>(rowp + ((((lane >> 3)) ^ (row & 7)) << 4) + (lane & 7) * 2) = h;
            *reinterpret_cast<__half*>(rowp + ((((lane >> 3) + 4) ^ (row & 7)) << 4) + (lane & 7) * 2) = l;
        }
    }
    __syncthreads();
    uint4* kd = reinterpret_cast<uint4*>(kpack + ((size_t)z * NKB + j) * SM::K_STAGE);
    uint4* vd = reinterpret_cast<uint4*>(vpack + ((size_t)z * NKB + j) * SM::V_STAGE);
    for (int i = threadIdx.x; i < SM::K_STAGE / 16; i += 256) kd[i] = reinterpret_cast<const uint4*>(pK)[i];
    for (int i = threadIdx.x; i < SM::V_STAGE / 16; i += 256) vd[i] = reinterpret_cast<const uint4*>(pV)[i];
}

}  // namespace

long long* g_attn_trace = nullptr;   // debug: device buffer of 8 x 40 clock64 stamps (keepop_attn_trace)

bool attention_tc_eligible(int Lq, int Lk, int dh) { return (dh == 128 || dh == 64) && Lq % 128 == 0 && Lk % AT_KB == 0 && Lk <= 1024 && Lq > 0 && Lk > 0; }

size_t attention_tc_pack_bytes(int nb, int Lk, int dh) { return (size_t)2 * nb * Lk * dh * 4; }

void attention_tc_configure_device() {
    static unsigned long long configured = 0;
    if (first_use_on_current_device(&configured)) {
        CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<128>::TOTAL));
        CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<64>::TOTAL));
        CUDA_CHECK(cudaFuncSetAttribute(attn_pack_kv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<128>::K_STAGE + AttnSmem<128>::V_STAGE));
        CUDA_CHECK(cudaFuncSetAttribute(attn_pack_kv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<64>::K_STAGE + AttnSmem<64>::V_STAGE));
    }
}

void attention_tc(const float* q, int ldq, long long q_bs, const float* k, int ldk, long long k_bs, const float* v, int ldv, long long v_bs,
                  float* out, int ldo, long long o_bs, int nb, int Lq, int Lk, int dh, float scale, const unsigned char* region, int n_win,
                  cudaStream_t s, int win_side, int wsz, int map_w, int shift, int heads, void* kv_pack) {
    KEEP_CHECK(attention_tc_eligible(Lq, Lk, dh), "attention_tc: unsupported shape (Lq %d, Lk %d, dh %d)", Lq, Lk, dh);
    KEEP_CHECK(!region || (Lq == Lk && n_win > 0), "attention_tc: the region mask needs Lq == Lk");
    KEEP_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 && o_bs % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 31) == 0,
               "attention_tc: operands must be 32-byte aligned with strides that are multiples of 8 floats");
    static_assert(AttnSmem<128>::TOTAL <= 227 * 1024, "attention kernel exceeds the shared-memory budget");
    attention_tc_configure_device();
    AttnTcArgs a;
    a.q = q; a.k = k; a.v = v; a.out = out;
    a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.o_bs = o_bs;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
    a.nb = nb; a.Lq = Lq; a.Lk = Lk; a.scale = scale; a.region = region; a.n_win = n_win > 0 ? n_win : 1;
    a.trace = g_attn_trace;
    a.win_side = win_side; a.wsz_log2 = 0; a.map_w = map_w; a.shift = shift; a.heads = heads > 1 ? heads : 1;
    KEEP_CHECK(a.heads == 1 || (win_side == 0 && nb % a.heads == 0), "attention_tc: multi-head mode needs plain rows and nb = batches * heads");
    if (win_side > 0) {
        KEEP_CHECK(wsz > 0 && (wsz & (wsz - 1)) == 0 && (map_w & (map_w - 1)) == 0 && win_side * wsz == map_w && Lq == wsz * wsz && Lk == Lq &&
                       nb % (win_side * win_side) == 0, "attention_tc: bad window geometry");
        while ((1 << a.wsz_log2) < wsz) ++a.wsz_log2;
    }
    a.kpack = a.vpack = nullptr;
    if (kv_pack) {   // convert K / V once per key block (attention_tc_pack_bytes(nb, Lk, dh) bytes of scratch), then stream the stages by TMA
        KEEP_CHECK((reinterpret_cast<uintptr_t>(kv_pack) & 127) == 0, "attention_tc: the K / V pack scratch must be 128-byte aligned");
        uint8_t* kp = (uint8_t*)kv_pack;
        uint8_t* vp = kp + (size_t)nb * Lk * dh * 4;
        const unsigned blocks = (unsigned)(nb * (Lk / AT_KB));
        if (dh == 128) launch_k(attn_pack_kv_kernel<128>, dim3(blocks), dim3(256), (size_t)(AttnSmem<128>::K_STAGE + AttnSmem<128>::V_STAGE), s, a, kp, vp);
        else launch_k(attn_pack_kv_kernel<64>, dim3(blocks), dim3(256), (size_t)(AttnSmem<64>::K_STAGE + AttnSmem<64>::V_STAGE), s, a, kp, vp);
        CUDA_CHECK(cudaGetLastError());
        a.kpack = kp; a.vpack = vp;
    }
    if (dh == 128) launch_k(attn_tc_kernel<128>, dim3((unsigned)(nb * (Lq / 128))), dim3(AT_THREADS), (size_t)AttnSmem<128>::TOTAL, s, a);
    else launch_k(attn_tc_kernel<64>, dim3((unsigned)(nb * (Lq / 128))), dim3(AT_THREADS), (size_t)AttnSmem<64>::TOTAL, s, a);
    CUDA_CHECK(cudaGetLastError());
}

KEEP_STAMP_SETTER(stamp_set_attn)

}  // namespace keep
