// keep_b200 — tcgen05 implicit-GEMM convolution / linear for sm_100a (NHWC, fp16 operands, fp32 accumulate in TMEM).
//
// One persistent, warp-specialised kernel covers every 3x3 stride-1 convolution (optionally with the
// generator's nearest-x2 upsample and the CFT channel concat folded into the gather) and every 1x1
// convolution / nn.Linear on the KEEP path whose Cout is a multiple of 16:
//
//   warps 0-3   epilogue      tcgen05.ld TMEM -> registers -> {bias, activation, +residual, cast} -> HBM
//   warp  4     MMA issuer    one thread issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) and tcgen05.commit
//   warp  8     weight loader one thread streams pre-packed fp16 weight panels with 1-D TMA (cp.async.bulk), or
//                             parks the layer's whole panel set in shared memory once when it fits
//   14 warps    A producers   (1x1 layers: two groups of 7 on alternate stages) coalesced NHWC loads of the (16+2)x(8+2) input halo,
//                             GroupNorm/InstanceNorm apply + swish/ReLU in registers, fp16 (hi, lo) split,
//                             st.shared into the UMMA SWIZZLE_128B K-major layout
//
// Implicit GEMM without im2col traffic: the halo tile of 64 input channels is staged ONCE in shared
// memory; the nine filter taps are nine *shifted views* of it, expressed purely through the UMMA
// shared-memory descriptor (start address + SBO = one halo row), so each activation byte is read from
// L2/HBM once per CTA tile instead of nine times.  Accumulators are double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Small-M layers (16^2..64^2 maps of the serial
// per-frame chain) are split over K (channel blocks) across CTAs with a deterministic fixed-order reduce.
//
// Replaces the cuDNN / cuBLAS dispatch behind nn.Conv2d / nn.Linear in vqgan_arch.py:161-181,219-243,
// 260-286,311-335 and keep_arch.py:78-87,391-393,445-455,929,936-938.
#include <vector>

#ifndef KEEP_TC_HPITCH
#define KEEP_TC_HPITCH 10
#endif
#ifndef KEEP_PDL_CONV_TRIGGER
#define KEEP_PDL_CONV_TRIGGER 0
#endif
#ifndef KEEP_TC_PARTIAL32
#define KEEP_TC_PARTIAL32 1   // split-K partial epilogue: 32 columns per TMEM round trip
#endif
#ifndef KEEP_TC_GROUPS_3X3
#define KEEP_TC_GROUPS_3X3 1   // producer groups on 3x3 layers (split-precision mode): 2 = alternate stages, 1 = all warps on one stage
#endif
#ifndef KEEP_TC_STACKED
#define KEEP_TC_STACKED 1   // stacked [Wh ; Wl] weight panels for 64-wide N tiles in the split-precision mode (0: three N = 64 MMAs per K step)
#endif

#include "ops.h"
#include "tc.h"

namespace keep {
namespace {

// Warp roles are laid out by SM sub-partition (warp % 4 picks the scheduler).  The MMA issuer needs ~90 instructions per
// filter tap; sharing a scheduler with four busy producer warps stretched that to 350-550 cycles per tap, more than the
// MMAs themselves take (measured: 88-105 clk per N=64 MMA in the kernel vs 55-64 in isolation, tools/ubench/umma_rate.cu).
// So sub-partition 0 holds the MMA issuer, the weight loader, one epilogue warp and only two producer warps (12, 16); the other
// twelve producer warps live on sub-partitions 1-3 next to the other three epilogue warps (tcgen05.ld ties epilogue warp w to
// TMEM lanes 32*(w%4)...).
// 640 threads = 20 warps: the register file gives 65 536 / 640 -> 96 registers per thread.  With 22 warps (704 threads, 16
// producer warps) ptxas is capped at 80 and every variant spilled a little; 14 producer warps without spills are faster
// (measured 170.9 -> 175.8 frames/s; 64->64 3x3 @512^2 tile 6.9K -> 6.3K cycles; -DKEEP_TC_THREADS=704 restores the old shape)
#ifndef KEEP_TC_THREADS
#define KEEP_TC_THREADS 640
#endif
#ifndef KEEP_TC_THREADS_1X1
#define KEEP_TC_THREADS_1X1 KEEP_TC_THREADS            // (1x1 / linear variants can be built with their own CTA size: A/B)
#endif
constexpr int tc_threads(int win) { return win == 1 ? KEEP_TC_THREADS_1X1 : KEEP_TC_THREADS; }
constexpr int kEpiWarps = 4, kMmaWarp = 4, kLoadWarp = 8;
constexpr int MAX_SA = 8, MAX_SB = 8;  // barrier slots (actual pipeline depths come from the launch arguments)
// channels per A stage: 64 with fp16 operands (4 MMA K-steps of 16); 32 in the split-precision mode, whose 128-byte rows
// hold [hi 32 ch | lo 32 ch] side by side (2 K-steps each), so a stage and a weight panel have the same geometry in both
__host__ __device__ constexpr int cb_of(int passes) { return passes == 3 ? 32 : 64; }
// A stage = halo tile in the UMMA SWIZZLE_128B K-major layout: one 128-byte row (64 channels, fp16) per halo pixel,
// 16-byte chunks XOR-swizzled by (row & 7).  The halo is 18 rows x 10 columns but rows are PITCHED at 16 pixels, so
// that (a) every 8-pixel group of an MMA operand view starts SBO = 16*128 = 2048 B after the previous one (a multiple
// of the 1024-byte swizzle atom: all groups share one swizzle phase) and (b) a filter tap (dy, dx) is just the start
// address base + (dy*16 + dx)*128 with descriptor base_offset = dx.  (The first version used the no-swizzle
// "interleaved" core-matrix layout: correct, but the tensor core fetched it 16 bytes per cycle, ~300 cycles per MMA.)
constexpr int HPITCH_PX = KEEP_TC_HPITCH;
constexpr int A_SUB_BYTES = (18 * HPITCH_PX * 128 + 1023) / 1024 * 1024;   // one A stage

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B, K-major (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30)
// (unused for swizzled K-major: 1), SBO>>4 [32,46) = distance between 8-row groups, version 1 [46,48),
// base_offset [49,52) = (start >> 7) & 7 when the start is not 1024-byte aligned, layout type 2 = SWIZZLE_128B [61,64).

// The kernel's code footprint matters: five warp roles run different code at once and the SM's instruction cache
// is small (a first version that inlined the generic activation switch 48x ran 10x slower, stalled on fetch).
// Producer prologue activations are only ever swish (VQGAN) or ReLU (GMFlow); everything else is out of line.
template <bool EXACT>
__device__ __forceinline__ float swish_f(float v) {
    if (EXACT) {   // v * rcp(1 + ex2(-v log2 e)): 5 instructions, <= 3 ulp (__expf/__fdividef add range fix-ups: 8+; an IEEE divide ~20)
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        return v * r;
    }
    float t;   // x * sigmoid(x) = 0.5 x (1 + tanh(x/2)): one MUFU op
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
    return 0.5f * v * (1.0f + t);
}
__device__ __noinline__ float act_slow(float v, int act) { return apply_act(v, act); }

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// PASSES = 1: fp16 operands (10-bit mantissa, the precision class of cuDNN's default TF32 convolutions).
// PASSES = 3: split-precision: A = Ah + Al, W = Wh + Wl (fp16 pairs, ~22 mantissa bits), D += Ah*Wh + Ah*Wl + Al*Wh
//             with fp32 accumulation in TMEM -> fp32-grade results on the tensor cores (the "decision path" needs
//             this: one flipped argmax over the 1024 code logits changes a 32x32-pixel block).
// optional per-role timeline (debug): a.trace != null -> CTA 0 stamps clock64() at role milestones of its first tiles
#ifndef KEEP_TC_TRACE_FINE
#define KEEP_TC_TRACE_FINE 0   // 1 (debug builds): extra producer milestones in rows 10-13 of a 256-entry trace buffer
#endif
#define TC_TRACE_FINE(slot, idx)                                                                   \
    do {                                                                                           \
        if (KEEP_TC_TRACE_FINE) TC_TRACE(slot, idx);                                               \
    } while (0)
#define TC_TRACE(slot, idx)                                                                        \
    do {                                                                                           \
        if (a.trace && blockIdx.x == 0 && (idx) < 16) a.trace[(slot) * 16 + (idx)] = clock64();    \
    } while (0)

// A_BF16 (split-precision mode only): both operand pairs (Ah, Al), (Wh, Wl) are bf16 instead of fp16 -- fp32's exponent
// range, 16 mantissa bits per pair instead of 22.  For layers whose input is a raw (un-normalised) feature map of
// unbounded magnitude: fp16 overflows to inf beyond 65504 (KEEP_FLAG_TC_WIDE).
template <int PASSES, bool IN_F16, int WIN, bool A_BF16 = false>
__global__ void __launch_bounds__(tc_threads(WIN), 1) conv_tc_kernel(const TcConvArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms are 1024-byte aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CBK = cb_of(PASSES);                   // channels per A stage / weight panel
    constexpr int UPP = CBK / 8;                         // 8-channel producer units per pixel
    // Producer groups: in the split-precision mode the 14 producer warps work as TWO groups of 7 that take alternate stages.
    // A stage's critical path is load latency + convert + proxy fence (the fence -- MEMBAR.ALL.CTA -- also waits for any
    // prefetched global load of the same thread, which is why register double-buffering inside one thread bought nothing);
    // with two groups one stage's latency hides behind the other's conversion.
    constexpr int NGROUPS = (PASSES == 3 && (WIN == 1 || KEEP_TC_GROUPS_3X3 == 2)) ? 2 : 1;   // (3x3 layers: one group -- MAXIT 3 at 83 % fill and spills cost more than the overlap gives)
    constexpr int kProdThreads = tc_threads(WIN) - 6 * 32;   // producers: the warps >= 5 other than the loader (8)
    constexpr int GT = kProdThreads / NGROUPS;           // threads per group = arrivals per stage
    constexpr int PPI = GT / UPP;                        // pixels per producer iteration
    constexpr int NPIX = WIN == 3 ? 18 * 10 : (WIN == 2 ? 17 * 9 : 128);   // pixels per stage (halo tile, or the 128 of a 1x1)
    constexpr int MAXIT = (NPIX + PPI - 1) / PPI;        // units per producer thread and stage
    constexpr int KSTEPS = CBK / 16;                     // MMA K-steps per operand tile
    constexpr int A_STAGE_BYTES = WIN == 1 ? 128 * 128 : A_SUB_BYTES;   // 1x1: 128 pixels, no halo
    const int SA = a.sa_stages, SB = a.sb_stages;
    const int b_stage_bytes = a.bn * 128;                // one weight panel: bn rows x 128 bytes
    uint8_t* sA = smem;
    uint8_t* sB = smem + SA * A_STAGE_BYTES;
    // weight region: SB streaming stages, or (a.w_resident) every panel of the layer, loaded once per CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (a.w_resident ? a.ncb * WIN * WIN : SB) * b_stage_bytes);
    // barrier map: a_full[MAX_SA] a_empty[MAX_SA] b_full[MAX_SB] b_empty[MAX_SB] acc_full[2] acc_empty[2]
    const uint32_t bar0 = smem_u32(bars);
    auto A_FULL = [&](int s) { return bar0 + 8u * s; };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (MAX_SA + s); };
    auto B_FULL = [&](int s) { return bar0 + 8u * (2 * MAX_SA + s); };
    auto B_EMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_SA + MAX_SB + s); };
    auto ACC_FULL = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + s); };
    auto ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_SA + 2 * MAX_SB + 4);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);          // [256] bias of the current N tile

    pdl_early_trigger();
    if (threadIdx.x == 0) TC_TRACE(9, 0);                 // kernel entry
    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_SA; ++s) { mbar_init(A_FULL(s), GT); mbar_init(A_EMPTY(s), 1); }
        for (int s = 0; s < MAX_SB; ++s) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(ACC_FULL(s), 1); mbar_init(ACC_EMPTY(s), kEpiWarps * 32); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_TRACE(9, 1);                 // barriers + TMEM ready
    // griddepcontrol.wait is executed per role, right before its first access to memory the previous kernel may still be
    // writing (activations, residual, outputs): everything above and the roles' index setup -- and the weight loader's first
    // TMA transfers, weights being constants -- overlap the previous kernel when it triggered early (pdl_prologue_light).

    // WIN = 3: 3x3 s1 | 2: 3x3 stride-2 as a 2x2 window over the virtual space-to-depth input | 1: 1x1
    constexpr int win = WIN, TAPS = WIN * WIN;
    constexpr bool conv3 = WIN > 1;        // halo-tile modes
    constexpr int hcols = 8 + WIN - 1;     // halo columns (rows are pitched at HPITCH_PX pixels)
    // stride-2 mode: virtual channel block cb = parity p * ncbr + real block; tap (a,b) x parity (py,px) maps to filter
    // position ky = 2a + py - pad_t, kx = 2b + px - pad_l; combinations outside the 3x3 filter are skipped entirely
    auto stage_live = [&](int cb, int tap) -> bool {
        if (!a.s2d) return true;
        const int p = cb / a.ncbr;
        const int ky = 2 * (tap >> 1) + (p >> 1) - a.pad_t, kx = 2 * (tap & 1) + (p & 1) - a.pad_l;
        return ky >= 0 && ky <= 2 && kx >= 0 && kx <= 2;
    };
    const int mt_per_img = a.tiles_y * a.tiles_x;
    const int m_tiles = a.n * mt_per_img;
    // Work items.  Default: item = (n tile, m tile, k split), n tile fastest, items strided over the CTAs.
    // A-stationary (a.a_stat; 1x1 layers with several N tiles and many M tiles): item = (m tile, k split); the CTA walks
    // all N tiles of the item back to back against ONE production of the activation stages, which stay in shared memory
    // until the last N tile has consumed them (an N = 1024 linear otherwise converts the same activations eight times).
    const int nt_inner = a.a_stat ? a.ntile_n : 1;
    const int total = m_tiles * a.splitk * (a.a_stat ? 1 : a.ntile_n);       // (host checks < 2^31)
    const int cb_per = a.cb_per;

    auto decode = [&](int w, int nti, int& nt, int& img, int& ty, int& tx, int& ks) {
        uint32_t r = (uint32_t)w;
        int mt;
        if (a.cluster_k) {   // cluster split-K: the K splits of one output tile are the CTAs of one cluster (k split fastest)
            const uint32_t q = fdiv(r, a.fd_splitk);
            ks = (int)(r - q * a.fd_splitk.d); r = q;
            const uint32_t q2 = fdiv(r, a.fd_ntile);
            nt = (int)(r - q2 * a.fd_ntile.d);
            mt = (int)q2;
        } else {
            if (a.a_stat) nt = nti;
            else { const uint32_t q = fdiv(r, a.fd_ntile); nt = (int)(r - q * a.fd_ntile.d); r = q; }
            const uint32_t q2 = fdiv(r, a.fd_mtiles);
            mt = (int)(r - q2 * a.fd_mtiles.d);
            ks = (int)q2;
        }
        img = (int)fdiv((uint32_t)mt, a.fd_mtimg);
        const int t2 = mt - img * mt_per_img;
        ty = (int)fdiv((uint32_t)t2, a.fd_tilesx);
        tx = t2 - ty * a.tiles_x;
    };

    if (warp > 4 && warp != kLoadWarp) {
        // =========================== A producers ===========================
        const int pw = warp - (warp > kLoadWarp ? 6 : 5);                           // 0..15
        const int grp = NGROUPS == 2 ? (pw & 1) : 0;                                // this warp's group
        const int pt = (NGROUPS == 2 ? (pw >> 1) : pw) * 32 + lane;                 // thread index within the group
        const int pl = pt % UPP;                            // 8-channel plane of the stage handled by this thread
        const int Hl = a.h * a.up, Wl = a.w * a.up;
        constexpr int npix = NPIX;
        const int cin = a.c0 + a.c1;
        const int ushift = a.up - 1;
        // halo units of this thread: pixel p = p_first + it*PPI -> (row, col) within the halo; tile independent, and so
        // are the shared-memory offsets of its 16-byte chunks (row-XOR swizzle; the lo half sits 4 chunks after the hi half)
        const int p_first = pt / UPP;
        int hy[MAXIT], hx[MAXIT], soff[MAXIT];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int p = p_first + it * PPI;
            hy[it] = p / hcols;
            hx[it] = p - hy[it] * hcols;
            const int row = conv3 ? hy[it] * HPITCH_PX + hx[it] : p;
            soff[it] = p < npix ? row * 128 + ((pl ^ (row & 7)) << 4) : -1;
        }
        int stage = 0, phase = 0, trace_i = 0, seq = 0;
        bool waited = false;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            if (pt == 0 && grp == 0) TC_TRACE_FINE(10, trace_i);
            int nt, img, ty, tx, ks;
            decode(w, 0, nt, img, ty, tx, ks);
            const int cb0 = ks * cb_per, cb1 = min(a.ncb, cb0 + cb_per);
            const int oy0 = ty * 16 - (a.s2d ? a.pad_t : 1), ox0 = tx * 8 - (a.s2d ? a.pad_l : 1);
            // pixel index of every halo unit (-1: outside the image / the tile): once per tile, except in the stride-2 mode
            // where it depends on the parity of the virtual channel block
            int pixv[MAXIT];                // (host checks n*h*w < 2^31)
            auto locate = [&](int py, int px) {
#pragma unroll
                for (int it = 0; it < MAXIT; ++it) {
                    bool ok = soff[it] >= 0;
                    int pix;
                    if (conv3) {
                        int iy = oy0 + hy[it], ix = ox0 + hx[it];
                        if (a.s2d) { iy = 2 * iy + py; ix = 2 * ix + px; }
                        ok = ok && iy >= 0 && iy < Hl && ix >= 0 && ix < Wl;
                        pix = (img * a.h + (iy >> ushift)) * a.w + (ix >> ushift);
                    } else {
                        const int q = ty * 128 + p_first + it * PPI;     // pixel index within the image (host checks n*h*w < 2^31)
                        ok = ok && q < a.h * a.w;
                        pix = img * a.h * a.w + q;
                    }
                    pixv[it] = ok ? pix : -1;
                }
            };
            if (!a.s2d) locate(0, 0);
            if (pt == 0 && grp == 0) TC_TRACE_FINE(11, trace_i);
            for (int cb = cb0; cb < cb1; ++cb) {
                if (NGROUPS == 2 && ((seq++ & 1) != grp)) {   // the other group's stage: only keep the ring position in step
                    if (++stage == SA) { stage = 0; phase ^= 1; }
                    continue;
                }
                const int par = a.s2d ? cb / a.ncbr : 0;    // stride-2 mode: input parity (py, px) of this virtual block
                if (a.s2d) locate(par >> 1, par & 1);
                const int ch = (a.s2d ? (cb - par * a.ncbr) : cb) * CBK + pl * 8;   // first of this thread's 8 real channels
                const bool ch_ok = ch < cin;
                const uint8_t* src; int sc_ch, cc;
                if (ch < a.c0) { src = reinterpret_cast<const uint8_t*>(a.in0); sc_ch = a.ld0; cc = ch; }
                else { src = reinterpret_cast<const uint8_t*>(a.in1); sc_ch = a.c1; cc = ch - a.c0; }
                constexpr int ESZ = IN_F16 ? 2 : 4;
                if (!waited) {   // first activation load of this thread: the previous kernel must have completed (PDL)
                    pdl_wait();
                    waited = true;
                    if (pt == 0 && grp == 0) { TC_TRACE(9, 2); keep_stamp_here(); }
                }
                // ---- issue every global load of this stage first (memory-level parallelism), then transform
                uint4 raw[MAXIT][IN_F16 ? 1 : 2];
#pragma unroll
                for (int it = 0; it < MAXIT; ++it) {
                    if (ch_ok && pixv[it] >= 0) {
                        const uint8_t* g = src + ((size_t)pixv[it] * sc_ch + cc) * ESZ;
                        if (IN_F16) raw[it][0] = *reinterpret_cast<const uint4*>(g);
                        else ldg256(reinterpret_cast<const float*>(g), reinterpret_cast<float*>(&raw[it][0]));
                    }
                }
                if (pt == 0 && grp == 0) TC_TRACE_FINE(12, trace_i);
                float sc[8], sh[8];
                if (a.pre_scale && ch_ok) {
                    const float4* ps = reinterpret_cast<const float4*>(a.pre_scale + (size_t)img * cin + ch);
                    const float4* pb = reinterpret_cast<const float4*>(a.pre_shift + (size_t)img * cin + ch);
                    float4 s0 = ps[0], s1 = ps[1], b0 = pb[0], b1 = pb[1];
                    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
                }
                if (pt == 0 && grp == 0) TC_TRACE(0, trace_i);
                mbar_wait(A_EMPTY(stage), phase ^ 1);
                if (pt == 0 && grp == 0) TC_TRACE(1, trace_i);
                uint8_t* dst = sA + stage * A_STAGE_BYTES;
#pragma unroll
                for (int it = 0; it < MAXIT; ++it) {
                    if (soff[it] < 0) continue;
                    uint4 o = make_uint4(0u, 0u, 0u, 0u), ol = make_uint4(0u, 0u, 0u, 0u);
                    if (ch_ok && pixv[it] >= 0) {
                        float v[8];
                        if (!IN_F16) {
                            const float* f0 = reinterpret_cast<const float*>(&raw[it][0]);
                            const float* f1 = reinterpret_cast<const float*>(&raw[it][IN_F16 ? 0 : 1]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { v[j] = f0[j]; v[4 + j] = f1[j]; }
                        } else {
                            const __half2* hh = reinterpret_cast<const __half2*>(&raw[it][0]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hh[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
                        }
                        if (a.pre_scale) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
                        }
                        __half2 hq[4];
                        if (PASSES == 3) {
                            if (a.pre_act == ACT_SWISH) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = swish_f<true>(v[j]);
                            } else if (a.pre_act == ACT_RELU) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                            }
                            __half2 lq[4];
                            if (A_BF16) {   // same split with bf16 parts (bit patterns carried in the __half2 registers)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                                    const float2 r = __bfloat1622float2(hb);
                                    const __nv_bfloat162 lb = __floats2bfloat162_rn(v[2 * j] - r.x, v[2 * j + 1] - r.y);
                                    hq[j] = *reinterpret_cast<const __half2*>(&hb);
                                    lq[j] = *reinterpret_cast<const __half2*>(&lb);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) hq[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {   // residual (lo) operand: v - float(fp16(v))
                                    const float2 r = __half22float2(hq[j]);
                                    lq[j] = __floats2half2_rn(v[2 * j] - r.x, v[2 * j + 1] - r.y);
                                }
                            }
                            ol.x = *reinterpret_cast<uint32_t*>(&lq[0]); ol.y = *reinterpret_cast<uint32_t*>(&lq[1]);
                            ol.z = *reinterpret_cast<uint32_t*>(&lq[2]); ol.w = *reinterpret_cast<uint32_t*>(&lq[3]);
                        } else if (a.pre_exact) {       // fp16 operands, but the activation still in fp32 (hybrid precision)
                            if (a.pre_act == ACT_SWISH) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = swish_f<true>(v[j]);
                            } else if (a.pre_act == ACT_RELU) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) hq[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) hq[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                            if (a.pre_act == ACT_SWISH) {   // packed: x*sigmoid(x) = h + h*tanh(h), h = x/2 (one MUFU per 2 elements)
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const __half2 hh = __hmul2(hq[j], __float2half2_rn(0.5f));
                                    uint32_t t, hi = *reinterpret_cast<const uint32_t*>(&hh);
                                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hi));
                                    hq[j] = __hfma2(hh, *reinterpret_cast<__half2*>(&t), hh);
                                }
                            } else if (a.pre_act == ACT_RELU) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) hq[j] = __hmax2(hq[j], __float2half2_rn(0.0f));
                            }
                        }
                        o.x = *reinterpret_cast<uint32_t*>(&hq[0]); o.y = *reinterpret_cast<uint32_t*>(&hq[1]);
                        o.z = *reinterpret_cast<uint32_t*>(&hq[2]); o.w = *reinterpret_cast<uint32_t*>(&hq[3]);
                    }
                    // 16-byte chunk `pl` of the row lands at chunk (pl ^ (row & 7)); the lo half 4 chunks (64 bytes) further
                    *reinterpret_cast<uint4*>(dst + soff[it]) = o;
                    if (PASSES == 3) *reinterpret_cast<uint4*>(dst + (soff[it] ^ 64)) = ol;
                }
                if (pt == 0 && grp == 0) TC_TRACE_FINE(13, trace_i);
                fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor-core (async) proxy
                mbar_arrive(A_FULL(stage));
                if (pt == 0 && grp == 0) { TC_TRACE(2, trace_i); ++trace_i; }
                if (++stage == SA) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kLoadWarp) {
        // =========================== weight loader (1-D TMA) ===========================
        // real weights were packed at engine creation: their first panels are fetched while the previous kernel still runs
        if (!a.wt_static) pdl_wait();
        if (lane == 0 && a.w_resident) {
            // the layer's whole panel set (one N tile, no K split) fits next to the activation stages: fetch it once
            // per CTA instead of once per 128-pixel tile (which costs 64 B/clk/SM of L2 bandwidth at the MMA floor)
            if ((int)blockIdx.x < total) {
                const int npanels = a.ncb * TAPS;
                mbar_arrive_expect_tx(B_FULL(0), (uint32_t)(npanels * b_stage_bytes));
                const uint8_t* g = reinterpret_cast<const uint8_t*>(a.wt);
                for (int i = 0; i < npanels; ++i)
                    tma_bulk_g2s(smem_u32(sB + i * b_stage_bytes), g + (size_t)i * b_stage_bytes, (uint32_t)b_stage_bytes, B_FULL(0));
            }
        } else if (lane == 0) {
            int stage = 0, phase = 0, trace_l = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x)
            for (int nti = 0; nti < nt_inner; ++nti) {
                int nt, img, ty, tx, ks;
                decode(w, nti, nt, img, ty, tx, ks);
                const int cb0 = ks * cb_per, cb1 = min(a.ncb, cb0 + cb_per);
                for (int cb = cb0; cb < cb1; ++cb) {
                    for (int tap = 0; tap < TAPS; ++tap) {
                        if (!stage_live(cb, tap)) continue;
                        mbar_wait(B_EMPTY(stage), phase ^ 1);
                        mbar_arrive_expect_tx(B_FULL(stage), (uint32_t)b_stage_bytes);
                        const uint8_t* g = reinterpret_cast<const uint8_t*>(a.wt) + (size_t)img * a.wt_img_stride * sizeof(__half) +
                                           ((size_t)((size_t)nt * a.ncb + cb) * TAPS + tap) * (size_t)b_stage_bytes;
                        tma_bulk_g2s(smem_u32(sB + stage * b_stage_bytes), g, (uint32_t)b_stage_bytes, B_FULL(stage));
                        if (cb == cb0 && tap == 0) TC_TRACE(8, trace_l);
                        if (++stage == SB) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =========================== MMA issuer ===========================
        // The whole warp runs the (warp-uniform) control flow so that stage counters, shared-memory addresses and
        // descriptors live in the uniform datapath; one elected lane issues the tcgen05 instructions.  A first version
        // ran this loop on a single divergent lane with per-MMA 64-bit descriptor construction: ~150 instructions per
        // filter tap on one thread made *instruction issue of this warp* the bottleneck of the whole kernel.
        // f16 x f16 -> f32, K-major A/B; A_BF16: a_format (bits [7,10)) = b_format (bits [10,13)) = 1 = bf16 (kind::f16 traps
        // on mixed f16 / bf16 operands, so the wide variant's weight panels hold bf16 pairs too)
        const uint32_t idesc = (1u << 4) | (A_BF16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(a.bn >> 3) << 17) | ((128u >> 4) << 24);
        // descriptor = {lo: start>>4 | LBO(1)<<16, hi: SBO>>4 | version 1<<14 | base_offset<<17 | SWIZZLE_128B(2)<<29}
        constexpr uint32_t a_sbo = conv3 ? (uint32_t)(HPITCH_PX * 128) : 1024u;
        constexpr uint32_t a_hi = ((a_sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        constexpr uint32_t b_hi = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        constexpr uint32_t lo0 = 1u << 16;                                      // LBO field (unused for swizzled K-major)
        constexpr uint32_t kstep = 32u >> 4;                                    // K=16 halfs = 32 bytes inside the 128-byte row
        constexpr uint32_t lo_off = 64u >> 4;                                   // split precision: the lo half of a row
        // Measured on B200: with a 128-byte-row-shifted start address the swizzle XOR is applied on the absolute shared
        // memory address bits, so base_offset stays 0 (setting it to dx double-counts the phase and corrupts the tile).
        // The tap loop is unrolled at compile time: every descriptor is "stage base + constant".  (With run-time window
        // sizes the issuing warp executed ~90 mostly dependent instructions per tap -- 500+ cycles of single-warp latency
        // against 200-400 cycles of MMA work -- and the tensor pipe idled: profiles/r1_trace_conv64_*.txt.)
        const bool leader = elect_one();
        const uint32_t b_step = (uint32_t)b_stage_bytes >> 4;
        const uint32_t b_base = lo0 | (smem_u32(sB) >> 4);
        // stacked panels (split precision, 64-wide N tile): B = [Wh ; Wl] as 128 rows of 64 bytes, SWIZZLE_64B (layout type 4),
        // 8-row groups 512 bytes apart.  Per K step: D[:, 0:128] += Ah * [Wh ; Wl]^T (N = 128), D[:, 0:64] += Al * Wh^T (N = 64);
        // the epilogue adds columns j and 64 + j.  Two A fetches instead of three on the operand-fetch-bound N = 64 layers.
        const bool stacked = KEEP_TC_STACKED && PASSES == 3 && a.bn == 64;
        constexpr uint32_t b_hi_st = ((512u >> 4) & 0x3FFF) | (1u << 14) | (4u << 29);
        const uint32_t idesc_n128 = (idesc & ~(0x3Fu << 17)) | ((128u >> 3) << 17);
        // the MMAs of one (channel block, tap): K-steps x {lo*hi, hi*lo, hi*hi}
        auto issue_tap = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + k * kstep);
                if (PASSES == 3 && stacked) {
                    const uint64_t bd = ((uint64_t)b_hi_st << 32) | (b_lo + k * kstep);
                    umma_f16(d_tmem, ad, bd, idesc_n128, k == 0 ? acc : 1u);
                    umma_f16(d_tmem, ad + lo_off, bd, idesc, 1u);
                    continue;
                }
                const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + k * kstep);
                if (PASSES == 3) {   // small cross terms first, then the leading term
                    umma_f16(d_tmem, ad + lo_off, bd, idesc, k == 0 ? acc : 1u);
                    umma_f16(d_tmem, ad, bd + lo_off, idesc, 1u);
                    umma_f16(d_tmem, ad, bd, idesc, 1u);
                } else {
                    umma_f16(d_tmem, ad, bd, idesc, k == 0 ? acc : 1u);
                }
            }
        };
        const int acc_cols = stacked ? 128 : a.bn;        // TMEM columns per accumulator buffer
        int sa = 0, pa = 0, as = 0, pacc = 0, trace_m = 0;
        if (a.w_resident && WIN != 2) {
            // ---- weights resident: one wait per activation stage, then TAPS x KSTEPS x PASSES back-to-back MMAs
            mbar_wait(B_FULL(0), 0);
            tc_fence_after();
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                mbar_wait(ACC_EMPTY(as), pacc ^ 1);
                tc_fence_after();
                if (lane == 0) TC_TRACE(3, trace_m);
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * acc_cols);
                uint32_t b_cur = b_base;                                       // panels [cb][tap]; one N tile, no K split
                for (int cb = 0; cb < a.ncb; ++cb) {
                    mbar_wait(A_FULL(sa), pa);
                    tc_fence_after();
                    if (lane == 0 && cb == 0) TC_TRACE(4, trace_m);
                    const uint32_t a_base = lo0 | (smem_u32(sA + sa * A_STAGE_BYTES) >> 4);
                    if (leader) {
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap)
                            issue_tap(d_tmem, a_base + (uint32_t)((((tap / WIN) * HPITCH_PX + (tap % WIN)) * 128) >> 4),
                                      b_cur + (uint32_t)tap * b_step, (cb | tap) ? 1u : 0u);
                        umma_commit(A_EMPTY(sa));
                    }
                    __syncwarp();
                    b_cur += (uint32_t)TAPS * b_step;
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
                if (leader) umma_commit(ACC_FULL(as));
                __syncwarp();
                if (lane == 0) { TC_TRACE(5, trace_m); ++trace_m; }
                if (++as == 2) { as = 0; pacc ^= 1; }
            }
        } else {
            // ---- weights streamed: one panel per (channel block, tap) through the SB-deep ring
            int sb = 0, pb = 0;
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int sa0 = sa, pa0 = pa;      // A-stationary: every N tile of the item re-reads the same activation stages
                for (int nti = 0; nti < nt_inner; ++nti) {
                    int nt, img, ty, tx, ks;
                    decode(w, nti, nt, img, ty, tx, ks);
                    const int cb0 = ks * cb_per, cb1 = min(a.ncb, cb0 + cb_per);
                    const bool last_nt = nti == nt_inner - 1;
                    sa = sa0; pa = pa0;
                    mbar_wait(ACC_EMPTY(as), pacc ^ 1);
                    tc_fence_after();
                    if (lane == 0) TC_TRACE(3, trace_m);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * acc_cols);
                    uint32_t acc = 0;
                    for (int cb = cb0; cb < cb1; ++cb) {
                        mbar_wait(A_FULL(sa), pa);     // (returns at once when an earlier N tile already saw this phase)
                        tc_fence_after();
                        if (lane == 0 && cb == cb0) TC_TRACE(4, trace_m);
                        const uint32_t a_base = lo0 | (smem_u32(sA + sa * A_STAGE_BYTES) >> 4);
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap) {
                            if (WIN == 2 && !stage_live(cb, tap)) continue;
                            mbar_wait(B_FULL(sb), pb);
                            tc_fence_after();
                            if (leader) {
                                issue_tap(d_tmem, a_base + (uint32_t)((((tap / WIN) * HPITCH_PX + (tap % WIN)) * 128) >> 4),
                                          b_base + (uint32_t)sb * b_step, acc);
                                umma_commit(B_EMPTY(sb));            // frees the weight stage when these MMAs retire
                            }
                            __syncwarp();
                            acc = 1;
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        if (leader && last_nt) umma_commit(A_EMPTY(sa));
                        __syncwarp();
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                    if (leader) umma_commit(ACC_FULL(as));
                    __syncwarp();
                    if (lane == 0) { TC_TRACE(5, trace_m); ++trace_m; }
                    if (++as == 2) { as = 0; pacc ^= 1; }
                }
            }
        }
        // every MMA of this CTA is issued: let the next kernel's CTAs be scheduled (KEEP_PDL_CONV_TRIGGER=1).  The next kernel
        // is almost always a short one (split-K reduce / GroupNorm finalize): its blocks then sit at griddepcontrol.wait while
        // this grid's last epilogues drain, and start the moment it completes
#if KEEP_PDL_CONV_TRIGGER
        if (leader) pdl_trigger();
#endif
    } else if (warp < kEpiWarps) {
        // =========================== epilogue (warps 0-3 <-> TMEM lane quarters) ===========================
        // Each thread owns one accumulator row (= one output pixel) and walks its BN columns 16 at a time:
        // tcgen05.ld -> (+bias from shared memory) -> activation -> (+residual, loaded while the TMEM load is in
        // flight) -> 64 contiguous bytes to HBM.  Everything stays in registers (a first version spilled the row to
        // local memory and re-fetched the bias from L2 every 16 columns: 8K cycles per tile, the kernel's bottleneck).
        int as = 0, pacc = 0, trace_e = 0, bias_nt = -1;
        bool waited = false;
        const int row = warp * 32 + lane;          // accumulator row = output pixel within the tile
        const int r = row >> 3, c = row & 7;
        for (int w = blockIdx.x; w < total; w += gridDim.x)
        for (int nti = 0; nti < nt_inner; ++nti) {
            int nt, img, ty, tx, ks;
            decode(w, nti, nt, img, ty, tx, ks);
            if (nt != bias_nt) {   // stage this N tile's bias (zeros when absent / beyond cout) in shared memory
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = threadIdx.x; i < a.bn; i += kEpiWarps * 32) {
                    const int nn = nt * a.bn + i;
                    s_bias[i] = (a.bias && (a.splitk == 1 || a.cluster_k) && nn < a.cout) ? a.bias[nn] : 0.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                bias_nt = nt;
            }
            long long pixel;
            bool ok;
            if (conv3) {
                const int oy = ty * 16 + r, ox = tx * 8 + c;
                ok = oy < a.ho && ox < a.wo;
                pixel = ((long long)img * a.ho + oy) * a.wo + ox;
            } else {
                const int q = ty * 128 + row;
                ok = q < a.h * a.w;
                pixel = (long long)(img * a.h * a.w + q);
            }
            const int n0 = nt * a.bn;
            mbar_wait(ACC_FULL(as), pacc);
            tc_fence_after();
            if (!waited) { pdl_wait(); waited = true; }   // residual reads / output writes below (returns at once: the producers passed it)
            if (threadIdx.x == 0) TC_TRACE(6, trace_e);
            const bool stacked = KEEP_TC_STACKED && PASSES == 3 && a.bn == 64;   // accumulator = [Ah*Wh + Al*Wh | Ah*Wl]: add columns j and 64 + j
            const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * (stacked ? 128 : a.bn));
            const bool partial_out = a.splitk > 1;
            if (a.cluster_k) {
                // cluster split-K: park this CTA's fp32 partial tile in its (now idle) activation stages, row = pixel,
                // 16-byte pieces XOR-swizzled by the row so that both this row-per-thread write and the piece-per-thread
                // reads of the reduction below are bank-conflict free; the cluster reduces it after the roles join
                const int row = warp * 32 + lane;
                uint8_t* dstrow = sA + (size_t)row * a.bn * 4;
                for (int j = 0; j < a.bn; j += 16) {
                    uint32_t rr[16];
                    __syncwarp();
                    tmem_ld16(t0 + (uint32_t)j, rr);
                    tmem_ld_wait();
                    if (stacked) {
                        uint32_t r2[16];
                        tmem_ld16(t0 + 64u + (uint32_t)j, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(__uint_as_float(rr[e]) + __uint_as_float(r2[e]));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        *reinterpret_cast<uint4*>(dstrow + ((((j >> 2) + e) ^ (row & 7)) << 4)) = make_uint4(rr[4 * e], rr[4 * e + 1], rr[4 * e + 2], rr[4 * e + 3]);
                }
                tc_fence_before();
                mbar_arrive(ACC_EMPTY(as));
                if (threadIdx.x == 0) { TC_TRACE(7, trace_e); ++trace_e; }
                if (++as == 2) { as = 0; pacc ^= 1; }
                continue;
            }
            if (KEEP_TC_PARTIAL32 && partial_out && !stacked) {
                // split-K partial tile: a plain TMEM -> partial-buffer copy, 32 columns (two loads in flight) per round trip --
                // on the small layers of the per-frame chain the epilogue follows the CTA's only K loop, so its TMEM latency
                // is on the layer's critical path
                for (int j = 0; j < a.bn; j += 32) {
                    uint32_t rr[32];
                    __syncwarp();
                    tmem_ld16(t0 + (uint32_t)j, rr);
                    tmem_ld16(t0 + (uint32_t)j + 16u, rr + 16);
                    tmem_ld_wait();
                    const int nn = n0 + j;
                    if (ok && nn < a.cout) {
                        float* o = a.partial + (size_t)ks * a.M * a.cout + (size_t)pixel * a.cout + nn;
                        const float* v = reinterpret_cast<const float*>(rr);
                        stg256(o, v);
                        stg256(o + 8, v + 8);
                        if (nn + 16 < a.cout) { stg256(o + 16, v + 16); stg256(o + 24, v + 24); }
                    }
                }
                tc_fence_before();
                mbar_arrive(ACC_EMPTY(as));
                if (threadIdx.x == 0) { TC_TRACE(7, trace_e); ++trace_e; }
                if (++as == 2) { as = 0; pacc ^= 1; }
                continue;
            }
            for (int j = 0; j < a.bn; j += 16) {
                uint32_t rr[16];
                __syncwarp();                              // tcgen05.ld is .sync.aligned: reconverge after divergent stores
                tmem_ld16(t0 + (uint32_t)j, rr);
                const int nn = n0 + j;
                const bool col_ok = nn < a.cout;           // cout % 16 == 0 (warp-uniform)
                const bool live = ok && col_ok;
                const size_t off = (size_t)pixel * a.cout + nn;
                float4 rs[4];
                const bool has_res = live && !partial_out && a.res != nullptr;
                if (has_res) {
                    if (a.res_dt == F32) {
                        ldg256(reinterpret_cast<const float*>(a.res) + off, reinterpret_cast<float*>(&rs[0]));
                        ldg256(reinterpret_cast<const float*>(a.res) + off + 8, reinterpret_cast<float*>(&rs[2]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) rs[e] = ld4(reinterpret_cast<const __half*>(a.res), off + 4 * e);
                    }
                }
                tmem_ld_wait();
                if (stacked) {   // (warp-uniform) second half of the accumulator: the Ah * Wl term (both loads behind one wait: measured slower)
                    uint32_t r2[16];
                    tmem_ld16(t0 + 64u + (uint32_t)j, r2);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(__uint_as_float(rr[e]) + __uint_as_float(r2[e]));
                }
#ifdef KEEP_TC_EPI_TRACE
                if (threadIdx.x == 0 && trace_e == 0) TC_TRACE(8, j >> 4);
#endif
                if (!col_ok || (!live && !a.gn_part)) continue;
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[e]);
                if (partial_out) {
                    if (live) {
                        float* o = a.partial + (size_t)ks * a.M * a.cout + off;
                        stg256(o, v);
                        stg256(o + 8, v + 8);
                    }
#ifdef KEEP_TC_EPI_TRACE
                    if (threadIdx.x == 0 && trace_e == 0) TC_TRACE(9, j >> 4);
#endif
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 b = *reinterpret_cast<const float4*>(s_bias + j + 4 * e);
                    v[4 * e] += b.x; v[4 * e + 1] += b.y; v[4 * e + 2] += b.z; v[4 * e + 3] += b.w;
                }
                if (a.act == ACT_RELU) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.0f);
                } else if (a.act == ACT_GELU) {   // GMFlow / transformer FFNs: N = 1024 outputs per row -- inline, 16 independent erff chains
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752440f));
                } else if (a.act != ACT_NONE) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = act_slow(v[e], a.act);
                }
                if (has_res) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { v[4 * e] += rs[e].x; v[4 * e + 1] += rs[e].y; v[4 * e + 2] += rs[e].z; v[4 * e + 3] += rs[e].w; }
                }
                if (a.gn_part) {
                    // GroupNorm(32) statistics of the final values, for the consumer's normalisation (vqgan_arch.py:16-17): this
                    // chunk's 16 channels are 16 / cpg groups; per lane (= pixel) group sums, then a warp reduce-scatter over the
                    // 32 pixels in a fixed tree (deterministic): 16 values -> 16 shuffles; lane 2i ends up with value i
                    float w16[16];
                    const int cpg = a.gn_cpg;               // 2, 4, 8 or 16 (warp-uniform)
                    {
                        float gs[8], gq[8];                 // channel pairs first, then pairs of pairs up to the group width
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            gs[g] = ok ? v[2 * g] + v[2 * g + 1] : 0.0f;
                            gq[g] = ok ? fmaf(v[2 * g], v[2 * g], v[2 * g + 1] * v[2 * g + 1]) : 0.0f;
                        }
                        if (cpg >= 4) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) { gs[g] = gs[2 * g] + gs[2 * g + 1]; gq[g] = gq[2 * g] + gq[2 * g + 1]; }
                        }
                        if (cpg >= 8) {
#pragma unroll
                            for (int g = 0; g < 2; ++g) { gs[g] = gs[2 * g] + gs[2 * g + 1]; gq[g] = gq[2 * g] + gq[2 * g + 1]; }
                        }
                        if (cpg >= 16) { gs[0] += gs[1]; gq[0] += gq[1]; }
#pragma unroll
                        for (int g = 0; g < 8; ++g) { w16[2 * g] = gs[g]; w16[2 * g + 1] = gq[g]; }   // slots >= 16 / cpg: unused
                    }
#pragma unroll
                    for (int half = 8; half >= 1; half >>= 1) {
                        const bool up = (lane & (half * 2)) != 0;
#pragma unroll
                        for (int e = 0; e < half; ++e) {
                            const float keep = up ? w16[e + half] : w16[e];
                            const float send = up ? w16[e] : w16[e + half];
                            w16[e] = keep + __shfl_xor_sync(0xffffffffu, send, half * 2);
                        }
                    }
                    w16[0] += __shfl_xor_sync(0xffffffffu, w16[0], 1);
                    const int vi = lane >> 1, gslot = vi >> 1;
                    if ((lane & 1) == 0 && gslot < (16 >> (cpg == 2 ? 1 : (cpg == 4 ? 2 : (cpg == 8 ? 3 : 4))))) {
                        const int mt_in_img = ty * a.tiles_x + tx;
                        a.gn_part[(((size_t)img * 32 + ((nn >> a.gn_cpg_shift) + gslot)) * a.gn_P + (size_t)mt_in_img * 4 + warp) * 2 + (vi & 1)] = w16[0];
                    }
                    if (!ok) continue;
                }
                if (a.out_dt == F32) {
                    float* o = reinterpret_cast<float*>(a.out) + off;
                    stg256(o, v);
                    stg256(o + 8, v + 8);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        st4(reinterpret_cast<__half*>(a.out), off + 4 * e, make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]));
                }
#ifdef KEEP_TC_EPI_TRACE
                if (threadIdx.x == 0 && trace_e == 0) TC_TRACE(9, j >> 4);
#endif
            }
            tc_fence_before();
            mbar_arrive(ACC_EMPTY(as));
            if (threadIdx.x == 0) { TC_TRACE(7, trace_e); ++trace_e; }
            if (++as == 2) { as = 0; pacc ^= 1; }
        }
    }

    if (a.cluster_k) {
        // ---- cluster split-K reduction over distributed shared memory (replaces the partial round trip through L2 and the
        // separate reduce kernel): after the first cluster barrier every CTA holds its partial tile in shared memory; the
        // CTA of rank r then sums rows [r*R, (r+1)*R) over all ranks IN RANK ORDER (deterministic), applies bias /
        // activation / residual and writes the final rows with coalesced 16-byte stores.
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (warp < kEpiWarps) {
            const int S = a.splitk, rank = (int)(blockIdx.x % (unsigned)S);
            int nt, img, ty, tx, ks;
            decode(blockIdx.x, 0, nt, img, ty, tx, ks);
            const int R = (128 + S - 1) / S, r0 = rank * R, r1 = min(128, r0 + R);
            const int q4 = a.bn >> 2;                       // 16-byte pieces per row
            const int n0 = nt * a.bn;
            const uint32_t sbase = smem_u32(sA);
            for (int idx = threadIdx.x; idx < (r1 - r0) * q4; idx += kEpiWarps * 32) {
                const int row = r0 + idx / q4, c4 = idx - (idx / q4) * q4;
                const int col = n0 + c4 * 4;
                long long pixel;
                bool ok;
                if (conv3) {
                    const int oy = ty * 16 + (row >> 3), ox = tx * 8 + (row & 7);
                    ok = oy < a.ho && ox < a.wo;
                    pixel = ((long long)img * a.ho + oy) * a.wo + ox;
                } else {
                    const long long q = (long long)ty * 128 + row;
                    ok = q < (long long)a.h * a.w;
                    pixel = (long long)img * a.h * a.w + q;
                }
                if (!ok || col >= a.cout) continue;
                const uint32_t laddr = sbase + (uint32_t)(row * a.bn * 4 + ((c4 ^ (row & 7)) << 4));
                // all S remote loads in flight at once (one after the other they cost S x the DSMEM latency: the round-1
                // version of this loop made the cluster path slower than the separate reduce kernel), summed in rank order
                float4 pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    pk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < S) {
                        uint32_t raddr;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(k));
                        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pk[k].x), "=f"(pk[k].y), "=f"(pk[k].z), "=f"(pk[k].w) : "r"(raddr));
                    }
                }
                float4 v = pk[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) { v.x += pk[k].x; v.y += pk[k].y; v.z += pk[k].z; v.w += pk[k].w; }
                const float4 b = *reinterpret_cast<const float4*>(s_bias + c4 * 4);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                if (a.act != ACT_NONE) { v.x = act_slow(v.x, a.act); v.y = act_slow(v.y, a.act); v.z = act_slow(v.z, a.act); v.w = act_slow(v.w, a.act); }
                const size_t off = (size_t)pixel * a.cout + col;
                if (a.res) {
                    const float4 r4 = a.res_dt == F32 ? ld4(reinterpret_cast<const float*>(a.res), off) : ld4(reinterpret_cast<const __half*>(a.res), off);
                    v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                }
                if (a.out_dt == F32) st4(reinterpret_cast<float*>(a.out), off, v);
                else st4(reinterpret_cast<__half*>(a.out), off, v);
            }
        }
        // no CTA may exit (and release its shared memory) while a peer is still reading it
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(9, 3);                 // all roles done
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int tc_cb(int passes) { return cb_of(passes); }

int tc_pick_bn(int cout, long long m_tiles, int passes) {
    // N tile: a multiple of 16 up to 256.  Wider tiles amortise the A-operand transform; narrower ones give
    // more CTAs when the layer is small.  The split-precision mode triples the MMAs per tile -> BN <= 128 keeps
    // both TMEM accumulators and a deep weight pipeline.
    if (cout <= 64) return (cout + 15) / 16 * 16;
    if (cout == 96) return 96;
    if (cout == 192 && passes == 1) return 192;
    if (passes == 1 && cout % 256 == 0 && m_tiles * (cout / 256) >= 148) return 256;
    if (cout % 128 != 0 && cout % 96 == 0) return 96;
    return 128;
}

bool tc_is_s2d(const ConvArgs& a) {
    return a.kh == 3 && a.kw == 3 && a.stride == 2 && a.up == 1 && a.c1 == 0 && a.pad_t == a.pad_l && (a.pad_t == 0 || a.pad_t == 1) &&
           a.h % 2 == 0 && a.w % 2 == 0 && a.ho == a.h / 2 && a.wo == a.w / 2;
}

int tc_virtual_cin(const ConvArgs& a, int passes) {
    const int cin = a.c0 + a.c1, cb = cb_of(passes);
    return tc_is_s2d(a) ? 4 * ((cin + cb - 1) / cb) * cb : cin;
}

bool tc_eligible(const ConvArgs& a) {
    const int cin = a.c0 + a.c1;
    if (tc_is_s2d(a)) return a.cout % 16 == 0 && cin % 8 == 0 && cin >= 32;
    const bool k3 = a.kh == 3 && a.kw == 3 && a.stride == 1 && a.pad_t == 1 && a.pad_l == 1 && a.ho == a.h * a.up && a.wo == a.w * a.up;
    const bool k1 = a.kh == 1 && a.kw == 1 && a.stride == 1 && a.pad_t == 0 && a.pad_l == 0 && a.up == 1;
    if (!k3 && !k1) return false;
    if (a.cout % 16 != 0 || cin % 8 != 0 || cin < 32) return false;
    if (a.c1 > 0 && a.c0 % 8 != 0) return false;
    if (k1 && ((long long)a.h * a.w) % 8 != 0) return false;
    if (a.up != 1 && a.up != 2) return false;
    return true;
}

// a panel = bn rows x 128 bytes (64 halfs): 64 input channels, or [hi | lo] of 32 input channels (split precision)
size_t tc_packed_weight_halfs(int cin, int cout, int taps, int bn, int passes) {
    const int cb = cb_of(passes), ncb = (cin + cb - 1) / cb, ntile = (cout + bn - 1) / bn;
    return (size_t)ntile * ncb * taps * bn * 64;
}

// position `idx` (halfs) inside a packed panel set -> (n tile, channel block, tap, row, logical channel k within the block,
// hi/lo part); the 16-byte chunk j of a row is stored at chunk position j ^ (row & 7)   (SWIZZLE_128B, K-major)
struct PanelPos { int nt, cb, tap, row, k, part; };
// Split-precision layers with a 64-wide N tile use the STACKED panel instead: 128 rows of 64 bytes -- rows 0-63 the hi parts
// (32 channels), rows 64-127 the lo parts -- in the SWIZZLE_64B K-major layout (16-byte chunk j of row r at chunk
// j ^ ((r >> 1) & 3)).  One N = 128 MMA then forms Ah*Wh and Ah*Wl side by side and one N = 64 MMA adds Al*Wh: two A-operand
// fetches per K step instead of three (the N = 64 MMAs are bound by shared-memory operand fetch, profiles/r1_ubench_umma_rate.md).
__host__ __device__ inline bool tc_stacked(int bn, int passes) { return KEEP_TC_STACKED && passes == 3 && bn == 64; }
__host__ __device__ inline PanelPos panel_pos(size_t idx, int bn, int taps, int ncb, int passes) {
    PanelPos q;
    size_t r = idx;
    const int e = (int)(r % 8); r /= 8;
    if (tc_stacked(bn, passes)) {
        const int chunk = (int)(r % 4); r /= 4;
        const int prow = (int)(r % 128); r /= 128;       // physical row of the 128 x 64-byte panel
        q.tap = (int)(r % taps); r /= taps;
        q.cb = (int)(r % ncb); r /= ncb;
        q.nt = (int)r;
        const int cl = chunk ^ ((prow >> 1) & 3);
        q.row = prow & 63; q.part = prow >> 6; q.k = (cl << 3) + e;
        return q;
    }
    const int chunk = (int)(r % 8); r /= 8;
    q.row = (int)(r % bn); r /= bn;
    q.tap = (int)(r % taps); r /= taps;
    q.cb = (int)(r % ncb); r /= ncb;
    q.nt = (int)r;
    const int cl = chunk ^ (q.row & 7);                  // logical chunk
    if (passes == 3) { q.part = cl >> 2; q.k = ((cl & 3) << 3) + e; }
    else { q.part = 0; q.k = (cl << 3) + e; }
    return q;
}
__host__ __device__ inline __half split_part(float v, int part, int wide = 0) {
    if (wide) {   // bf16 (hi, lo) pair, bit patterns stored in the fp16 panel (KEEP_FLAG_TC_WIDE: both MMA operands are bf16)
        const __nv_bfloat16 hb = __float2bfloat16_rn(v);
        const __nv_bfloat16 r = part == 0 ? hb : __float2bfloat16_rn(v - __bfloat162float(hb));
        return *reinterpret_cast<const __half*>(&r);
    }
    const __half hi = __float2half_rn(v);
    return part == 0 ? hi : __float2half_rn(v - __half2float(hi));
}

// OIHW fp32 (host) -> [ntile][cb][tap] panels
void tc_pack_weights(const float* w_oihw, int cout, int cin, int kh, int kw, int bn, int passes, __half* out, int wide) {
    const int taps = kh * kw, cb = cb_of(passes), ncb = (cin + cb - 1) / cb;
    const size_t total = tc_packed_weight_halfs(cin, cout, taps, bn, passes);
    for (size_t idx = 0; idx < total; ++idx) {
        const PanelPos q = panel_pos(idx, bn, taps, ncb, passes);
        const int o = q.nt * bn + q.row, i = q.cb * cb + q.k;
        float v = 0.0f;
        if (o < cout && i < cin) v = w_oihw[(((size_t)o * cin + i) * kh + q.tap / kw) * kw + q.tap % kw];
        out[idx] = split_part(v, q.part, wide);
    }
}

namespace {
// device-side repack: fp32 [(tap*cin + ci)][cout] (the CUDA-core path's layout) -> tcgen05 fp16 swizzled panels
__global__ void tc_repack_kernel(const float* __restrict__ w, int cin, int cout, int taps, int bn, int ncb, int passes, int s2d_pad,
                                 size_t total, __half* __restrict__ out, int wide) {
    pdl_prologue_light();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const PanelPos q = panel_pos(idx, bn, taps, ncb, passes);
    const int cb = cb_of(passes);
    const int o = q.nt * bn + q.row;
    float v = 0.0f;
    if (s2d_pad < 0) {
        const int i = q.cb * cb + q.k;
        if (o < cout && i < cin) v = w[((size_t)q.tap * cin + i) * cout + o];
    } else {   // virtual space-to-depth: block cb = parity * ncbr + real block; tap = (a, b) of the 2x2 cell window
        const int ncbr = ncb / 4, par = q.cb / ncbr, i = (q.cb - par * ncbr) * cb + q.k;
        const int ky = 2 * (q.tap >> 1) + (par >> 1) - s2d_pad, kx = 2 * (q.tap & 1) + (par & 1) - s2d_pad;
        if (o < cout && i < cin && ky >= 0 && ky <= 2 && kx >= 0 && kx <= 2) v = w[((size_t)(ky * 3 + kx) * cin + i) * cout + o];
    }
    out[idx] = split_part(v, q.part, wide);
}

// Activation matrix -> tcgen05 "weight" panels, so that C[z] = A[z] * B[z]^T (attention QK^T, PV) runs on the same kernel:
// B[z] is (N x K) with element (n, k) at src[z*bstride + n*ld_n + k*ld_k] (ld_k = 1: row-major; ld_n = 1: transposed view).
// One thread per 8 consecutive k of one row: 8 fp32 in (two 16-byte loads when the source is row-major; eight loads that
// coalesce across the warp's rows when it is a transposed view), one 16-byte hi chunk (+ one lo chunk) out.
__global__ void __launch_bounds__(256) tc_pack_matrix_kernel(const float* __restrict__ src, long long bstride, int ld_n, int ld_k,
                                                             int N, int K, int bn, int ncb, int ntile, int passes, float alpha,
                                                             size_t per_batch, size_t total_units, __half* __restrict__ out) {
    pdl_prologue_light();
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total_units) return;
    const int cb = cb_of(passes), upr = cb / 8;
    size_t r = u;
    int row, j;
    if (ld_n == 1) { row = (int)(r % bn); r /= bn; j = (int)(r % upr); r /= upr; }
    else { j = (int)(r % upr); r /= upr; row = (int)(r % bn); r /= bn; }
    const int cbi = (int)(r % ncb); r /= ncb;
    const int nt = (int)(r % ntile);
    const size_t z = r / ntile;
    const int n = nt * bn + row, k0 = cbi * cb + j * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (n < N) {
        const float* p = src + z * bstride + (size_t)n * ld_n + (size_t)k0 * ld_k;
        if (ld_k == 1 && k0 + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (k0 + i < K) v[i] = p[(size_t)i * ld_k];
        }
    }
    __half2 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x = alpha * v[2 * i], y = alpha * v[2 * i + 1];
        hi[i] = __floats2half2_rn(x, y);
        const float2 f = __half22float2(hi[i]);
        lo[i] = __floats2half2_rn(x - f.x, y - f.y);
    }
    __half* base = out + z * per_batch + ((size_t)(nt * ncb + cbi) * bn + row) * 64;
    *reinterpret_cast<uint4*>(base + ((j ^ (row & 7)) << 3)) = *reinterpret_cast<const uint4*>(hi);
    if (passes == 3) *reinterpret_cast<uint4*>(base + (((j + 4) ^ (row & 7)) << 3)) = *reinterpret_cast<const uint4*>(lo);
}
}  // namespace

size_t tc_pack_matrix(const float* src, long long bstride, int ld_n, int ld_k, int nbatch, int N, int K, int bn, int passes,
                      float alpha, __half* out, cudaStream_t s) {
    const int cb = cb_of(passes), ncb = (K + cb - 1) / cb, ntile = (N + bn - 1) / bn;
    KEEP_CHECK(!tc_stacked(bn, passes), "tc_pack_matrix: 64-wide N tiles use the stacked panel layout (weights only)");
    const size_t per_batch = tc_packed_weight_halfs(K, N, 1, bn, passes);
    if (out) {
        const size_t units = (size_t)nbatch * ntile * ncb * bn * (cb / 8);
        launch_k(tc_pack_matrix_kernel, dim3((unsigned)((units + 255) / 256)), dim3(256), 0, s, src, bstride, ld_n, ld_k, N, K, bn, ncb, ntile, passes,
                                                                             alpha, per_batch, units, out);
        CUDA_CHECK(cudaGetLastError());
    }
    return per_batch;
}

void tc_repack_device(const float* w_kc, int cin, int cout, int taps, int bn, int passes, int s2d_pad, __half* out, cudaStream_t s, int wide) {
    // stride-2 mode: `cin` real channels are seen as 4 * ceil(cin/cb) * cb virtual channels with a 2x2 (taps = 4) window
    const int cb = cb_of(passes);
    const int vcin = s2d_pad >= 0 ? 4 * ((cin + cb - 1) / cb) * cb : cin;
    const int vtaps = s2d_pad >= 0 ? 4 : taps;
    const int ncb = (vcin + cb - 1) / cb;
    const size_t total = tc_packed_weight_halfs(vcin, cout, vtaps, bn, passes);
    launch_k(tc_repack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, w_kc, cin, cout, vtaps, bn, ncb, passes, s2d_pad, total, out, wide);
    CUDA_CHECK(cudaGetLastError());
}

// GroupNorm(32) statistics slots per image that the producing kernel writes (ConvArgs::gn_part): 4 per m tile (one per
// epilogue warp) without split-K, one per 1024-element block of the reduce kernel with it; 0 = the layer cannot emit them
int conv_gn_slots(const ConvArgs& a, int splitk) {
    if (!tc_eligible(a) || a.cout % 32 != 0 || a.out_dt != F32) return 0;
    const int cpg = a.cout / 32;
    if (cpg != 2 && cpg != 4 && cpg != 8 && cpg != 16) return 0;
    const long long hw = (long long)a.ho * a.wo;
    if (splitk > 1) {
        if (cpg < 4 || (hw * a.cout) % 1024 != 0 || 1024 % a.cout != 0) return 0;
        return (int)(hw * a.cout / 1024);
    }
    const bool s2d = tc_is_s2d(a);
    if (s2d || a.kh == 3) return cdiv(a.ho, 16) * cdiv(a.wo, 8) * 4;
    return cdiv((long long)a.h * a.w, 128) * 4;
}

int tc_pick_splitk(long long m_tiles, int ntile_n, int ncb) {
    static const int target = getenv("KEEP_TC_SPLIT_TARGET") ? atoi(getenv("KEEP_TC_SPLIT_TARGET")) : 80;   // CTAs to aim for
    static const int nosplit = getenv("KEEP_TC_SPLIT_MIN") ? atoi(getenv("KEEP_TC_SPLIT_MIN")) : 96;          // enough tiles: no split
    const long long ctas = m_tiles * ntile_n;
    if (ctas >= nosplit || ncb < 2) return 1;
    long long s = (target + ctas - 1) / ctas;
    if (s > ncb) s = ncb;
    return (int)(s < 1 ? 1 : s);
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int env_swap() {
    static int v = -1;
    if (v < 0) v = env_int("KEEP_TC_BASE_OFFSET", 0) == 1 ? 1 : 0;
    return v;
}

long long* g_tc_trace = nullptr;   // debug: device buffer of 10*16 clock64 stamps (keepop_tc_trace)

constexpr int SMEM_FIXED = 1024 + 8 * (2 * MAX_SA + 2 * MAX_SB + 4) + 16 + 256 * (int)sizeof(float);   // alignment slack, barriers, TMEM slot, bias
constexpr int SMEM_BUDGET = 227 * 1024 - SMEM_FIXED;

// weights resident in shared memory for the CTA's lifetime: one N tile, no K split, per-layer (not per-image) panels, and
// the whole panel set fits next to at least two activation stages (KEEP_TC_RESIDENT=0 disables)
static int a_stage_bytes(int win) { return win == 1 ? 128 * 128 : A_SUB_BYTES; }

static bool pick_resident(const TcConvArgs& t, int splitk) {
    static int en = -1;
    if (en < 0) en = env_int("KEEP_TC_RESIDENT", 1);
    if (!en || t.ntile_n != 1 || splitk != 1 || t.wt_img_stride != 0 || t.s2d) return false;
    return (size_t)t.ncb * t.taps * t.bn * 128 + 2 * (size_t)a_stage_bytes(t.win) <= (size_t)SMEM_BUDGET;
}

// A-stationary walk over the N tiles (see the kernel): 1x1 layers with several N tiles, every activation stage of an item
// resident at once (<= MAX_SA stages next to >= 3 weight stages), and enough (m tile, k split) items to fill the grid
static bool pick_a_stationary(const TcConvArgs& t, int splitk, int grid_cap) {
    static int en = -1;
    if (en < 0) en = env_int("KEEP_TC_ASTAT", 1);
    if (!en || t.win != 1 || t.ntile_n < 2) return false;
    const int cb_per = cdiv(t.ncb, splitk);
    if (cb_per > MAX_SA) return false;
    if ((size_t)cb_per * a_stage_bytes(1) + 3 * (size_t)t.bn * 128 > (size_t)SMEM_BUDGET) return false;
    const long long items = (long long)t.n * t.tiles_y * t.tiles_x * splitk;
    return items >= grid_cap;
}

static void pick_stages(int win, int bn, size_t resident_bytes, int min_sa, int& sa, int& sb) {
    const int a_stage = a_stage_bytes(win), b_stage = bn * 128;
    sa = min_sa > 2 ? min_sa : 2;
    if (resident_bytes) {
        sb = 0;
        while (sa < MAX_SA && resident_bytes + (size_t)(sa + 1) * a_stage <= (size_t)SMEM_BUDGET) ++sa;
        return;
    }
    sb = 2;
    // grow the weight pipeline first (up to 9 weight stages are consumed per activation stage), then the activation one
    while (sb < 6 && sa * a_stage + (sb + 1) * b_stage <= SMEM_BUDGET) ++sb;
    while (sa < (min_sa > 4 ? min_sa : 4) && (sa + 1) * a_stage + sb * b_stage <= SMEM_BUDGET) ++sa;
    while (sb < MAX_SB && sa * a_stage + (sb + 1) * b_stage <= SMEM_BUDGET) ++sb;
    while (sa < MAX_SA && (sa + 1) * a_stage + sb * b_stage <= SMEM_BUDGET) ++sa;
}

using Kern = void (*)(const TcConvArgs);
static const Kern kerns[2][2][3] = {
    {{conv_tc_kernel<1, false, 1>, conv_tc_kernel<1, false, 2>, conv_tc_kernel<1, false, 3>},
     {conv_tc_kernel<1, true, 1>, conv_tc_kernel<1, true, 2>, conv_tc_kernel<1, true, 3>}},
    {{conv_tc_kernel<3, false, 1>, conv_tc_kernel<3, false, 2>, conv_tc_kernel<3, false, 3>},
     {conv_tc_kernel<3, true, 1>, conv_tc_kernel<3, true, 2>, conv_tc_kernel<3, true, 3>}}};
// wide-range variant (bf16 activation pairs): split-precision mode, fp32 feature maps only
static const Kern kerns_wide[3] = {conv_tc_kernel<3, false, 1, true>, conv_tc_kernel<3, false, 2, true>, conv_tc_kernel<3, false, 3, true>};

// opt every variant into 227 KB of dynamic shared memory, once per device (also called at engine creation, so that no
// attribute call happens while a CUDA graph is being captured)
void tc_configure_device() {
    static unsigned long long configured = 0;
    if (!first_use_on_current_device(&configured)) return;
    for (int i = 0; i < 12; ++i)
        CUDA_CHECK(cudaFuncSetAttribute(kerns[i / 6][(i / 3) % 2][i % 3], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    for (int i = 0; i < 3; ++i)
        CUDA_CHECK(cudaFuncSetAttribute(kerns_wide[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

int conv2d_tc(const ConvArgs& a, const __half* packed, int bn, int passes, int splitk, float* partial, int num_sms,
              cudaStream_t s) {
    KEEP_CHECK(tc_eligible(a), "conv2d_tc: layer not eligible for the tcgen05 kernel");
    KEEP_CHECK(passes == 1 || passes == 3, "conv2d_tc: passes must be 1 or 3");
    const int cb = cb_of(passes);
    TcConvArgs t;
    t.in0 = a.in0; t.in1 = a.in1; t.in0_dt = a.in0_dt; t.in1_dt = a.in1_dt; t.c0 = a.c0; t.c1 = a.c1;
    KEEP_CHECK(a.ld0 == 0 || (a.kh == 1 && a.c1 == 0 && a.ld0 >= a.c0 && a.ld0 % 8 == 0), "conv2d_tc: a strided A operand needs a 1x1 layer with one source and ld0 %% 8 == 0");
    t.ld0 = a.ld0 > 0 ? a.ld0 : a.c0;
    t.n = a.n; t.h = a.h; t.w = a.w; t.up = a.up;
    t.pre_scale = a.pre_scale; t.pre_shift = a.pre_shift; t.pre_act = a.pre_act; t.pre_exact = a.pre_exact;
    t.wt = packed; t.bias = a.bias;
    t.wt_img_stride = a.wt_img_stride;
    const bool s2d = tc_is_s2d(a);
    t.s2d = s2d ? 1 : 0;
    t.win = s2d ? 2 : (a.kh == 3 ? 3 : 1);
    t.taps = t.win * t.win; t.cout = a.cout; t.bn = bn;
    t.ho = a.ho; t.wo = a.wo;
    t.ncbr = (a.c0 + a.c1 + cb - 1) / cb;
    t.ncb = s2d ? 4 * t.ncbr : t.ncbr;
    t.pad_t = a.pad_t; t.pad_l = a.pad_l;
    if (t.win > 1) { t.tiles_y = cdiv(a.ho, 16); t.tiles_x = cdiv(a.wo, 8); }
    else { t.tiles_y = cdiv((long long)a.h * a.w, 128); t.tiles_x = 1; }
    t.ntile_n = cdiv(a.cout, bn);
    t.wt_static = (a.wt_static && a.wt_img_stride == 0) ? 1 : 0;
    // split-K inside a thread-block cluster (<= 8 CTAs, portable size) whenever the layer splits at all: the partial
    // tiles meet in distributed shared memory instead of L2 and no second kernel is needed (opt-in: KEEP_TC_CLUSTER=8; measured slower than the two-kernel path, profiles/r1_experiments.md;
    // the value caps the cluster size)
    static const int cluster_max = std::min(8, std::max(0, env_int("KEEP_TC_CLUSTER", 0)));
    const bool want_cluster = splitk > 1 && cluster_max >= 2 && bn % 32 == 0;
    if (want_cluster && splitk > cluster_max) splitk = cluster_max;
    {   // every K-split must own at least one channel block
        const int cb_per = cdiv(t.ncb, splitk < 1 ? 1 : splitk);
        splitk = cdiv(t.ncb, cb_per);
    }
    t.cluster_k = (want_cluster && splitk > 1) ? 1 : 0;
    t.splitk = splitk; t.partial = partial;
    t.act = a.act; t.res = a.res; t.res_dt = a.res_dt; t.out = a.out; t.out_dt = a.out_dt;
    t.gn_part = nullptr; t.gn_cpg = 0; t.gn_P = 0; t.gn_cpg_shift = 0;
    if (a.gn_part && splitk == 1) {   // (split layers: the reduce kernel emits the statistics)
        KEEP_CHECK(!t.cluster_k && a.cout % 32 == 0 && a.gn_P == t.tiles_y * t.tiles_x * 4, "conv2d_tc: bad GroupNorm statistics request");
        t.gn_part = a.gn_part; t.gn_cpg = a.cout / 32; t.gn_P = a.gn_P;
        while ((1 << t.gn_cpg_shift) < t.gn_cpg) ++t.gn_cpg_shift;
        KEEP_CHECK(t.gn_cpg == 2 || t.gn_cpg == 4 || t.gn_cpg == 8 || t.gn_cpg == 16, "conv2d_tc: GroupNorm statistics need 64..512 channels");
    }
    t.M = (long long)a.n * a.ho * a.wo;
    KEEP_CHECK((long long)a.n * a.h * a.w < (1ll << 31) && t.M < (1ll << 31), "conv2d_tc: more than 2^31 pixels");
    int cols = 32;
    while (cols < 2 * (tc_stacked(bn, passes) ? 128 : bn)) cols *= 2;   // (stacked panels: 128 accumulator columns per 64-wide tile)
    KEEP_CHECK(cols <= 512, "conv2d_tc: BN %d needs more than 512 TMEM columns", bn);
    t.tmem_cols = cols;
    t.swap_lbo_sbo = env_swap();
    t.w_resident = pick_resident(t, splitk) ? 1 : 0;
    t.a_stat = (!t.w_resident && !t.cluster_k && pick_a_stationary(t, splitk, num_sms > 0 ? num_sms : 148)) ? 1 : 0;
    const size_t resident_bytes = t.w_resident ? (size_t)t.ncb * t.taps * bn * 128 : 0;
    // A-stationary: room for the stages of two items when that fits, so the next item is produced while this one is consumed
    const int cb_per_item = cdiv(t.ncb, splitk);
    pick_stages(t.win, bn, resident_bytes, t.a_stat ? cb_per_item : 0, t.sa_stages, t.sb_stages);
    if (t.a_stat) {
        const int a_stage = a_stage_bytes(1), b_stage = bn * 128;
        while (t.sa_stages < std::min(MAX_SA, 2 * cb_per_item) && (t.sa_stages + 1) * a_stage + 3 * b_stage <= SMEM_BUDGET) {
            ++t.sa_stages;
            while (t.sb_stages > 3 && t.sa_stages * a_stage + t.sb_stages * b_stage > SMEM_BUDGET) --t.sb_stages;
        }
    }
    t.trace = g_tc_trace;
    if (t.cluster_k && (size_t)t.sa_stages * a_stage_bytes(t.win) < (size_t)128 * bn * 4) t.cluster_k = 0;   // partial tile must fit the A stages
    KEEP_CHECK(splitk == 1 || t.cluster_k || partial, "conv2d_tc: split-K needs a partial buffer");
    const size_t smem = SMEM_FIXED + (size_t)t.sa_stages * a_stage_bytes(t.win) + (t.w_resident ? resident_bytes : (size_t)t.sb_stages * bn * 128);
    KEEP_CHECK(smem <= 227 * 1024, "conv2d_tc: %zu bytes of shared memory", smem);
    KEEP_CHECK(a.c1 == 0 || a.in0_dt == a.in1_dt, "conv2d_tc: concatenated sources must share a dtype");
    const long long total = (long long)t.n * t.tiles_y * t.tiles_x * splitk * (t.a_stat ? 1 : t.ntile_n);
    KEEP_CHECK(total < (1ll << 30), "conv2d_tc: too many work items");
    t.cb_per = cdiv(t.ncb, splitk);
    t.fd_ntile = make_fastdiv((uint32_t)t.ntile_n);
    t.fd_mtiles = make_fastdiv((uint32_t)(t.n * t.tiles_y * t.tiles_x));
    t.fd_mtimg = make_fastdiv((uint32_t)(t.tiles_y * t.tiles_x));
    t.fd_tilesx = make_fastdiv((uint32_t)t.tiles_x);
    t.fd_splitk = make_fastdiv((uint32_t)splitk);
    // num_sms > 0: persistent grid capped at that many CTAs.  num_sms < 0 (low-priority side branch): short-lived CTAs
    // of about -num_sms work items each and as many of them as that takes -- they soak up whatever SMs the
    // latency-critical main stream leaves idle and hand an SM back within a few microseconds when it wants one.
    const int grid = num_sms > 0 ? (int)std::min<long long>(total, num_sms)
                                 : (int)std::min<long long>(total, std::max<long long>(1, (total + (-num_sms) - 1) / (-num_sms)));
    const bool f16 = a.in0_dt == F16;
    tc_configure_device();
    const bool wide = a.a_wide && passes == 3 && !f16;
    const Kern kern = wide ? kerns_wide[t.win - 1] : kerns[passes == 3 ? 1 : 0][f16 ? 1 : 0][t.win - 1];
    if (t.cluster_k) {   // one work item per CTA, the splitk CTAs of a tile form one cluster
        launch_k_cluster(kern, dim3((unsigned)total), dim3(tc_threads(t.win)), smem, s, splitk, t);
        CUDA_CHECK(cudaGetLastError());
        return 1;
    }
    launch_k(kern, dim3(grid), dim3(tc_threads(t.win)), smem, s, t);
    CUDA_CHECK(cudaGetLastError());
    if (a.no_reduce) {
        KEEP_CHECK(splitk == a.splitk && !a.bias && !a.res && a.act == ACT_NONE, "conv2d_tc: no_reduce needs the K split as requested and a plain epilogue");
        return 1;
    }
    if (splitk > 1 && a.ln_out) {
        KEEP_CHECK(a.out_dt == F32 && !a.gn_part, "conv2d_tc: the reduce + LayerNorm kernel writes fp32 and no GroupNorm statistics");
        splitk_reduce_ln(partial, splitk, t.M * a.cout, a.cout, a.bias, a.act, a.res, a.res_dt, (float*)a.out, a.ln_g, a.ln_b, a.ln_eps, a.ln_out,
                         a.ln_add2, a.ln_add2_rows, a.ln_out2, s);
    } else if (splitk > 1)
        splitk_reduce(partial, splitk, t.M * a.cout, a.cout, a.bias, a.act, a.res, a.res_dt, a.out, a.out_dt, s, a.gn_part, a.gn_P,
                      (long long)a.ho * a.wo, a.gn_fin_gamma, a.gn_fin_beta, a.gn_fin_scale, a.gn_fin_shift, a.gn_tickets);
    else
        KEEP_CHECK(!a.gn_fin_scale && !a.ln_out, "conv2d_tc: finalize-in-reduce / LayerNorm-in-reduce requested for a layer without a split-K reduce");
    return splitk > 1 ? 2 : 1;
}

KEEP_STAMP_SETTER(stamp_set_conv_tc)

}  // namespace keep

// op-level test hook (capi.cu): pack on the fly, run, free.  use_tc: 1 = fp16 operands, 3 = split precision
int keepop_conv2d_tc(const keep::ConvArgs& a_in, const float* w_oihw_host, int passes, cudaStream_t s, const float* gn_gamma,
                     const float* gn_beta, float* gn_scale, float* gn_shift) {
    using namespace keep;
    ConvArgs a = a_in;
    const int cin = a.c0 + a.c1;
    KEEP_CHECK(tc_eligible(a), "keepop_conv2d(use_tc): layer not eligible for the tcgen05 kernel");
    const long long m_tiles = a.kh == 3 ? (long long)a.n * cdiv(a.ho, 16) * cdiv(a.wo, 8) : (long long)a.n * cdiv((long long)a.h * a.w, 128);
    const int bn = tc_pick_bn(a.cout, m_tiles, passes);
    const bool s2d = tc_is_s2d(a);
    const int vcin = tc_virtual_cin(a, passes);
    __half* dw = nullptr;
    float* part = nullptr;
    if (!s2d) {
        std::vector<__half> packed(tc_packed_weight_halfs(cin, a.cout, a.kh * a.kw, bn, passes));
        tc_pack_weights(w_oihw_host, a.cout, cin, a.kh, a.kw, bn, passes, packed.data(), (a.a_wide && passes == 3) ? 1 : 0);
        CUDA_CHECK(cudaMalloc((void**)&dw, packed.size() * sizeof(__half)));
        CUDA_CHECK(cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
    } else {   // a.wt already holds the [9*cin][cout] fp32 layout on the device
        CUDA_CHECK(cudaMalloc((void**)&dw, tc_packed_weight_halfs(vcin, a.cout, 4, bn, passes) * sizeof(__half)));
        tc_repack_device(a.wt, cin, a.cout, 9, bn, passes, a.pad_t, dw, s, (a.a_wide && passes == 3) ? 1 : 0);
    }
    const int splitk = tc_pick_splitk(m_tiles, cdiv(a.cout, bn), cdiv(vcin, tc_cb(passes)));
    if (splitk > 1) CUDA_CHECK(cudaMalloc((void**)&part, (size_t)splitk * a.n * a.ho * a.wo * a.cout * sizeof(float)));
    float* gn_part = nullptr;
    int* tickets = nullptr;
    if (gn_scale) {   // GroupNorm(32) statistics of the output from the producing kernel + the finalize launch
        a.out_dt = F32;
        const int P = conv_gn_slots(a, splitk);
        KEEP_CHECK(P > 0, "keepop_conv2d_gn: this layer cannot emit GroupNorm statistics");
        CUDA_CHECK(cudaMalloc((void**)&gn_part, (size_t)a.n * P * 64 * sizeof(float)));
        CUDA_CHECK(cudaMemsetAsync(gn_part, 0xff, (size_t)a.n * P * 64 * sizeof(float), s));   // NaN pattern: every slot must be written
        a.gn_part = gn_part; a.gn_P = P;
        static const bool fin_en = !(getenv("KEEP_GN_REDUCE_FINAL") && getenv("KEEP_GN_REDUCE_FINAL")[0] == '0');
        if (fin_en && splitk > 1 && P <= kGnReduceFinalMaxP) {   // the engine's choice for such layers: finalize inside the reduce
            CUDA_CHECK(cudaMalloc((void**)&tickets, a.n * sizeof(int)));
            CUDA_CHECK(cudaMemsetAsync(tickets, 0, a.n * sizeof(int), s));
            a.gn_fin_gamma = gn_gamma; a.gn_fin_beta = gn_beta; a.gn_fin_scale = gn_scale; a.gn_fin_shift = gn_shift; a.gn_tickets = tickets;
        }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    try {
        CUDA_CHECK(cudaStreamSynchronize(s));   // panels are complete: the loader may run ahead of griddepcontrol.wait
        a.wt_static = 1;
        conv2d_tc(a, dw, bn, passes, splitk, part, sms, s);
        if (gn_part && !a.gn_fin_scale) gn_finalize_parts(gn_part, a.n, a.gn_P, a.ho * a.wo, a.cout, 1e-6f, gn_gamma, gn_beta, gn_scale, gn_shift, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
    } catch (...) {
        cudaFree(dw);
        cudaFree(part);
        cudaFree(gn_part);
        cudaFree(tickets);
        throw;
    }
    cudaFree(dw);
    cudaFree(part);
    cudaFree(gn_part);
    cudaFree(tickets);
    return 0;
}
