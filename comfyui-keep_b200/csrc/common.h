// keep_b200 — shared host/device helpers for the sm_100a KEEP engine.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <utility>
#include <stdexcept>

namespace keep {

// ------------------------------------------------------------------------------------------
// errors: everything throws; the C-ABI layer converts to an error code + thread-local message
// (the reference's nodes rely on catching exceptions, nodes.py:83-88 — never abort()).
// ------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define KEEP_CHECK(cond, ...)                                                         \
    do {                                                                              \
        if (!(cond)) {                                                                \
            char _b[512];                                                             \
            snprintf(_b, sizeof(_b), __VA_ARGS__);                                    \
            char _c[768];                                                             \
            snprintf(_c, sizeof(_c), "%s:%d: %s", __FILE__, __LINE__, _b);            \
            throw ::keep::Error(_c);                                                  \
        }                                                                             \
    } while (0)

#define CUDA_CHECK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            char _c[768];                                                             \
            snprintf(_c, sizeof(_c), "%s:%d: CUDA error %s: %s", __FILE__, __LINE__,  \
                     cudaGetErrorName(_e), cudaGetErrorString(_e));                   \
            throw ::keep::Error(_c);                                                  \
        }                                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------
// tensors: feature maps are NHWC; token matrices are (rows, C) == NHWC with h=rows, w=1
// ------------------------------------------------------------------------------------------
enum DType : int { F32 = 0, F16 = 1 };

static inline size_t dtype_size(DType d) { return d == F32 ? 4 : 2; }

struct Tensor {
    void* p = nullptr;
    int n = 0, h = 0, w = 0, c = 0;
    DType dt = F32;
    // GroupNorm(32) partial statistics of this tensor written by the kernel that produced it ([n][gn_P][32][2] fp32; see
    // ConvArgs::gn_part): consumed (and released) by the one Engine::gn() that normalises the tensor, else by tfree()
    float* gn_part = nullptr; int gn_P = 0;
    // ... or, for split-K layers with few slots, already finalized by the reduce kernel into the consumer's affine
    // ([2][n][c]: scale, shift; aff_gamma = the consuming norm's weight, checked by Engine::gn())
    float* aff = nullptr; const float* aff_gamma = nullptr;
    size_t numel() const { return (size_t)n * h * w * c; }
    size_t bytes() const { return numel() * dtype_size(dt); }
    size_t rows() const { return (size_t)n * h * w; }
    float* f() const { return (float*)p; }
    __half* hf() const { return (__half*)p; }
};

// activation applied in an epilogue / prologue
enum Act : int { ACT_NONE = 0, ACT_SWISH = 1, ACT_RELU = 2, ACT_LRELU02 = 3, ACT_GELU = 4, ACT_SIGMOID = 5 };

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float ldf(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __half* p, size_t i) { return __half2float(p[i]); }
__device__ __forceinline__ void stf(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(__half* p, size_t i, float v) { p[i] = __float2half_rn(v); }

// 4 consecutive elements (i must be a multiple of 4 and the row 16B/8B aligned)
__device__ __forceinline__ float4 ld4(const float* p, size_t i) { return *reinterpret_cast<const float4*>(p + i); }
__device__ __forceinline__ float4 ld4(const __half* p, size_t i) {
    uint2 u = *reinterpret_cast<const uint2*>(p + i);
    __half2 a = *reinterpret_cast<__half2*>(&u.x), b = *reinterpret_cast<__half2*>(&u.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, size_t i, float4 v) { *reinterpret_cast<float4*>(p + i) = v; }
__device__ __forceinline__ void st4(__half* p, size_t i, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + i) = u;
}

// 256-bit global accesses (sm_100: LDG / STG.E.ENL2.256): one full 32-byte sector per lane and instruction; p 32-byte aligned
__device__ __forceinline__ void ldg256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_SWISH: return v / (1.0f + expf(-v));                       // x * sigmoid(x)
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_LRELU02: return v > 0.0f ? v : 0.2f * v;
        case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // exact (erf) GELU
        case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Division of a 31-bit dividend by a run-time constant, magic numbers from the host: q = umulhi(n, mul) >> shr.
// (p = 31 + ceil(log2 d), mul = ceil(2^p / d) < 2^32, exact for n < 2^31.)  The kernels' work-item decoding used 64-bit
// divisions -- ~400 cycles of dependent subroutine each, several per role before the first load or MMA of every CTA.
struct FastDiv { uint32_t mul = 0, shr = 0, d = 1; };
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d < 1 ? 1 : d;
    if (f.d == 1) return f;
    uint32_t l = 0;
    while ((1ull << l) < f.d) ++l;
    const uint32_t pbits = 31 + l;
    f.mul = (uint32_t)(((1ull << pbits) + f.d - 1) / f.d);
    f.shr = pbits - 32;
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr); }
#endif

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: true exactly once per (call site's mask, current device),
// so a second engine on another GPU of the same process configures its own context (thread-safe)
bool first_use_on_current_device(unsigned long long* mask);   // engine.cu

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): ~560 small dependent kernels per frame make launch latency a first-order cost.
// Every kernel starts with pdl_prologue(): it lets the NEXT kernel's CTAs be scheduled right away (launch_dependents) and
// then waits until the PREVIOUS kernel has fully completed and flushed (wait), so correctness is that of plain stream
// order while CTA dispatch / prologue latency of kernel N+1 overlaps the tail of kernel N.  Both instructions are no-ops
// for a kernel launched without the attribute.
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Measured: triggering at the top parks the next kernel's CTAs (220 KB smem each for the tcgen05 kernel) on SMs the running
// kernel still needs (-21%).  So kernels only *wait*; the dependent launch is released implicitly as this grid's CTAs retire,
// which still overlaps the grid-boundary drain with the next kernel's dispatch.
// Experiment knob (KEEP_PDL_TRIGGER_MAX, default 0 = off): let grids of at most that many CTAs trigger at the top so the next
// kernel's prologue overlaps them.  Measured on B200 (profiles/r1_ab_pdl_side.md): 72 or 148 -> 87 frames/s vs 109 with
// wait-only -- the cascade of parked CTAs (each holding ~200 KB of shared memory) starves the low-priority GMFlow branch.
#ifndef KEEP_PDL_TRIGGER_MAX
#define KEEP_PDL_TRIGGER_MAX 0
#endif
__device__ __forceinline__ void pdl_early_trigger() {
    if (KEEP_PDL_TRIGGER_MAX > 0 && gridDim.x * gridDim.y * gridDim.z <= (unsigned)KEEP_PDL_TRIGGER_MAX) pdl_trigger();
}
// Debug timeline (tools/timeline.py): when a stamp buffer is set, the first thread of every kernel appends %globaltimer at the
// moment its grid may start (right after griddepcontrol.wait).  b[0] = running count, b[1..] = stamps.  One copy of the
// pointer per translation unit (no relocatable device code); KEEP_STAMP_SETTER defines the TU's setter.
constexpr unsigned long long KEEP_STAMP_CAP = 1ull << 16;
static __device__ unsigned long long* g_stamp_buf = nullptr;
__device__ __forceinline__ void keep_stamp() {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long* b = g_stamp_buf;
        if (b) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            const unsigned long long i = atomicAdd(b, 1ull);
            if (i < KEEP_STAMP_CAP) b[1 + i] = t;
        }
    }
}
#define KEEP_STAMP_SETTER(fn) \
    void fn(unsigned long long* p) { cudaMemcpyToSymbol(g_stamp_buf, &p, sizeof(p)); }
__device__ __forceinline__ void pdl_prologue() { pdl_early_trigger(); pdl_wait(); keep_stamp(); }
// Experiment knobs (both default OFF, measured on B200, profiles/r2_experiments.md): early `launch_dependents` from the short
// kernels between two tcgen05 convolutions (pdl_prologue_tiny: split-K reduce, GroupNorm finalize, LayerNorm, small softmax,
// ...) or from every non-convolution kernel (pdl_prologue_light).  The next convolution then runs its prologue (barrier
// init, TMEM allocation, cold instruction fetch, index setup, the weight loader's first TMA transfers) while the short
// kernel executes -- the convolution kernel waits per role for exactly that reason -- and the per-frame chain alone does get
// shorter.  But a parked convolution CTA holds a whole SM (200 KB of shared memory, 56K registers) doing nothing, and the
// clip as a whole is SM-time bound (GMFlow on the side stream fills every SM the chain leaves idle): 160.3 frames/s without
// any early trigger, 154.6 with the tiny set, 150.3 with the light set (which also parks convolution CTAs three kernels
// ahead: bgemm -> softmax -> bgemm -> conv).
#ifndef KEEP_PDL_TINY_TRIGGER
#define KEEP_PDL_TINY_TRIGGER 0
#endif
#ifndef KEEP_PDL_LIGHT_TRIGGER
#define KEEP_PDL_LIGHT_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_prologue_tiny() {
    if (KEEP_PDL_TINY_TRIGGER) pdl_trigger(); else pdl_early_trigger();
    pdl_wait();
    keep_stamp();
}
__device__ __forceinline__ void pdl_prologue_light() {
    if (KEEP_PDL_LIGHT_TRIGGER) pdl_trigger(); else pdl_early_trigger();
    pdl_wait();
    keep_stamp();
}
// stamp from one designated thread of block 0 (kernels whose thread 0 does not pass griddepcontrol.wait first)
__device__ __forceinline__ void keep_stamp_here() {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long* b = g_stamp_buf;
        if (b) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            const unsigned long long i = atomicAdd(b, 1ull);
            if (i < KEEP_STAMP_CAP) b[1 + i] = t;
        }
    }
}

bool pdl_enabled();   // engine.cu: KEEP_PDL env (default on)

void launch_log_add(const void* func, dim3 grid, dim3 block);   // engine.cu: no-op unless the debug launch log is on

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    launch_log_add(reinterpret_cast<const void*>(kernel), grid, block);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
// same, as a thread-block cluster of `cluster_x` consecutive CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args&&... args) {
    launch_log_add(reinterpret_cast<const void*>(kernel), grid, block);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
#endif

}  // namespace keep
