// Micro-benchmark: cycles per tcgen05.mma (M=128, N, K=16, kind::f16, SWIZZLE_128B K-major operands) as a function of how the
// A operand view is laid out in shared memory -- aligned 1024-byte swizzle atoms vs the row-shifted / odd-pitch "halo views"
// the implicit-GEMM convolution uses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

template <int N, int A_SBO, int PITCH, int WIN, int KSTEPS, int PASSES, int BP>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t a_hi = (((uint32_t)A_SBO >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        constexpr uint32_t b_hi = ((1024u >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
        const uint32_t a_base = (1u << 16) | (smem_u32(smem) >> 4);
        const uint32_t b_base = (1u << 16) | (smem_u32(smem + 64 * 1024) >> 4);
        uint32_t parity = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep)
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                constexpr int dummy = 0;
                const int dy = tap / WIN, dx = tap % WIN;
                const uint32_t a_lo = a_base + (uint32_t)(((dy * PITCH + dx) * 128) >> 4);
                const uint32_t b_lo = b_base + (uint32_t)(((tap % BP) * N * 128) >> 4);
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                    const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + kk * 2);
                    const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + kk * 2);
                    if (PASSES == 3) {
                        umma_f16(tmem, ad + 4, bd, idesc, (rep | tap | kk) ? 1u : 0u);
                        umma_f16(tmem, ad, bd + 4, idesc, 1u);
                        umma_f16(tmem, ad, bd, idesc, 1u);
                    } else {
                        umma_f16(tmem, ad, bd, idesc, (rep | tap | kk) ? 1u : 0u);
                    }
                }
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), parity);
            parity ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, int A_SBO, int PITCH, int WIN, int KSTEPS, int PASSES, int BP>
void run(const char* name, long long* d) {
    auto kern = k<N, A_SBO, PITCH, WIN, KSTEPS, PASSES, BP>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int iters = 100;
    for (int grid : {1, 148}) {
        cudaMemset(d, 0, 8);
        kern<<<grid, 128, 210 * 1024>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const long long mmas = 2ll * iters * 9 * KSTEPS * (PASSES == 3 ? 3 : 1);
        printf("grid %3d  %-52s %8.1f clk/MMA  (math %d)  %s\n", grid, name, (double)cyc / mmas, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    // <N, A_SBO bytes between 8-row groups, PITCH rows between dy views, WIN, KSTEPS, PASSES, B panels cycled>
    run<64, 1024, 0, 3, 4, 1, 9>("N64  aligned atoms, one view", d);
    run<64, 1280, 10, 3, 4, 1, 9>("N64  halo pitch 10, 3x3 shifted views (SBO 1280)", d);
    run<64, 2048, 16, 3, 4, 1, 9>("N64  halo pitch 16, 3x3 shifted views (SBO 2048)", d);
    run<64, 2048, 16, 1, 4, 1, 9>("N64  pitch 16, dx=0 views only", d);
    run<64, 1024, 8, 1, 4, 1, 9>("N64  aligned views (SBO 1024, dy*8 rows)", d);
    run<64, 1024, 0, 3, 4, 1, 1>("N64  aligned, same B panel", d);
    run<128, 1024, 0, 3, 4, 1, 9>("N128 aligned atoms", d);
    run<128, 1280, 10, 3, 4, 1, 9>("N128 halo pitch 10 shifted views", d);
    run<256, 1024, 0, 3, 4, 1, 3>("N256 aligned atoms", d);
    run<256, 1280, 10, 3, 4, 1, 3>("N256 halo pitch 10 shifted views", d);
    run<64, 1024, 0, 3, 2, 3, 9>("N64  tc3 interleaved, aligned", d);
    run<64, 1280, 10, 3, 2, 3, 9>("N64  tc3 interleaved, halo pitch 10", d);
    run<128, 1024, 0, 3, 2, 3, 9>("N128 tc3 interleaved, aligned", d);
    run<128, 1280, 10, 3, 2, 3, 9>("N128 tc3 interleaved, halo pitch 10", d);
    return 0;
}
