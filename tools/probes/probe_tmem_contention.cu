// Probe: does tcgen05.ld traffic (epilogue warps) slow tcgen05.mma down, and the other way round?  One CTA: thread 0 issues a
// stream of M=128, N=80, K=16 MMAs (A from tensor memory or from shared memory) while 8 warps loop over tcgen05.ld.32x32b.x16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tmem_contention probe_tmem_contention.cu
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int N = 80, MMAS = 360, LDS_PER_WARP = 200;
template <bool TS, bool DO_MMA, bool DO_LD>
__global__ void __launch_bounds__(384, 1) k_cont(long long *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 3 * 1024; i += 384) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp < 4) {
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 496;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && DO_MMA) {
        const uint64_t da = desc_nosw(smem_u32(sm), 2048, 128), db = desc_nosw(smem_u32(sm + 4096), N * 16, 128);
        const long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < MMAS; i++) {
            const uint32_t d = tmem + (i & 1) * N;                     // accumulators at columns 0..159
            if (TS)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(tmem + 496), "l"(db), "r"(idesc_bf16(128, N)), "r"(i >= 2 ? 1u : 0u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc_bf16(128, N)), "r"(i >= 2 ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        out[0] = clock64() - t0;
    }
    if (warp >= 4 && DO_LD) {
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < LDS_PER_WARP; i++) {
            uint32_t v[16];
            // other columns than the accumulators: 192 + 16*(i%16)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 192 + 16 * (i & 15)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] ^ v[15];
        }
        const long long t1 = clock64();
        if (lane == 0) out[warp] = t1 - t0 + (acc == 0x12345678u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <bool TS, bool DO_MMA, bool DO_LD>
static void run(long long *dout, const char *name) {
    cudaMemset(dout, 0, 16 * 8);
    cudaFuncSetAttribute(k_cont<TS, DO_MMA, DO_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k_cont<TS, DO_MMA, DO_LD><<<1, 384, 16384>>>(dout);
    long long h[16];
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s", name);
    if (DO_MMA) printf(" %6.1f cycles per MMA (%d MMAs);", (double)h[0] / MMAS, MMAS);
    if (DO_LD) printf(" %6.1f cycles per tcgen05.ld.x16+wait per warp (8 warps: %.0f B/cycle)", (double)h[4] / LDS_PER_WARP, 8.0 * 2048 * LDS_PER_WARP / h[4]);
    printf("\n");
}

int main() {
    long long *dout;
    cudaMalloc(&dout, 16 * 8);
    run<true, true, false>(dout, "MMA A=tmem alone");
    run<false, true, false>(dout, "MMA A=smem alone");
    run<true, false, true>(dout, "8 warps of tcgen05.ld alone");
    run<true, true, true>(dout, "MMA A=tmem + 8 warps tcgen05.ld");
    run<false, true, true>(dout, "MMA A=smem + 8 warps tcgen05.ld");
    return 0;
}
