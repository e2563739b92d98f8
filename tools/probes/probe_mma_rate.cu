// Probe: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, of where A comes from (shared memory / tensor
// memory) and of how many independent accumulators consecutive MMAs rotate over.  One CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mma_rate probe_mma_rate.cu && ./probe_mma_rate
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int COUNT = 96;
template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k_rate(long long *out) {
    extern __shared__ __align__(128) uint8_t sm[];     // A: 4 KB at 0, B: up to 8 KB at 4096 (zeros)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 3 * 1024; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
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
    {   // zero the A columns (496..503) of every lane
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 496;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t da = desc_nosw(smem_u32(sm), 2048, 128), db = desc_nosw(smem_u32(sm + 4096), N * 16, 128);
        for (int rep = 0; rep < 3; rep++) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < COUNT; i++) {
                const uint32_t d = tmem + (i % NACC) * N;
                if (TS)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(tmem + 496), "l"(db), "r"(idesc_bf16(128, N)), "r"(i >= NACC ? 1u : 0u) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(da), "l"(db), "r"(idesc_bf16(128, N)), "r"(i >= NACC ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
            out[rep] = clock64() - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool TS, int NACC>
static void run(long long *dout) {
    cudaFuncSetAttribute(k_rate<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k_rate<N, TS, NACC><<<1, 128, 16384>>>(dout);
    long long h[3];
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d A=%s accumulators=%d: %6.1f cycles per MMA (N/2 = %d)\n", N, TS ? "tmem" : "smem", NACC, (double)h[2] / COUNT, N / 2);
}

int main() {
    long long *dout;
    cudaMalloc(&dout, 64);
    run<32, false, 1>(dout);  run<32, false, 2>(dout);  run<32, false, 4>(dout);
    run<64, false, 1>(dout);  run<64, false, 2>(dout);
    run<80, false, 1>(dout);  run<80, false, 2>(dout);  run<80, false, 3>(dout);
    run<80, true, 1>(dout);   run<80, true, 2>(dout);   run<80, true, 3>(dout);
    run<96, true, 1>(dout);   run<96, true, 2>(dout);
    run<128, true, 1>(dout);  run<128, true, 2>(dout);  run<128, false, 1>(dout);
    run<160, true, 1>(dout);  run<160, true, 2>(dout);
    run<240, true, 1>(dout);  run<240, true, 2>(dout);  run<256, false, 1>(dout);
    return 0;
}
