// Probe: tcgen05.mma with the A operand in tensor memory (written with tcgen05.st.32x32b), B from shared memory.
// Confirms the layout engine 17 of qnet.cu relies on: TMEM lane = row m, 32-bit column j holds K elements (2j, 2j+1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_ts_mma probe_ts_mma.cu && ./probe_ts_mma
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int N = 80;
// a: [128][16] bf16 row-major; b: [N][16] bf16 row-major (K-major); d: [128][N] f32
__global__ void __launch_bounds__(128, 1) k_probe(const __nv_bfloat16 *a, const __nv_bfloat16 *b, float *d) {
    __shared__ __align__(128) uint8_t sB[2 * N * 16];   // no-swizzle K-major: [K chunk][row][16 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * 2; i += 128) {            // (row n, chunk h)
        const int n = i >> 1, h = i & 1;
        *reinterpret_cast<uint4 *>(sB + h * N * 16 + n * 16) = *reinterpret_cast<const uint4 *>(b + n * 16 + h * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // A: thread = row m = tid, its 16 bf16 as 8 x b32 into columns 96..103 of its own lane
    const uint4 lo = *reinterpret_cast<const uint4 *>(a + tid * 16), hi = *reinterpret_cast<const uint4 *>(a + tid * 16 + 8);
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 96;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(ta), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t db = desc_nosw(smem_u32(sB), N * 16, 128);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(tmem + 96), "l"(db), "r"(idesc_bf16(128, N)), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) d[tid * N + c + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> a(128 * 16), b(N * 16);
    std::vector<float> af(128 * 16), bf(N * 16), d(128 * N);
    srand(1);
    for (size_t i = 0; i < a.size(); i++) { a[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); af[i] = __bfloat162float(a[i]); }
    for (size_t i = 0; i < b.size(); i++) { b[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); bf[i] = __bfloat162float(b[i]); }
    __nv_bfloat16 *da, *db; float *dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    k_probe<<<1, 128>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            float ref = 0;
            for (int k = 0; k < 16; k++) ref += af[m * 16 + k] * bf[n * 16 + k];
            worst = fmax(worst, fabs(ref - d[m * N + n]));
        }
    printf("probe_ts_mma: max |D - A B^T| = %g  (%s)\n", worst, worst < 1e-4 ? "OK: lane = row, column j = K elements 2j, 2j+1" : "MISMATCH");
    return worst < 1e-4 ? 0 : 2;
}
