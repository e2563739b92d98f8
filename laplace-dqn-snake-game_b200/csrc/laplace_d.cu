// laplace_d.cu — Laplace deviation-matrix kernels (compute_D.jl:9-31, 66-81; la_utils.jl:14-36, 154-169).
//
// D is P x K Float64 column-major: column k is the k-th snapshot of the flattened Q-net weights.
// k_center_columns restates, per row p, the reference's streaming statistics
//     for c in eachcol(D): n += 1; d = x - mean; mean += d / n; m2 += d * (x - mean)
// followed by `D .-= mean` and var = m2 ./ max(n-1, 1), in Float64 with explicit round-to-nearest
// intrinsics so that no multiply-add is contracted: the results are bit-identical to the Julia loop.
// The column order is a true sequential dependency, so the parallel axis is p (coalesced: thread p reads
// D[p + k P]); 16 columns per thread are kept in flight through a shared-memory cp.async ring.  The per-column
// division by the running count uses a shared table of correctly rounded reciprocals plus one exact-remainder
// correction (div_by_count) instead of the ~20-instruction IEEE division sequence.
#include "common.h"

namespace snk {

#ifndef SNK_CENTER_TPB
#define SNK_CENTER_TPB 128
#endif
#ifndef SNK_CENTER_DEPTH
#define SNK_CENTER_DEPTH 16
#endif
#ifndef SNK_CENTER_GROUP
#define SNK_CENTER_GROUP 2
#endif
#ifndef SNK_CENTER_CARVE
#define SNK_CENTER_CARVE (-1)   /* percent of the L1/shared array given to shared memory; -1 = driver default */
#endif
#ifndef SNK_CENTER_MINB
#define SNK_CENTER_MINB 7    /* 72 registers, no spills: measured faster than 10 CTAs/SM at 48 registers with spills */
#endif
constexpr int CENTER_TPB = SNK_CENTER_TPB;
constexpr int CENTER_DEPTH = SNK_CENTER_DEPTH;   // columns in flight per thread (16 KB of ring per CTA)
constexpr int CENTER_GROUP = SNK_CENTER_GROUP;    // columns per cp.async commit group
constexpr long long CENTER_FAST_MAX_K = 1ll << 20;

// d / n for an integer count n <= 2^20 with y = RN(1/n), bit-identical to the IEEE quotient:
//   q = RN(d y);  r = d - n q (exact: a multiple of ulp(q) below 2^12 ulps, one fma);  result = RN(q + r y).
// q + r y differs from the true quotient Q = q + r/n by less than 2^-52 ulp(Q), while Q = d/n is never a rounding
// midpoint (an odd 54-bit significand times n does not fit 53 bits) and lies at least ulp/(2n) >= 2^-21 ulp away from
// one, so the last rounding returns RN(Q).  The argument needs r and r y free of underflow/overflow, hence the
// magnitude window; zeros, subnormals, huge values, Inf and NaN take the IEEE division.
__device__ __forceinline__ double div_by_count(double d, double n, double y) {
    const double ad = fabs(d);
    if (ad > 0x1p-900 && ad < 0x1p+900) {
        const double q = __dmul_rn(d, y);
        const double r = __fma_rn(-q, n, d);
        return __fma_rn(r, y, q);
    }
    return __ddiv_rn(d, n);
}

// Column ring in shared memory: every thread keeps CENTER_DEPTH columns of its own row in flight with 8-byte cp.async
// copies (rows of odd length leave columns only 8-byte aligned, which rules out 16-byte and bulk/TMA copies), one
// commit group per CENTER_GROUP columns.  A thread only ever reads the slots it filled itself, so the ring needs no
// block barrier and the register file holds just the running statistics.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Streams row p of D through the ring: for every block of CENTER_DEPTH columns starting at kb calls begin(kb) once and
// then body(kb, c, x) for c = 0..CENTER_DEPTH-1 (while kb + c < K) in column order, x = D[p + (kb + c) P].  The block
// is fully unrolled, so ring slots are compile-time offsets and the tail tests are uniform integer compares.
template <typename Begin, typename Body>
__device__ __forceinline__ void stream_row(const double *__restrict__ D, long long P, long long K, long long p,
                                           double (*ring)[CENTER_TPB], Begin begin, Body body) {
    constexpr int NG = CENTER_DEPTH / CENTER_GROUP;
    static_assert(CENTER_DEPTH % CENTER_GROUP == 0, "ring depth must be a whole number of commit groups");
    double *mine = &ring[0][threadIdx.x];        // slot s of this thread = mine[s * CENTER_TPB]
    const double *src = D + p;                   // next column to request (runs ahead past the end, never dereferenced there)
#pragma unroll
    for (int g = 0; g < NG; g++) {
#pragma unroll
        for (int j = 0; j < CENTER_GROUP; j++) {
            const int c = g * CENTER_GROUP + j;
            if (c < K) cp_async8(mine + c * CENTER_TPB, src);
            src += P;
        }
        cp_async_commit();
    }
    for (long long kb = 0; kb < K; kb += CENTER_DEPTH) {
        const int rem = (int)(K - kb < 2 * CENTER_DEPTH ? K - kb : 2 * CENTER_DEPTH);   // columns from kb on, clamped
        begin(kb);
#pragma unroll
        for (int g = 0; g < NG; g++) {
            cp_async_wait<NG - 1>();
            double x[CENTER_GROUP];
#pragma unroll
            for (int j = 0; j < CENTER_GROUP; j++) x[j] = mine[(g * CENTER_GROUP + j) * CENTER_TPB];
#pragma unroll
            for (int j = 0; j < CENTER_GROUP; j++) {
                const int c = g * CENTER_GROUP + j;
                if (CENTER_DEPTH + c < rem) cp_async8(mine + c * CENTER_TPB, src);
                src += P;
            }
            cp_async_commit();
#pragma unroll
            for (int j = 0; j < CENTER_GROUP; j++) {
                const int c = g * CENTER_GROUP + j;
                if (c < rem) body(kb, c, x[j]);
            }
        }
    }
    cp_async_wait<0>();
}

template <bool FAST>
__global__ void __launch_bounds__(CENTER_TPB, SNK_CENTER_MINB) k_center_columns(double *__restrict__ D, long long P, long long K,
                                                               double *__restrict__ mean_out,
                                                               double *__restrict__ var_out) {
    static_assert(CENTER_TPB % CENTER_DEPTH == 0, "the reciprocal table is refreshed on ring-block boundaries");
    __shared__ double ring[CENTER_DEPTH][CENTER_TPB];
    __shared__ double rcp[CENTER_TPB];          // RN(1/n) for the CENTER_TPB columns being processed
    const long long p_raw = (long long)blockIdx.x * CENTER_TPB + threadIdx.x;
    const bool active = p_raw < P;
    const long long p = active ? p_raw : P - 1;  // idle lanes shadow the last row (loads only) so the block barriers stay uniform
    double mean = 0.0, m2 = 0.0, n = 0.0;       // n: running column count (exact in Float64)
    int kt = 0;                                  // first table entry of the current ring block
    stream_row(D, P, K, p, ring,
        [&](long long kb) {
            kt = (int)(kb % CENTER_TPB);
            if (FAST && kt == 0) {               // uniform across the block: every thread walks the columns in step
                __syncthreads();
                rcp[threadIdx.x] = __drcp_rn((double)(kb + threadIdx.x + 1));
                __syncthreads();
            }
        },
        [&](long long, int c, double x) {
            n += 1.0;
            const double d = __dsub_rn(x, mean);
            mean = __dadd_rn(mean, FAST ? div_by_count(d, n, rcp[kt + c]) : __ddiv_rn(d, n));
            m2 = __dadd_rn(m2, __dmul_rn(d, __dsub_rn(x, mean)));
        });
    if (active) {
        if (mean_out != nullptr) mean_out[p] = mean;
        if (var_out != nullptr) var_out[p] = __ddiv_rn(m2, (double)(K - 1 > 1 ? K - 1 : 1));
    }
    double *dst = D + p;
    stream_row(D, P, K, p, ring, [](long long) {},
        [&](long long, int, double x) {
            if (active) *dst = __dsub_rn(x, mean);
            dst += P;
        });
}

// deviation_matrix[:, position] = Float64.(theta)   (compute_D.jl:67-71, la_utils.jl:154-158)
__global__ void k_store_snapshot(double *__restrict__ D, long long P, long long k, const float *__restrict__ theta) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) D[k * P + p] = (double)theta[p];
}

// sample_model (la_utils.jl:83-95):  w = mean + 1/sqrt(2) * sqrt.(Gamma_diag) * z1 + 1/sqrt(2(K-1)) * D * z2
// with Gamma_diag = |var| (compute_Gamma_diag, la_utils.jl:74-81).  One thread per parameter p (row streamed through the
// same cp.async ring as k_center_columns: HBM-bound, D is read once); D z2 is accumulated in
// column order in Float64 (the reference's BLAS gemv order is unspecified: parity is to ~1e-15 relative, not bit-exact).
__global__ void __launch_bounds__(CENTER_TPB, SNK_CENTER_MINB) k_laplace_sample(const double *__restrict__ mean, const double *__restrict__ var,
                                                                                 const double *__restrict__ D, long long P, long long K,
                                                                                 const double *__restrict__ z1, const double *__restrict__ z2,
                                                                                 double *__restrict__ w) {
    __shared__ double ring[CENTER_DEPTH][CENTER_TPB];
    const long long p_raw = (long long)blockIdx.x * CENTER_TPB + threadIdx.x;
    const bool active = p_raw < P;
    const long long p = active ? p_raw : P - 1;
    double acc = 0.0;
    stream_row(D, P, K, p, ring, [](long long) {}, [&](long long kb, int c, double x) { acc = fma(x, __ldg(z2 + kb + c), acc); });
    if (!active) return;
    const double g = fabs(var[p]);
    w[p] = mean[p] + (1.0 / sqrt(2.0)) * sqrt(g) * z1[p] + (1.0 / sqrt(2.0 * (double)(K - 1))) * acc;
}

}  // namespace snk

using namespace snk;

extern "C" int snk_laplace_sample_weights(const double *mean, const double *var, const double *D, int64_t P, int64_t K,
                                          const double *z1, const double *z2, double *w, void *cuda_stream) {
    snk::DeviceGuard guard__(snk::device_of(w));
    SNK_REQUIRE(mean && var && D && z1 && z2 && w, "null argument");
    SNK_REQUIRE(P > 0 && K > 1, "need P > 0 and K > 1");
    static const bool carve = [] {
        cudaFuncSetAttribute(k_laplace_sample, cudaFuncAttributePreferredSharedMemoryCarveout, SNK_CENTER_CARVE);
        return true;
    }();
    (void)carve;
    k_laplace_sample<<<(unsigned)((P + CENTER_TPB - 1) / CENTER_TPB), CENTER_TPB, 0, (cudaStream_t)cuda_stream>>>(mean, var, D, P, K, z1, z2, w);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}

extern "C" int snk_d_store_snapshot(double *D, int64_t P, int64_t K, int64_t position, const float *theta, void *cuda_stream) {
    snk::DeviceGuard guard__(snk::device_of(D));
    SNK_REQUIRE(D != nullptr && theta != nullptr, "null argument");
    SNK_REQUIRE(P > 0 && position >= 0 && position < K, "position must be in 0..K-1 (0-based column)");
    k_store_snapshot<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(D, P, position, theta);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}

extern "C" int snk_center_columns(double *D, int64_t P, int64_t K, double *mean, double *var, void *cuda_stream) {
    snk::DeviceGuard guard__(snk::device_of(D));
    SNK_REQUIRE(D != nullptr, "null D");
    SNK_REQUIRE(P > 0 && K > 0, "P and K must be positive");
    unsigned grid = (unsigned)((P + CENTER_TPB - 1) / CENTER_TPB);
    // Carve-out left at the driver default (-1): measured 0.81 ms at the config-5a size for 50-60 % and the default alike;
    // the maximum carve-out is 25 % slower (the 8-byte-aligned columns share sectors between warps and want the L1).
    static const bool carve = [] {
        cudaFuncSetAttribute(k_center_columns<true>, cudaFuncAttributePreferredSharedMemoryCarveout, SNK_CENTER_CARVE);
        cudaFuncSetAttribute(k_center_columns<false>, cudaFuncAttributePreferredSharedMemoryCarveout, SNK_CENTER_CARVE);
        return true;
    }();
    (void)carve;
    if (K <= CENTER_FAST_MAX_K) k_center_columns<true><<<grid, CENTER_TPB, 0, (cudaStream_t)cuda_stream>>>(D, P, K, mean, var);
    else k_center_columns<false><<<grid, CENTER_TPB, 0, (cudaStream_t)cuda_stream>>>(D, P, K, mean, var);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}
