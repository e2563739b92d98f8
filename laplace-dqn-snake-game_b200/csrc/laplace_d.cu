// laplace_d.cu — Laplace deviation-matrix kernels (compute_D.jl:9-31, 66-81; la_utils.jl:14-36, 154-169).
//
// D is P x K Float64 column-major: column k is the k-th snapshot of the flattened Q-net weights.
// k_center_columns restates, per row p, the reference's streaming statistics
//     for c in eachcol(D): n += 1; d = x - mean; mean += d / n; m2 += d * (x - mean)
// followed by `D .-= mean` and var = m2 ./ max(n-1, 1), in Float64 with explicit round-to-nearest
// intrinsics so that no multiply-add is contracted: the results are bit-identical to the Julia loop.
// The column order is a true sequential dependency, so the parallel axis is p (coalesced: thread p reads
// D[p + k P]); 8 columns are loaded ahead of the dependent arithmetic to keep HBM requests in flight.
#include "common.h"

namespace snk {

constexpr int CENTER_TPB = 128;
constexpr int CENTER_AHEAD = 8;

__global__ void __launch_bounds__(CENTER_TPB) k_center_columns(double *__restrict__ D, long long P, long long K,
                                                               double *__restrict__ mean_out,
                                                               double *__restrict__ var_out) {
    const long long p = (long long)blockIdx.x * CENTER_TPB + threadIdx.x;
    if (p >= P) return;
    double mean = 0.0, m2 = 0.0;
    long long k = 0;
    for (; k + CENTER_AHEAD <= K; k += CENTER_AHEAD) {
        double x[CENTER_AHEAD];
#pragma unroll
        for (int j = 0; j < CENTER_AHEAD; j++) x[j] = __ldcs(D + (k + j) * P + p);
#pragma unroll
        for (int j = 0; j < CENTER_AHEAD; j++) {
            double d = __dsub_rn(x[j], mean);
            mean = __dadd_rn(mean, __ddiv_rn(d, (double)(k + j + 1)));
            m2 = __dadd_rn(m2, __dmul_rn(d, __dsub_rn(x[j], mean)));
        }
    }
    for (; k < K; k++) {
        double x = __ldcs(D + k * P + p);
        double d = __dsub_rn(x, mean);
        mean = __dadd_rn(mean, __ddiv_rn(d, (double)(k + 1)));
        m2 = __dadd_rn(m2, __dmul_rn(d, __dsub_rn(x, mean)));
    }
    if (mean_out != nullptr) mean_out[p] = mean;
    if (var_out != nullptr) var_out[p] = __ddiv_rn(m2, (double)(K - 1 > 1 ? K - 1 : 1));
    k = 0;
    for (; k + CENTER_AHEAD <= K; k += CENTER_AHEAD) {
        double x[CENTER_AHEAD];
#pragma unroll
        for (int j = 0; j < CENTER_AHEAD; j++) x[j] = __ldcs(D + (k + j) * P + p);
#pragma unroll
        for (int j = 0; j < CENTER_AHEAD; j++) D[(k + j) * P + p] = __dsub_rn(x[j], mean);
    }
    for (; k < K; k++) D[k * P + p] = __dsub_rn(D[k * P + p], mean);
}

// deviation_matrix[:, position] = Float64.(theta)   (compute_D.jl:67-71, la_utils.jl:154-158)
__global__ void k_store_snapshot(double *__restrict__ D, long long P, long long k, const float *__restrict__ theta) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) D[k * P + p] = (double)theta[p];
}

// sample_model (la_utils.jl:83-95):  w = mean + 1/sqrt(2) * sqrt.(Gamma_diag) * z1 + 1/sqrt(2(K-1)) * D * z2
// with Gamma_diag = |var| (compute_Gamma_diag, la_utils.jl:74-81).  One thread per parameter p; D z2 is accumulated in
// column order in Float64 (the reference's BLAS gemv order is unspecified: parity is to ~1e-15 relative, not bit-exact).
__global__ void k_laplace_sample(const double *__restrict__ mean, const double *__restrict__ var, const double *__restrict__ D,
                                 long long P, long long K, const double *__restrict__ z1, const double *__restrict__ z2,
                                 double *__restrict__ w) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double acc = 0.0;
    for (long long k = 0; k < K; k++) acc = fma(D[k * P + p], z2[k], acc);
    const double g = fabs(var[p]);
    w[p] = mean[p] + (1.0 / sqrt(2.0)) * sqrt(g) * z1[p] + (1.0 / sqrt(2.0 * (double)(K - 1))) * acc;
}

}  // namespace snk

using namespace snk;

extern "C" int snk_laplace_sample_weights(const double *mean, const double *var, const double *D, int64_t P, int64_t K,
                                          const double *z1, const double *z2, double *w, void *cuda_stream) {
    SNK_REQUIRE(mean && var && D && z1 && z2 && w, "null argument");
    SNK_REQUIRE(P > 0 && K > 1, "need P > 0 and K > 1");
    k_laplace_sample<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(mean, var, D, P, K, z1, z2, w);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}

extern "C" int snk_d_store_snapshot(double *D, int64_t P, int64_t K, int64_t position, const float *theta, void *cuda_stream) {
    SNK_REQUIRE(D != nullptr && theta != nullptr, "null argument");
    SNK_REQUIRE(P > 0 && position >= 0 && position < K, "position must be in 0..K-1 (0-based column)");
    k_store_snapshot<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(D, P, position, theta);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}

extern "C" int snk_center_columns(double *D, int64_t P, int64_t K, double *mean, double *var, void *cuda_stream) {
    SNK_REQUIRE(D != nullptr, "null D");
    SNK_REQUIRE(P > 0 && K > 0, "P and K must be positive");
    unsigned grid = (unsigned)((P + CENTER_TPB - 1) / CENTER_TPB);
    k_center_columns<<<grid, CENTER_TPB, 0, (cudaStream_t)cuda_stream>>>(D, P, K, mean, var);
    SNK_CUDA(cudaGetLastError());
    return SNK_OK;
}
