// takens.cu — audio side of the hot path before Rips: delay tau, Takens embedding + min-max
// normalisation, pairwise Euclidean distances (float64 arithmetic, float32 matrices out).
//
// Replaces:
//   compute_tau               /root/reference/scripts/utils.py:92-104  (first zero crossing of the autocorrelation)
//   takens_embedding          /root/reference/scripts/utils.py:107-116
//   compute_audio_persistence /root/reference/scripts/utils.py:123-130 (min-max normalisation)
//   ripser(point_cloud)       -> sklearn.metrics.pairwise_distances (Gram-trick Euclidean in float64,
//                                SURVEY.md Appendix A.1 step 2) followed by ripser's float32 cast
//
// K = 3: a tensor-core contraction would buy nothing — the kernels are
// HBM/latency bound; one warp per window, lanes along samples / points.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace takens {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kMaxDim = 8;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// one warp per window
__global__ void __launch_bounds__(128) tau_kernel(const double* __restrict__ wins, long long B, int L, long long stride,
                                                  int max_lag, int* __restrict__ tau) {
    extern __shared__ double tsm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* s = tsm + (size_t)wib * L;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long b = gw; b < B; b += nw) {
        const double* g = wins + b * stride;
        double sum = 0;
        for (int k = lane; k < L; k += 32) { double v = g[k]; s[k] = v; sum += v; }
        sum = wsum(sum);
        const double mean = sum / L;
        __syncwarp();
        for (int k = lane; k < L; k += 32) s[k] -= mean;
        __syncwarp();
        double a0 = 0;
        for (int k = lane; k < L; k += 32) a0 = fma(s[k], s[k], a0);
        a0 = wsum(a0) + 1e-10;
        int ml = max_lag < 0 ? L / 4 : max_lag;
        if (ml > L - 1) ml = L - 1;
        const int upper = ml < L ? ml : L;  // min(max_lag, len(ac))
        int res = ml / 10 > 1 ? ml / 10 : 1;
        for (int i = 1; i < upper; ++i) {
            double acc = 0;
            for (int k = lane; k + i < L; k += 32) acc = fma(s[k + i], s[k], acc);
            acc = wsum(acc);
            if (acc / a0 <= 0.0) { res = i; break; }
        }
        if (lane == 0) tau[b] = res;
        __syncwarp();
    }
}

// one warp per window: embed, subsample, min-max normalise -> pts (B, ldp, dim) float64, npts (B)
__global__ void __launch_bounds__(128) takens_kernel(const double* __restrict__ wins, long long B, int L,
                                                     long long stride, const int* __restrict__ tau, int tau_stride,
                                                     int dim, int sub, int ldp, double* __restrict__ pts,
                                                     int* __restrict__ npts, int normalise) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = gw; b < B; b += nw) {
        const double* g = wins + b * stride;
        const int t = tau[b * tau_stride];
        const int n_full = L - (dim - 1) * t;
        int n = n_full <= 0 ? 0 : (n_full + sub - 1) / sub;
        if (n > ldp) n = ldp;
        double mn[kMaxDim], rg[kMaxDim];
#pragma unroll
        for (int k = 0; k < kMaxDim; ++k) { mn[k] = 0.0; rg[k] = 1.0; }
        if (normalise && n > 0) {
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k) {
                if (k < dim) {
                    double lo = INFINITY, hi = -INFINITY;
                    for (int p = lane; p < n; p += 32) { double v = g[p * sub + k * t]; lo = fmin(lo, v); hi = fmax(hi, v); }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        lo = fmin(lo, __shfl_xor_sync(kFull, lo, o));
                        hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
                    }
                    mn[k] = lo;
                    rg[k] = (hi - lo == 0.0) ? 1.0 : hi - lo;
                }
            }
        }
        double* o = pts + b * (long long)ldp * dim;
        for (int p = lane; p < n; p += 32) {
#pragma unroll
            for (int k = 0; k < kMaxDim; ++k)
                if (k < dim) o[(size_t)p * dim + k] = (g[p * sub + k * t] - mn[k]) / rg[k];
        }
        if (lane == 0) npts[b] = n;
    }
}

// sklearn euclidean_distances semantics: d2 = (-2 x.y + |x|^2) + |y|^2, max(.,0), zero diagonal, sqrt
__device__ __forceinline__ float gram_dist(const double* __restrict__ a, const double* __restrict__ b, int dim,
                                           double na, double nb) {
    double dot = 0.0;
    for (int k = 0; k < dim; ++k) dot = fma(a[k], b[k], dot);
    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), na), nb);
    d2 = fmax(d2, 0.0);
    return (float)sqrt(d2);
}

// one CTA per cloud; D (B, ld, ld) float32, both triangles, zero diagonal; entries >= npts untouched
__global__ void __launch_bounds__(256) pairwise_kernel(const double* __restrict__ pts, const int* __restrict__ npts,
                                                       long long B, int ldp, int dim, int ld, float* __restrict__ D) {
    extern __shared__ double psm[];  // ldp*dim points + ldp norms
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const int n = npts ? npts[b] : ldp;
        const double* g = pts + b * (long long)ldp * dim;
        double* P = psm;
        double* nr = psm + (size_t)ldp * dim;
        for (int e = threadIdx.x; e < n * dim; e += blockDim.x) P[e] = g[e];
        __syncthreads();
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            double s = 0.0;
            for (int k = 0; k < dim; ++k) s = __dadd_rn(s, __dmul_rn(P[p * dim + k], P[p * dim + k]));
            nr[p] = s;
        }
        __syncthreads();
        float* Db = D + b * (long long)ld * ld;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e % n;
            Db[(size_t)i * ld + j] = (i == j) ? 0.0f : gram_dist(P + i * dim, P + j * dim, dim, nr[i], nr[j]);
        }
        __syncthreads();
    }
}

}  // namespace takens
}  // namespace tda

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

extern "C" int tda_compute_tau(const double* wins, long long B, int L, long long stride, int max_lag, int* tau,
                               void* stream) {
    if (!wins || !tau || B < 0 || L < 1) return TDA_E_ARG;
    if (B == 0) return 0;
    if (stride == 0) stride = L;
    const size_t smem = (size_t)4 * L * sizeof(double);
    if (smem > 200 * 1024) return TDA_E_SIZE;
    cudaFuncSetAttribute(tda::takens::tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long need = (B + 3) / 4;
    int grid = (int)(need < (long long)sm_count() * 8 ? need : (long long)sm_count() * 8);
    tda::ProfScope prof("compute_tau", (cudaStream_t)stream);
    tda::takens::tau_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(wins, B, L, stride, max_lag, tau);
    tda::count_launch();
    return (int)cudaGetLastError();
}

extern "C" int tda_takens_cloud(const double* wins, long long B, int L, long long stride, const int* tau,
                                int tau_stride, int dim, int subsample, int normalise, int ldp, double* pts,
                                int* npts, void* stream) {
    if (!wins || !tau || !pts || !npts || B < 0 || L < 1 || dim < 1 || dim > tda::takens::kMaxDim || subsample < 1 ||
        ldp < 1 || tau_stride < 0)
        return TDA_E_ARG;
    if (B == 0) return 0;
    if (stride == 0) stride = L;
    long long need = (B + 3) / 4;
    int grid = (int)(need < (long long)sm_count() * 16 ? need : (long long)sm_count() * 16);
    tda::ProfScope prof("takens_cloud", (cudaStream_t)stream);
    tda::takens::takens_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(wins, B, L, stride, tau, tau_stride, dim,
                                                                      subsample, ldp, pts, npts, normalise);
    tda::count_launch();
    return (int)cudaGetLastError();
}

extern "C" int tda_pairwise_dist_f32(const double* pts, const int* npts, long long B, int ldp, int dim, int ld, float* D,
                                     void* stream) {
    if (!pts || !D || B < 0 || ldp < 1 || dim < 1 || ld < ldp) return TDA_E_ARG;
    if (B == 0) return 0;
    const size_t smem = ((size_t)ldp * dim + ldp) * sizeof(double);
    if (smem > 200 * 1024) return TDA_E_SIZE;
    cudaFuncSetAttribute(tda::takens::pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = (int)(B < (long long)sm_count() * 8 ? B : (long long)sm_count() * 8);
    tda::ProfScope prof("pairwise_dist", (cudaStream_t)stream);
    tda::takens::pairwise_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pts, npts, B, ldp, dim, ld, D);
    tda::count_launch();
    return (int)cudaGetLastError();
}
