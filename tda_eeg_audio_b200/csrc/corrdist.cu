// corrdist.cu — windows of a band-passed multichannel signal -> Pearson correlation -> Euclidean
// correlation distance d = sqrt(2(1-r)), float64 arithmetic, float32 distance matrices out.
//
// Replaces, per window:
//   create_sliding_windows      /root/reference/notebooks/1_preprocesamiento.ipynb:314-364 (slicing only)
//   compute_correlation_matrix  /root/reference/notebooks/2_graph_construction.ipynb:86-97 (np.corrcoef, NaN->0)
//   correlation_to_distance     /root/reference/notebooks/2_graph_construction.ipynb:100-122 (method "euclidean";
//                               "abs" / "standard" / "sqrt" are selectable too)
//
// The windows are never materialised: a CTA reads its (C x win) slice straight out of the filtered
// recording (coalesced 8-byte loads along time), keeps it in shared memory (row stride padded to an
// odd number of doubles so the 4x4 register tiles read conflict-free), centres it, and accumulates
// the upper-triangular 4x4 tiles of the Gram in FP64 FMAs.  FP64 is required: d = sqrt(2(1-r))
// cancels catastrophically for r -> 1 and north_star asks 1e-5 relative on d.
// The epilogue (divide by the two standard deviations in numpy's order, clip, NaN->0, distance, zero
// diagonal, cast) is fused; the correlation matrix itself is written only on request.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace corrdist {

constexpr int kThreads = 128;
constexpr int kTile = 4;

__global__ void __launch_bounds__(kThreads) corrdist_kernel(const double* __restrict__ x, int R, int C, long long T,
                                                            long long strideR, int win, int step, int W, int method,
                                                            float* __restrict__ D, double* __restrict__ corr,
                                                            long long strideO) {
    extern __shared__ __align__(16) double sm[];
    const int Cp = (C + kTile - 1) / kTile * kTile;  // channels padded to the tile
    const int ldw = win | 1;                         // odd row stride (in doubles)
    double* xs = sm;                                 // Cp x ldw
    double* cs = sm + (size_t)Cp * ldw;              // Cp x Cp   Gram / covariance
    double* sd = cs + (size_t)Cp * Cp;               // Cp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kThreads / 32;
    const int nt = Cp / kTile;
    const int ntiles = nt * (nt + 1) / 2;

    for (long long item = blockIdx.x; item < (long long)R * W; item += gridDim.x) {
        const long long rec = item / W;
        const int w = (int)(item % W);
        const double* src = x + rec * strideR + (long long)w * step;
        // ---- load (coalesced along time) and centre each channel
        for (int c = warp; c < Cp; c += nwarp) {
            double* row = xs + (size_t)c * ldw;
            if (c < C) {
                const double* g = src + (long long)c * T;
                double s = 0;
                for (int k = lane; k < win; k += 32) { double v = g[k]; row[k] = v; s += v; }
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                const double mean = s / win;
                for (int k = lane; k < win; k += 32) row[k] -= mean;
            } else {
                for (int k = lane; k < win; k += 32) row[k] = 0.0;
            }
        }
        __syncthreads();
        // ---- upper-triangular 4x4 tiles of X X^T
        for (int t = tid; t < ntiles; t += kThreads) {
            // tile index -> (ti <= tj)
            int ti = 0, rem = t;
            while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
            const int tj = ti + rem;
            const double* a = xs + (size_t)(ti * kTile) * ldw;
            const double* b = xs + (size_t)(tj * kTile) * ldw;
            double acc[kTile][kTile];
#pragma unroll
            for (int p = 0; p < kTile; ++p)
#pragma unroll
                for (int q = 0; q < kTile; ++q) acc[p][q] = 0.0;
            for (int k = 0; k < win; ++k) {
                double av[kTile], bv[kTile];
#pragma unroll
                for (int p = 0; p < kTile; ++p) { av[p] = a[(size_t)p * ldw + k]; bv[p] = b[(size_t)p * ldw + k]; }
#pragma unroll
                for (int p = 0; p < kTile; ++p)
#pragma unroll
                    for (int q = 0; q < kTile; ++q) acc[p][q] = fma(av[p], bv[q], acc[p][q]);
            }
            const double inv = 1.0 / (double)(win - 1);
#pragma unroll
            for (int p = 0; p < kTile; ++p)
#pragma unroll
                for (int q = 0; q < kTile; ++q) cs[(size_t)(ti * kTile + p) * Cp + tj * kTile + q] = acc[p][q] * inv;
        }
        __syncthreads();
        for (int c = tid; c < C; c += kThreads) sd[c] = sqrt(cs[(size_t)c * Cp + c]);
        __syncthreads();
        // ---- epilogue over i <= j
        const long long oo = rec * strideO + (long long)w * C * C;
        float* Dw = D ? D + oo : nullptr;
        double* Cw = corr ? corr + oo : nullptr;
        for (int e = tid; e < C * C; e += kThreads) {
            const int i = e / C, j = e % C;
            if (i > j) continue;
            double r = cs[(size_t)i * Cp + j];
            r = r / sd[i];
            r = r / sd[j];                       // numpy: c /= stddev[:, None]; c /= stddev[None, :]
            if (r != r) r = 0.0;                 // np.clip propagates NaN, nan_to_num(nan=0.0): zero-variance channel
            else r = fmin(fmax(r, -1.0), 1.0);   // np.clip inside corrcoef
            if (Cw) { Cw[(size_t)i * C + j] = r; Cw[(size_t)j * C + i] = r; }
            if (Dw) {
                double d;
                if (method == 0) d = sqrt(2.0 * (1.0 - r));
                else if (method == 1) d = 1.0 - fabs(r);
                else if (method == 2) d = 1.0 - r;
                else d = sqrt(1.0 - r * r);
                d = fmax(d, 0.0);
                if (i == j) d = 0.0;
                const float f = (float)d;
                Dw[(size_t)i * C + j] = f;
                Dw[(size_t)j * C + i] = f;
            }
        }
        __syncthreads();
    }
}

// corrdist_mma.cu: the Gram on the FP64 tensor pipe (the default for the shapes it takes)
int launch_corrdist_mma(const double* x, int R, int C, long long T, long long strideR, int win, int step, long long W,
                        int method, float* D, double* corr, long long strideO, cudaStream_t stream);

static size_t smem_bytes(int C, int win) {
    const int Cp = (C + kTile - 1) / kTile * kTile;
    return ((size_t)Cp * (win | 1) + (size_t)Cp * Cp + Cp) * sizeof(double);
}

}  // namespace corrdist
}  // namespace tda

extern "C" int tda_corrdist_windows(const double* x, int R, int C, long long T, long long strideR, int win, int step,
                                    int method, float* D, double* corr, long long strideO, void* stream) {
    using namespace tda::corrdist;
    if (!x || (!D && !corr) || R < 0 || C < 1 || win < 2 || step < 1 || T < 0 || method < 0 || method > 3)
        return TDA_E_ARG;
    if (T < win || R == 0) return 0;  // no window fits: nothing to write
    const long long W = (T - win) / step + 1;
    const size_t smem = smem_bytes(C, win);
    if (smem > 227 * 1024) return TDA_E_SIZE;
    if (strideR == 0) strideR = (long long)C * T;
    if (strideO == 0) strideO = W * (long long)C * C;
    {
        // the Gram on the FP64 tensor pipe (corrdist_mma.cu) for every shape it takes (the 47 x 250
        // EEG windows among them); the register-tile FMA kernel below serves the others
        const int rc = launch_corrdist_mma(x, R, C, T, strideR, win, step, W, method, D, corr, strideO,
                                           (cudaStream_t)stream);
        if (rc != TDA_E_SIZE) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(corrdist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    long long items = (long long)R * W;
    long long grid = (long long)sms * per_sm;
    if (grid > items) grid = items;
    tda::ProfScope prof("corrdist", (cudaStream_t)stream);
    corrdist_kernel<<<(int)grid, kThreads, smem, (cudaStream_t)stream>>>(x, R, C, T, strideR, win, step, (int)W, method,
                                                                        D, corr, strideO);
    tda::count_launch();
    return (int)cudaGetLastError();
}

// ---- single-matrix helpers for the drop-in surface ------------------------------------------
namespace tda {
namespace corrdist {
// correlation_to_distance on an existing correlation matrix (float64 in, float64 out)
__global__ void corr_to_dist_kernel(const double* __restrict__ c, int n, int method, double* __restrict__ d) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    double r = c[e];
    if (r == r) r = fmin(fmax(r, -1.0), 1.0);  // np.clip keeps NaN
    double v;
    if (method == 0) v = sqrt(2.0 * (1.0 - r));
    else if (method == 1) v = 1.0 - fabs(r);
    else if (method == 2) v = 1.0 - r;
    else v = sqrt(1.0 - r * r);
    if (v == v) v = fmax(v, 0.0);              // np.maximum keeps NaN
    if (e / n == e % n) v = 0.0;
    d[e] = v;
}
// (D + D^T)/2, zero diagonal, clamp at 0, cast to float32: what compute_eeg_persistence does to
// its float64 input before ripser (/root/reference/scripts/utils.py:137-139)
__global__ void symmetrize_kernel(const double* __restrict__ D, long long B, int n, float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * n * n) return;
    const long long b = e / (n * n);
    const int ij = (int)(e % (n * n)), i = ij / n, j = ij % n;
    const double* M = D + b * n * n;
    double v = (M[i * n + j] + M[j * n + i]) / 2;
    if (i == j) v = 0.0;
    v = fmax(v, 0.0);
    out[e] = (float)v;
}
// validate_distance_matrix (/root/reference/scripts/tda_eeg_classification_v2.py:110-140) for a batch:
// one warp per matrix.  flags: TDA_DM_* bits; stats[b] = { max |D - D^T|, min D, max |diag| } with
// numpy's NaN propagation (np.max / np.min return NaN as soon as one is present).
__global__ void __launch_bounds__(128) validate_kernel(const double* __restrict__ D, long long B, int n,
                                                       int* __restrict__ flags, double* __restrict__ stats) {
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    const double* M = D + b * (long long)n * n;
    int f = 0;
    double mx_asym = 0.0, mn = __longlong_as_double(0x7FF0000000000000ll), mx_diag = 0.0;
    bool nan_asym = false, nan_any = false, nan_diag = false;
    for (int e = lane; e < n * n; e += 32) {
        const int i = e / n, j = e - i * n;
        const double a = M[e], t = M[(size_t)j * n + i];
        // np.allclose(D, D.T, rtol=1e-5, atol=1e-8): finite pairs |a - t| <= atol + rtol |t|, non-finite
        // ones must be equal (inf == inf is close, NaN never is)
        const double diff = fabs(a - t);
        const bool fin = isfinite(a) && isfinite(t);
        if (!(fin ? diff <= 1e-8 + 1e-5 * fabs(t) : a == t)) f |= TDA_DM_ASYMMETRIC;
        if (diff != diff) nan_asym = true; else mx_asym = fmax(mx_asym, diff);
        if (a < -1e-10) f |= TDA_DM_NEGATIVE;
        if (a != a) { f |= TDA_DM_NAN; nan_any = true; } else mn = fmin(mn, a);
        if (isinf(a)) f |= TDA_DM_INF;
        if (i == j) {
            if (!(fabs(a) <= 1e-10)) f |= TDA_DM_DIAGONAL;   // np.allclose(diag, 0, atol=1e-10)
            if (a != a) nan_diag = true; else mx_diag = fmax(mx_diag, fabs(a));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        f |= __shfl_xor_sync(0xFFFFFFFFu, f, o);
        mx_asym = fmax(mx_asym, __shfl_xor_sync(0xFFFFFFFFu, mx_asym, o));
        mn = fmin(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
        mx_diag = fmax(mx_diag, __shfl_xor_sync(0xFFFFFFFFu, mx_diag, o));
    }
    const bool na = __any_sync(0xFFFFFFFFu, nan_asym), nn = __any_sync(0xFFFFFFFFu, nan_any),
               nd = __any_sync(0xFFFFFFFFu, nan_diag);
    if (lane == 0) {
        const double qnan = __longlong_as_double(0x7FF8000000000000ll);
        flags[b] = f;
        stats[3 * b] = na ? qnan : mx_asym;
        stats[3 * b + 1] = nn ? qnan : mn;
        stats[3 * b + 2] = nd ? qnan : mx_diag;
    }
}
}  // namespace corrdist
}  // namespace tda

extern "C" int tda_validate_distance_f64(const double* D, long long B, int n, int* flags, double* stats,
                                         void* stream) {
    if (!D || !flags || !stats || B < 0 || n < 1) return TDA_E_ARG;
    if (B == 0) return 0;
    tda::corrdist::validate_kernel<<<(unsigned)((B + 3) / 4), 128, 0, (cudaStream_t)stream>>>(D, B, n, flags, stats);
    tda::count_launch();
    return (int)cudaGetLastError();
}

extern "C" int tda_corr_to_dist_f64(const double* corr, int n, int method, double* dist, void* stream) {
    if (!corr || !dist || n < 1 || method < 0 || method > 3) return TDA_E_ARG;
    tda::corrdist::corr_to_dist_kernel<<<(n * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(corr, n, method, dist);
    tda::count_launch();
    return (int)cudaGetLastError();
}

extern "C" int tda_symmetrize_f64_to_f32(const double* D, long long B, int n, float* out, void* stream) {
    if (!D || !out || B < 0 || n < 1) return TDA_E_ARG;
    if (B == 0) return 0;
    const long long tot = B * n * n;
    tda::corrdist::symmetrize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(D, B, n, out);
    tda::count_launch();
    return (int)cudaGetLastError();
}
