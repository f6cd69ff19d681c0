// features.cu — persistence statistics / entropy features and the per-recording window
// aggregation.
//
// Replaces (per diagram)  extract_features            /root/reference/scripts/utils.py:144-177
//                      ≡  extract_persistence_features /root/reference/scripts/tda_eeg_classification_v2.py:179-250
// and (per recording)  the mean/std-over-windows loop  /root/reference/scripts/tda_eeg_classification_v2.py:429-436
//
// HBM-bound and tiny: one warp per diagram, rows strided over lanes, float64 accumulators
// (the reference computes on float64 arrays that hold float32 values), two passes over rows that
// are L1-resident after the first.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace features {

constexpr uint32_t kFull = 0xFFFFFFFFu;

// reductions over the LPD lanes that share a diagram (LPD = 4: eight diagrams per warp)
template <int LPD> __device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = LPD / 2; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int LPD> __device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = LPD / 2; o; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
template <int LPD> __device__ __forceinline__ int wsumi(int v) {
#pragma unroll
    for (int o = LPD / 2; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// The kernel is issue-bound (float64 logarithm, division, square roots and the nine shuffle reductions per
// diagram), and a diagram of a 47-channel window has 46 + ~37 rows: with a whole warp per diagram most lanes idled
// in the second trip and every diagram paid a full warp's reductions and epilogue.  Measured per 849,600 diagrams:
// 32 lanes per diagram 0.95 ms, 16: 0.57, 8: 0.40, 4: 0.32, 2: 0.29 -- four lanes (eight diagrams per warp) it is.
template <int LPD>
__global__ void __launch_bounds__(256) pers_features_kernel(const float* __restrict__ bd, int cap,
                                                            const int* __restrict__ counts, int count_stride,
                                                            int B, double* __restrict__ feats, int feat_stride) {
    constexpr int DPW = 32 / LPD;   // diagrams per warp
    const int lane = threadIdx.x & 31, l = lane % LPD, sub = lane / LPD;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int b0 = gw * DPW; b0 < B; b0 += nw * DPW) {
        const int b = b0 + sub;
        const bool have = b < B;
        int n = have ? counts[(size_t)b * count_stride] : 0;
        if (n > cap) n = cap;
        const float2* rows = reinterpret_cast<const float2*>(bd) + (size_t)(have ? b : 0) * cap;
        // pass 1: counts and first moments over the finite rows
        int nf = 0;
        double sb = 0, sd = 0, sp = 0, mx = -INFINITY;
        for (int k = l; k < n; k += LPD) {
            float2 r = rows[k];
            if (isfinite(r.x) && isfinite(r.y)) {
                double bb = r.x, dd = r.y, pp = dd - bb;
                ++nf; sb += bb; sd += dd; sp += pp; mx = fmax(mx, pp);
            }
        }
        nf = wsumi<LPD>(nf); sb = wsum<LPD>(sb); sd = wsum<LPD>(sd); sp = wsum<LPD>(sp); mx = wmax<LPD>(mx);
        double* o = feats + (size_t)(have ? b : 0) * feat_stride;
        const double inv = nf > 0 ? (double)nf : 1.0;
        const double mb = sb / inv, md = sd / inv, mp = sp / inv;
        // pass 2: second central moments (np.std, ddof=0) and the entropy sum
        double vb = 0, vd = 0, vp = 0, ent = 0;
        const bool do_ent = nf > 1 && sp > 0;
        if (nf > 0) {
            for (int k = l; k < n; k += LPD) {
                float2 r = rows[k];
                if (isfinite(r.x) && isfinite(r.y)) {
                    double bb = r.x, dd = r.y, pp = dd - bb;
                    vb += (bb - mb) * (bb - mb); vd += (dd - md) * (dd - md); vp += (pp - mp) * (pp - mp);
                    if (do_ent) {
                        double pn = pp / sp;
                        if (pn > 0) ent += pn * log(pn + 1e-10);
                    }
                }
            }
        }
        vb = wsum<LPD>(vb); vd = wsum<LPD>(vd); vp = wsum<LPD>(vp); ent = wsum<LPD>(ent);
        if (!have) continue;   // (after the last shuffle: both halves of the warp take part in every one)
        if (nf == 0) {
            for (int q = l; q < 11; q += LPD) o[q] = (q == 1) ? (double)n : 0.0;
        } else if (l == 0) {
            const bool many = nf > 1;
            o[0] = nf;
            o[1] = n - nf;
            o[2] = mb;
            o[3] = many ? sqrt(vb / nf) : 0.0;
            o[4] = md;
            o[5] = many ? sqrt(vd / nf) : 0.0;
            o[6] = mp;
            o[7] = many ? sqrt(vp / nf) : 0.0;
            o[8] = mx;
            o[9] = sp;
            o[10] = do_ent ? -ent / log((double)nf + 1e-10) : 0.0;
        }
    }
}

// feats (R, Bd, Wn, 2, 11) -> table (R, Bd*44): column = band*44 + feat*4 + {h0 mean, h0 std, h1 mean, h1 std}
// (the order of /root/reference/features/feature_names.txt)
__global__ void aggregate_windows_kernel(const double* __restrict__ feats, int R, int Bd, int Wn,
                                         double* __restrict__ table) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)R * Bd * 22;
    if (t >= total) return;
    const int f = (int)(t % 11);
    const int dim = (int)((t / 11) % 2);
    const long long rb = t / 22;  // rec * Bd + band
    const double* src = feats + (rb * Wn) * 22 + dim * 11 + f;
    double s = 0;
    for (int w = 0; w < Wn; ++w) s += src[(size_t)w * 22];
    const double mean = Wn > 0 ? s / Wn : 0.0;
    double v = 0;
    for (int w = 0; w < Wn; ++w) {
        double x = src[(size_t)w * 22] - mean;
        v += x * x;
    }
    const int band = (int)(rb % Bd);
    const long long rec = rb / Bd;
    double* o = table + rec * (Bd * 44) + band * 44 + f * 4 + dim * 2;
    o[0] = mean;
    o[1] = Wn > 0 ? sqrt(v / Wn) : 0.0;
}

}  // namespace features
}  // namespace tda

extern "C" int tda_pers_features(const float* bd, int cap, const int* counts, int count_stride, int B,
                                 double* feats, int feat_stride, void* stream) {
    if (!bd || !counts || !feats || B < 0 || cap < 0 || count_stride < 1 || feat_stride < 11) return TDA_E_ARG;
    if (B == 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // four lanes per diagram whatever its size: the summation order, and with it the last bits of the features,
    // must not depend on the padding `cap` the caller happened to choose
    long long need = ((long long)B + 63) / 64;
    int grid = (int)(need < (long long)sms * 8 ? need : (long long)sms * 8);
    tda::ProfScope prof("pers_features", (cudaStream_t)stream);
    tda::features::pers_features_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(bd, cap, counts, count_stride, B,
                                                                                    feats, feat_stride);
    tda::count_launch();
    return (int)cudaGetLastError();
}

extern "C" int tda_aggregate_windows(const double* feats, int R, int Bd, int Wn, double* table, void* stream) {
    if (!feats || !table || R < 0 || Bd < 0 || Wn < 0) return TDA_E_ARG;
    long long total = (long long)R * Bd * 22;
    if (total == 0) return 0;
    int grid = (int)((total + 255) / 256);
    tda::ProfScope prof("aggregate_windows", (cudaStream_t)stream);
    tda::features::aggregate_windows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feats, R, Bd, Wn, table);
    tda::count_launch();
    return (int)cudaGetLastError();
}
