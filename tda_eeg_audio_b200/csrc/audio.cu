// audio.cu — the audio front end of the hot path, FP64:
//   tda_resample_poly_f64     scipy.signal.resample_poly(x, up, down) as the reference calls it in
//                             resample_audio (/root/reference/scripts/utils.py:77-79; 44.1 kHz -> 250 Hz
//                             is up 5 / down 882 with a 17,641-tap Kaiser FIR)
//   tda_hilbert_envelope_f64  abs(scipy.signal.hilbert(x)), the first half of compute_envelope
//                             (/root/reference/scripts/utils.py:56-63; the LP50 filtfilt that follows is
//                             tda_filtfilt_f64, form 1)
//
// resample: only the KEPT outputs are computed (scipy's upfirdn does the same).  One CTA owns a
// tile of 4*UP consecutive outputs of one sequence; the input span the tile needs (~6k samples)
// is staged once in shared memory with coalesced loads,
// every thread walks a strided slice of the polyphase taps (coalesced, L2-resident: the filter is
// 140 KB, the taps of the next step requested one step ahead) and keeps all 4*UP partial sums in
// registers; warp shuffles + one shared-memory pass finish the dot products.  One LDS.64 per FMA;
// four CTAs of four warps per SM.  ~45 ms for the full 1,416-recording data set against 81 ms per
// recording on the CPU.
// hilbert: cuFFT Z2Z (plain library FFT; N = 15,000 = 2^3 3 5^4) between two elementwise kernels.
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace audio {

constexpr int kMaxUp = 8;
constexpr int kPerPhase = 4;   // outputs per phase and tile (measured 8 / 4 / 2: 4.45 / 3.77 / 4.35 ms for 96 recordings: staged span against tap reuse)
constexpr int kThreads = 128;

struct ResParams {
    const double* x;
    long long n_seq, n_in, x_stride;
    int up, down;
    const double* hpoly;   // [up][qmax]: hpoly[p][q] = h_padded[p + q*up] (0 beyond the filter)
    int qmax;
    long long n_pre_remove, n_out;
    double* y;
    long long y_stride;
    int rel[kMaxUp * kPerPhase];   // input offset of the slot's output relative to the tile base
    int oidx[kMaxUp * kPerPhase];  // which output of the tile the slot accumulates
    int span;                      // staged input samples per tile
};

template <int UP> __global__ void __launch_bounds__(kThreads) resample_kernel(ResParams p) {
    constexpr int TO = UP * kPerPhase;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* xs = (double*)smem_raw;
    double* part = xs + p.span;  // [kThreads / 32][TO]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long seq = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * TO;          // first output (upfirdn index) of the tile
    const long long b0 = t0 / UP * p.down;                    // its newest input sample
    const long long lo = b0 - (p.qmax - 1);
    const double* xr = p.x + seq * p.x_stride;
    for (int i = tid; i < p.span; i += kThreads) {
        const long long g = lo + i;
        xs[i] = (g >= 0 && g < p.n_in) ? xr[g] : 0.0;
    }
    __syncthreads();
    double acc[TO];
#pragma unroll
    for (int a = 0; a < TO; ++a) acc[a] = 0.0;
    // the taps of the NEXT step are requested before this step's 8 UP multiply-adds: with two CTAs of four warps
    // per SM (80 KB of staged input each) nothing else hides the L2 round trip of the tap loads
    double hn[UP];
#pragma unroll
    for (int ph = 0; ph < UP; ++ph) hn[ph] = tid < p.qmax ? __ldg(p.hpoly + (size_t)ph * p.qmax + tid) : 0.0;
    for (int q = tid; q < p.qmax; q += kThreads) {
        const double* xq = xs + (p.qmax - 1 - q);
        double hv[UP];
#pragma unroll
        for (int ph = 0; ph < UP; ++ph) hv[ph] = hn[ph];
        if (q + kThreads < p.qmax) {
#pragma unroll
            for (int ph = 0; ph < UP; ++ph) hn[ph] = __ldg(p.hpoly + (size_t)ph * p.qmax + q + kThreads);
        }
#pragma unroll
        for (int ph = 0; ph < UP; ++ph) {
#pragma unroll
            for (int k = 0; k < kPerPhase; ++k) acc[ph * kPerPhase + k] = fma(hv[ph], xq[p.rel[ph * kPerPhase + k]], acc[ph * kPerPhase + k]);
        }
    }
#pragma unroll
    for (int a = 0; a < TO; ++a) {
        double v = acc[a];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) part[warp * TO + a] = v;
    }
    __syncthreads();
    if (tid < TO) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += part[w * TO + tid];
        const long long k = t0 + p.oidx[tid] - p.n_pre_remove;
        if (k >= 0 && k < p.n_out) p.y[seq * p.y_stride + k] = v;
    }
}

// ---------------------------------------------------------------------------------- hilbert
__global__ void to_complex_kernel(const double* x, long long n_seq, long long T, long long x_stride, double2* z) {
    const long long total = n_seq * T;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long s = q / T, t = q - s * T;
        z[q] = make_double2(x[s * x_stride + t], 0.0);
    }
}
// scipy.signal.hilbert: h[0] = 1, h[1 .. (N-1)/2] = 2 (N odd) or h[1 .. N/2-1] = 2, h[N/2] = 1 (N even), else 0
__global__ void analytic_mask_kernel(double2* z, long long n_seq, long long T) {
    const long long total = n_seq * T;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long t = q % T;
        double h;
        if (T % 2 == 0) h = (t == 0 || t == T / 2) ? 1.0 : (t < T / 2 ? 2.0 : 0.0);
        else h = (t == 0) ? 1.0 : (t <= (T - 1) / 2 ? 2.0 : 0.0);
        double2 v = z[q];
        v.x *= h; v.y *= h;
        z[q] = v;
    }
}
__global__ void abs_scale_kernel(const double2* z, long long n_seq, long long T, double inv_n, double* env, long long env_stride) {
    const long long total = n_seq * T;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long s = q / T, t = q - s * T;
        const double2 v = z[q];
        env[s * env_stride + t] = hypot(v.x * inv_n, v.y * inv_n);
    }
}

// cuFFT plans are expensive to build: one per (device, T, batch), kept for the life of the library
static std::mutex g_mu;
static std::map<std::tuple<int, long long, long long>, cufftHandle> g_plans;
static int get_plan(long long T, long long batch, cufftHandle* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, T, batch);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return 0; }
    cufftHandle h;
    int n[1] = {(int)T};
    cufftResult r = cufftPlanMany(&h, 1, n, nullptr, 1, (int)T, nullptr, 1, (int)T, CUFFT_Z2Z, (int)batch);
    if (r != CUFFT_SUCCESS) return 10000 + (int)r;
    g_plans[key] = h;
    *out = h;
    return 0;
}

}  // namespace audio
}  // namespace tda

using namespace tda;
using namespace tda::audio;

extern "C" int tda_resample_poly_f64(const double* x, long long n_seq, long long n_in, long long x_stride, int up,
                                     int down, const double* hpoly, int qmax, long long n_pre_remove,
                                     long long n_out, double* y, long long y_stride, void* stream) {
    if (!x || !hpoly || !y || n_seq < 0 || n_in < 1 || n_out < 0 || qmax < 1 || down < 1) return TDA_E_ARG;
    if (up < 1 || up > kMaxUp) return TDA_E_SIZE;
    if (n_seq == 0 || n_out == 0) return 0;
    ResParams p;
    p.x = x; p.n_seq = n_seq; p.n_in = n_in; p.x_stride = x_stride ? x_stride : n_in;
    p.up = up; p.down = down; p.hpoly = hpoly; p.qmax = qmax;
    p.n_pre_remove = n_pre_remove; p.n_out = n_out; p.y = y; p.y_stride = y_stride ? y_stride : n_out;
    const int TO = up * kPerPhase;
    // slot (phase ph, k-th output of that phase in the tile): output o has phase (o*down) % up and
    // newest input floor(o*down/up) past the tile base (tiles start at multiples of `up`)
    int cnt[kMaxUp] = {0};
    int maxrel = 0;
    for (int o = 0; o < TO; ++o) {
        const int ph = (int)(((long long)o * down) % up);
        const int rel = (int)(((long long)o * down) / up);
        const int slot = ph * kPerPhase + cnt[ph]++;
        p.rel[slot] = rel;
        p.oidx[slot] = o;
        if (rel > maxrel) maxrel = rel;
    }
    for (int ph = 0; ph < up; ++ph)
        if (cnt[ph] != kPerPhase) return TDA_E_SIZE;  // up and down must be coprime (scipy reduces them)
    p.span = qmax + maxrel;
    const size_t smem = ((size_t)p.span + (size_t)(kThreads / 32) * TO) * sizeof(double);
    if (smem > 227 * 1024) return TDA_E_SIZE;
    const long long t_last = n_pre_remove + n_out;  // outputs [n_pre_remove, t_last) are kept
    const long long tiles = (t_last + TO - 1) / TO;
    if (tiles > 0x7FFFFFFF || n_seq > 65535) return TDA_E_SIZE;
    dim3 grid((unsigned)tiles, (unsigned)n_seq);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof("resample_poly", st);
    cudaError_t e = cudaSuccess;
#define TDA_LAUNCH_RES(U)                                                                                        \
    case U:                                                                                                      \
        e = cudaFuncSetAttribute(resample_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) resample_kernel<U><<<grid, kThreads, smem, st>>>(p);                              \
        break;
    switch (up) {
        TDA_LAUNCH_RES(1) TDA_LAUNCH_RES(2) TDA_LAUNCH_RES(3) TDA_LAUNCH_RES(4)
        TDA_LAUNCH_RES(5) TDA_LAUNCH_RES(6) TDA_LAUNCH_RES(7) TDA_LAUNCH_RES(8)
    }
#undef TDA_LAUNCH_RES
    if (e != cudaSuccess) return (int)e;
    count_launch();
    return (int)cudaGetLastError();
}

extern "C" size_t tda_hilbert_envelope_workspace_bytes(long long n_seq, long long T) {
    if (n_seq < 0 || T < 1) return 0;
    return (size_t)n_seq * (size_t)T * sizeof(double2) + 256;
}

extern "C" int tda_hilbert_envelope_f64(const double* x, long long n_seq, long long T, long long x_stride, double* env,
                                        long long env_stride, void* ws, size_t ws_bytes, void* stream) {
    if (!x || !env || !ws || n_seq < 0 || T < 1) return TDA_E_ARG;
    if (T > 0x7FFFFFFF || n_seq > 0x7FFFFFFF) return TDA_E_SIZE;
    if (ws_bytes < tda_hilbert_envelope_workspace_bytes(n_seq, T)) return TDA_E_WORKSPACE;
    if (n_seq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    double2* z = (double2*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    cufftHandle plan = 0;
    int rc = get_plan(T, n_seq, &plan);
    if (rc) return rc;
    if (cufftSetStream(plan, st) != CUFFT_SUCCESS) return 10000;
    const long long total = n_seq * T;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    ProfScope prof("hilbert_envelope", st);
    to_complex_kernel<<<blocks, 256, 0, st>>>(x, n_seq, T, x_stride ? x_stride : T, z);
    if (cufftExecZ2Z(plan, (cufftDoubleComplex*)z, (cufftDoubleComplex*)z, CUFFT_FORWARD) != CUFFT_SUCCESS) return 10001;
    analytic_mask_kernel<<<blocks, 256, 0, st>>>(z, n_seq, T);
    if (cufftExecZ2Z(plan, (cufftDoubleComplex*)z, (cufftDoubleComplex*)z, CUFFT_INVERSE) != CUFFT_SUCCESS) return 10002;
    abs_scale_kernel<<<blocks, 256, 0, st>>>(z, n_seq, T, 1.0 / (double)T, env, env_stride ? env_stride : T);
    count_launch(5);
    return (int)cudaGetLastError();
}
