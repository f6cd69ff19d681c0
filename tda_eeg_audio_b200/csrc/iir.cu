// iir.cu — batched zero-phase IIR filtering (forward + backward pass, odd extension) in FP64,
// reproducing scipy's recursions operation by operation.
//
// Replaces:
//   signal.sosfiltfilt(sos, x)   /root/reference/notebooks/1_preprocesamiento.ipynb:262-263 (EEG bands, "sos" form)
//   sig_proc.filtfilt(b, a, x)   /root/reference/scripts/utils.py:63 (LP50 of the envelope), :74 (audio bands, "ba" form)
//
// Why the recursion is replicated exactly instead of "the ideal filter": the ba-form delta band
// (0.5-4 Hz at 250 Hz) has poles at |z| = 0.996 and b ~ 3e-6; the sos and ba results differ by
// 2.8e-4 relative in float64, so only the same direct-form-II-transposed
// update order in FP64, without FMA contraction, lands on the reference's numbers.
//
// Mapping: one thread owns one (band, sequence) job — the recursion is serial in time, and at the
// benchmark shape there are 332,760 independent jobs.  A CTA of 128 jobs of ONE band (its
// coefficients pinned in registers: the recursion's SASS is 36 DMUL/DADD per sample and nothing
// else) moves time tiles of 16 samples through shared memory so that every HBM access is a
// coalesced 128-byte row segment (lanes along time on the way in/out, lanes along jobs inside the
// recursion; odd row stride keeps both conflict-free).  The staging is software-pipelined: while the
// CTA runs the recursion on tile q, the sixteen values each thread contributes to tile q+1 are in
// flight from HBM into its registers.  The forward pass materialises the padded intermediate once
// (workspace); the backward pass walks the same tiles in reverse and writes only the un-padded
// samples.  Measured against the two earlier generations of this kernel (coefficients from the
// constant bank; unpipelined staging), outputs bit-equal: 27.9 / 16.8 / 9.9 ms for the five EEG bands
// of 256 recordings (profiles/r02_staged_ab.jsonl).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace iir {

constexpr int kJobs = 128;   // threads per CTA = jobs per CTA
constexpr int kTT = 16;      // samples per time tile (36 KB of staging per CTA)
constexpr int kRPW = 32 / kTT;  // tile rows one warp moves per step
constexpr int kLd = kTT + 1; // odd row stride
constexpr int kMaxBands = 8;
constexpr int kMaxSec = 4;   // sos sections
constexpr int kMaxTaps = 9;  // ba taps

struct Coef {
    // form 0: sos[s][6] = b0 b1 b2 a0 a1 a2 ; zi[s][2]
    // form 1: b[0..nt), a[0..nt) already divided by a[0] ; zi[0..nt-1)
    double c[kMaxBands][kMaxSec * 6];
    double zi[kMaxBands][kMaxSec * 2];
    int n;  // sections (form 0) or taps (form 1)
};

// The two recursions, the band's coefficients in registers (cr = Coef::c[band]):
//   sos: scipy _sosfilt: x_new = b0*x + z0 ; z0 = b1*x - a1*x_new + z1 ; z1 = b2*x - a2*x_new
//   ba : scipy lfilter (direct form II transposed):
//        y = Z[0] + b[0]*x ; Z[n] = Z[n+1] + x*b[n+1] - y*a[n+1] ; Z[last] = x*b[last] - y*a[last]
template <int FORM>
__device__ __forceinline__ double step_regs(double (&z)[8], const double (&cr)[kMaxSec * 6], int n, double x) {
    if (FORM == 0) {
#pragma unroll
        for (int s = 0; s < kMaxSec; ++s) {
            if (s < n) {
                const double xn = __dadd_rn(__dmul_rn(cr[s * 6 + 0], x), z[2 * s]);
                z[2 * s] = __dadd_rn(__dsub_rn(__dmul_rn(cr[s * 6 + 1], x), __dmul_rn(cr[s * 6 + 4], xn)), z[2 * s + 1]);
                z[2 * s + 1] = __dsub_rn(__dmul_rn(cr[s * 6 + 2], x), __dmul_rn(cr[s * 6 + 5], xn));
                x = xn;
            }
        }
        return x;
    } else {
        if (n == 1) return __dmul_rn(x, cr[0]);
        const double y = __dadd_rn(z[0], __dmul_rn(cr[0], x));
#pragma unroll
        for (int k = 0; k < kMaxTaps - 2; ++k) {
            if (k < n - 2)
                z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(x, cr[k + 1])), __dmul_rn(y, cr[kMaxTaps + k + 1]));
        }
#pragma unroll
        for (int k = 0; k < kMaxTaps - 1; ++k) {
            if (k == n - 2) z[k] = __dsub_rn(__dmul_rn(x, cr[k + 1]), __dmul_rn(y, cr[kMaxTaps + k + 1]));
        }
        return y;
    }
}

// value of the odd-extended signal at padded position k (scipy _arraytools.odd_ext)
__device__ __forceinline__ double ext_value(const double* __restrict__ x, long long T, int edge, long long k) {
    if (k < edge) return __dsub_rn(__dmul_rn(2.0, x[0]), x[edge - k]);
    if (k < edge + T) return x[k - edge];
    return __dsub_rn(__dmul_rn(2.0, x[T - 1]), x[T - 2 - (k - edge - T)]);
}

// jobs are (band, seq): job = band * n_seq + seq; a CTA works on groups of 128 sequences of one band
//   forward : in = x (n_seq rows, stride x_stride), out = mid (n_jobs rows of Text)
//   backward: in = mid, out = y (n_jobs rows of T, row stride T)
template <int FORM, bool BACKWARD>
__global__ void __launch_bounds__(kJobs, 4) iir_pass_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                            long long n_seq, int n_bands, long long T,
                                                            long long x_stride, int edge,
                                                            const __grid_constant__ Coef cf) {
    extern __shared__ __align__(16) double iir_smem[];
    double* tin = iir_smem;
    double* tout = iir_smem + kJobs * kLd;
    long long* inbase = reinterpret_cast<long long*>(iir_smem + 2 * kJobs * kLd);   // element offset of a job's input row
    long long* outbase = inbase + kJobs;                                             // ... of its output row
    constexpr int kRows = kJobs / ((kJobs / 32) * kRPW);   // rows of a tile one thread moves (16)
    constexpr int kRStep = (kJobs / 32) * kRPW;            // distance between them (8)
    const long long Text = T + 2LL * edge;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = warp * kRPW + lane / kTT;   // first tile row of this thread
    const int c = lane % kTT;                  // its column (time within the tile)
    const long long n_tiles = (Text + kTT - 1) / kTT;
    // job groups: ceil(n_seq / 128) groups per band, each within one band
    const long long gpb = (n_seq + kJobs - 1) / kJobs;
    const long long n_groups = gpb * n_bands;
    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int band = (int)(grp / gpb);                                   // uniform in the CTA
        const long long seq0 = (grp - (long long)band * gpb) * kJobs;
        const long long job = (long long)band * n_seq + seq0 + tid;
        const bool active = seq0 + tid < n_seq;
        double cr[kMaxSec * 6];
#pragma unroll
        for (int k = 0; k < kMaxSec * 6; ++k) {
            const bool used = FORM == 0 ? (k % 6 != 3) : (k < 2 * kMaxTaps);
            cr[k] = used ? cf.c[band][k] : 0.0;
            if (used) asm volatile("" : "+d"(cr[k]));   // a register, not a constant-bank operand
        }
        __syncthreads();   // the previous group's last tile has left tin / tout / the offset tables
        inbase[tid] = active ? (BACKWARD ? job * Text : (seq0 + tid) * x_stride) : 0;
        outbase[tid] = active ? (BACKWARD ? job * T : job * Text) : 0;
        double z[8];
        if (active) {
            double scale;
            if (!BACKWARD) scale = ext_value(in + (seq0 + tid) * x_stride, T, edge, 0);
            else scale = in[job * Text + (Text - 1)];
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = __dmul_rn(cf.zi[band][k], scale);
        }
        __syncthreads();
        const long long left = n_seq - seq0;
        const int rows_here = (int)(left < kJobs ? left : kJobs);
        double nx[kRows];
        // values of tile `tile` this thread stages: rows r0, r0 + 8, ..., column c
        auto fetch = [&](long long tile) {
            const long long k = tile * kTT + c;
            const bool interior = BACKWARD || (tile * kTT >= edge && tile * kTT + kTT <= edge + T);
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
                const int r = r0 + kRStep * i;
                double v = 0.0;
                if (r < rows_here && k < Text) {
                    const double* row = in + inbase[r];
                    if (BACKWARD) v = row[k];
                    else if (interior) v = row[k - edge];
                    else v = ext_value(row, T, edge, k);
                }
                nx[i] = v;
            }
        };
        fetch(BACKWARD ? n_tiles - 1 : 0);
#pragma unroll
        for (int i = 0; i < kRows; ++i) tin[(r0 + kRStep * i) * kLd + c] = nx[i];
        __syncthreads();
        for (long long q = 0; q < n_tiles; ++q) {
            const long long tile = BACKWARD ? (n_tiles - 1 - q) : q;
            const long long k0 = tile * kTT;
            if (q + 1 < n_tiles) fetch(BACKWARD ? tile - 1 : tile + 1);   // in flight during the recursion
            // ---- serial recursion, one job per thread
            if (active) {
                const int nvalid = (int)((Text - k0 < kTT) ? (Text - k0) : kTT);
                if (!BACKWARD) {
                    for (int cc = 0; cc < nvalid; ++cc)
                        tout[tid * kLd + cc] = step_regs<FORM>(z, cr, cf.n, tin[tid * kLd + cc]);
                } else {
                    for (int cc = nvalid - 1; cc >= 0; --cc)
                        tout[tid * kLd + cc] = step_regs<FORM>(z, cr, cf.n, tin[tid * kLd + cc]);
                }
            }
            __syncthreads();   // tout complete, tin consumed
            // ---- cooperative coalesced store of tile q, then the staged tile q + 1 takes tin
            {
                const long long k = k0 + c;
                const bool keep = BACKWARD ? (k >= edge && k < edge + T) : (k < Text);
                const long long ko = BACKWARD ? k - edge : k;
#pragma unroll
                for (int i = 0; i < kRows; ++i) {
                    const int r = r0 + kRStep * i;
                    if (keep && r < rows_here) out[outbase[r] + ko] = tout[r * kLd + c];
                }
            }
            if (q + 1 < n_tiles) {
#pragma unroll
                for (int i = 0; i < kRows; ++i) tin[(r0 + kRStep * i) * kLd + c] = nx[i];
            }
            __syncthreads();   // tin holds tile q + 1, tout is free
        }
    }
}

}  // namespace iir
}  // namespace tda

extern "C" size_t tda_filtfilt_workspace_bytes(long long n_seq, int n_bands, long long T, int padlen) {
    if (n_seq < 0 || n_bands < 1 || T < 1 || padlen < 0) return 0;
    return (size_t)n_seq * n_bands * (size_t)(T + 2LL * padlen) * sizeof(double);
}

extern "C" int tda_filtfilt_f64(const double* x, long long n_seq, long long T, long long x_stride, int form,
                                int n_bands, int n, const double* coef, const double* zi, int padlen, double* y,
                                void* ws, size_t ws_bytes, void* stream) {
    using namespace tda::iir;
    if (!x || !coef || !zi || !y || !ws || n_seq < 0 || T < 1 || n_bands < 1 || n_bands > kMaxBands || padlen < 0)
        return TDA_E_ARG;
    if (form == 0 && (n < 1 || n > kMaxSec)) return TDA_E_SIZE;
    if (form == 1 && (n < 1 || n > kMaxTaps)) return TDA_E_SIZE;
    if (form != 0 && form != 1) return TDA_E_ARG;
    if (padlen >= T) return TDA_E_ARG;  // scipy: "The length of the input vector x must be greater than padlen"
    if (n_seq == 0) return 0;
    if (ws_bytes < tda_filtfilt_workspace_bytes(n_seq, n_bands, T, padlen)) return TDA_E_WORKSPACE;
    if (x_stride == 0) x_stride = T;
    Coef cf;
    for (int b = 0; b < kMaxBands; ++b) {
        for (int k = 0; k < kMaxSec * 6; ++k) cf.c[b][k] = 0.0;
        for (int k = 0; k < kMaxSec * 2; ++k) cf.zi[b][k] = 0.0;
    }
    cf.n = n;
    for (int b = 0; b < n_bands; ++b) {
        if (form == 0) {
            for (int k = 0; k < n * 6; ++k) cf.c[b][k] = coef[(size_t)b * n * 6 + k];
            for (int k = 0; k < n * 2; ++k) cf.zi[b][k] = zi[(size_t)b * n * 2 + k];
        } else {
            const double a0 = coef[(size_t)b * 2 * n + n];
            for (int k = 0; k < n; ++k) {
                cf.c[b][k] = coef[(size_t)b * 2 * n + k] / a0;                 // b / a0
                cf.c[b][kMaxTaps + k] = coef[(size_t)b * 2 * n + n + k] / a0;  // a / a0
            }
            for (int k = 0; k < n - 1; ++k) cf.zi[b][k] = zi[(size_t)b * (n - 1) + k];
        }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long groups = ((n_seq + kJobs - 1) / kJobs) * n_bands;
    const long long maxb = (long long)sms * 4;  // 128 registers, 36 KB of staging per CTA -> 4 CTAs per SM
    const int grid = (int)(groups < maxb ? groups : maxb);
    cudaStream_t st = (cudaStream_t)stream;
    double* mid = (double*)ws;
    const int smem = 2 * kJobs * kLd * (int)sizeof(double) + 2 * kJobs * (int)sizeof(long long);
    if (form == 0) {
        cudaFuncSetAttribute(iir_pass_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(iir_pass_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    } else {
        cudaFuncSetAttribute(iir_pass_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(iir_pass_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    {
        tda::ProfScope prof(form == 0 ? "iir_sos_forward" : "iir_ba_forward", st);
        if (form == 0) iir_pass_kernel<0, false><<<grid, kJobs, smem, st>>>(x, mid, n_seq, n_bands, T, x_stride, padlen, cf);
        else iir_pass_kernel<1, false><<<grid, kJobs, smem, st>>>(x, mid, n_seq, n_bands, T, x_stride, padlen, cf);
        tda::count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    {
        tda::ProfScope prof(form == 0 ? "iir_sos_backward" : "iir_ba_backward", st);
        if (form == 0) iir_pass_kernel<0, true><<<grid, kJobs, smem, st>>>(mid, y, n_seq, n_bands, T, 0, padlen, cf);
        else iir_pass_kernel<1, true><<<grid, kJobs, smem, st>>>(mid, y, n_seq, n_bands, T, 0, padlen, cf);
        tda::count_launch();
    }
    return (int)cudaGetLastError();
}
