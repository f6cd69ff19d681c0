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
// benchmark shape there are 332,760 independent jobs.  A CTA holds 32 sequences x ALL bands: warp =
// band (its coefficients pinned in registers: the recursion's SASS is 36 DMUL/DADD per sample and
// nothing else), lane = sequence, so a recording is read from HBM once for all its bands.  What bounds
// the kernel is the memory path, not the FP64 pipe (a copy-only build of the previous generation ran
// in 90 % of the time of the real one; profiles/README.md), so the layout serves HBM:
//   * the padded intermediate between the two passes is stored TIME-MAJOR per CTA group,
//     mid[group][k][band*32 + seq]: the forward pass writes each sample of all its jobs as one
//     contiguous 1,280-byte row straight from registers, the backward pass streams whole 20 KB tiles
//     back with 16-byte cp.async copies — no 128-byte row segments at 120 KB strides;
//   * input tiles (forward: 32 channel rows x 16 samples; backward: 16 time rows of the intermediate)
//     are staged two tiles ahead with cp.async into a ring of three shared-memory tiles PER WARP (no
//     CTA-wide barrier anywhere in the steady state), filtered in place, and only the backward pass
//     transposes through shared memory to write the un-padded samples as coalesced row segments of y.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace iir {

constexpr int kSeq = 32;     // sequences per CTA (lanes); warps = bands
constexpr int kTT = 16;      // samples per time tile
constexpr int kLdIn = kTT + 1;  // row stride of a forward input tile (odd: lanes = rows read conflict-free)
constexpr int kRing = 3;     // tiles in flight per CTA
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
// NC > 0: the number of sections / taps is a compile-time constant, so a whole tile of the recursion is
// straight-line code that the scheduler interleaves across sections and samples (with the run-time
// count every section is a basic block of its own: measured 43 % of the FP64 pipe at full occupancy,
// profiles/r02b_iir_forward_ncu.json)
template <int FORM, int NC>
__device__ __forceinline__ double step_regs(double (&z)[8], const double (&cr)[kMaxSec * 6], int n_rt, double x) {
    const int n = NC > 0 ? NC : n_rt;
#ifdef IIR_COPY_ONLY
    return x;
#endif
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
__device__ __forceinline__ void cp_async_8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_16(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int FORM> __device__ __forceinline__ void load_coefficients(const Coef& cf, int band, double (&cr)[kMaxSec * 6]) {
#pragma unroll
    for (int k = 0; k < kMaxSec * 6; ++k) {
        const bool used = FORM == 0 ? (k % 6 != 3) : (k < 2 * kMaxTaps);
        cr[k] = used ? cf.c[band][k] : 0.0;
        if (used) asm volatile("" : "+d"(cr[k]));   // a register, not a constant-bank operand
    }
}

// Forward pass.  blockDim.x = 32 * n_bands; group g = sequences [32 g, 32 g + 32); warp = band.
//   in  = x (n_seq rows of T samples, stride x_stride)
//   out = mid, time-major per group: mid[(g * Text + k) * nth + band * 32 + lane], nth = blockDim.x
// Every warp is autonomous (its own ring of tiles, __syncwarp only): with CTA-wide barriers the five
// warps of a CTA, spread unevenly over four schedulers, waited for one another 2.5 cycles per issued
// instruction (profiles/r02c_iir_forward_ncu.json).  The bands of a group read the same x tile; the
// copies after the first hit L1 / L2.
template <int FORM, int NC>
__global__ void __launch_bounds__(kSeq * kMaxBands, 2) iir_forward_kernel(const double* __restrict__ x, double* __restrict__ mid,
                                                                          long long n_seq, long long T, long long x_stride,
                                                                          int edge, int* __restrict__ next_group,
                                                                          const __grid_constant__ Coef cf) {
    extern __shared__ __align__(16) double iir_smem[];   // per warp: ring of kRing tiles [32 rows][kLdIn]
    const long long Text = T + 2LL * edge;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, band = tid >> 5;
    double* ring = iir_smem + band * (kRing * kSeq * kLdIn);
    const int r0 = lane >> 4, c = lane & 15;   // staging: rows r0, r0 + 2, ..., column c
    const long long n_tiles = (Text + kTT - 1) / kTT;
    const long long n_groups = (n_seq + kSeq - 1) / kSeq;
    double cr[kMaxSec * 6];
    load_coefficients<FORM>(cf, band, cr);
    // groups are handed out through a device counter (the first gridDim.x statically): CTAs that run
    // ahead of the others take more of them, which keeps the last wave short
    __shared__ int grp_next;
    for (long long grp = blockIdx.x; grp < n_groups;) {
        const long long seq0 = grp * kSeq;
        const bool active = seq0 + lane < n_seq;
        const int rows_here = (int)(n_seq - seq0 < kSeq ? n_seq - seq0 : kSeq);
        const double* x_row0 = x + seq0 * x_stride;
        double* mid_g = mid + grp * Text * nth;
        __syncwarp();   // the previous group's last tile has been read
        double z[8];
        if (active) {
            const double scale = ext_value(x_row0 + lane * x_stride, T, edge, 0);
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = __dmul_rn(cf.zi[band][k], scale);
        }
        // stage tile `tile` (32 rows x 16 samples of x, odd-extended at the ends) into ring slot `slot`
        auto stage = [&](long long tile, int slot) {
            double* dst = ring + slot * (kSeq * kLdIn) + r0 * kLdIn + c;
            const long long k = tile * kTT + c;
            const bool interior = tile * kTT >= edge && tile * kTT + kTT <= edge + T;
            if (k < Text) {
                if (interior) {
                    const double* src = x_row0 + r0 * x_stride + (k - edge);
#pragma unroll
                    for (int i = 0; i < kSeq / 2; ++i) {
                        if (r0 + 2 * i < rows_here) cp_async_8(dst, src);
                        src += 2 * x_stride;
                        dst += 2 * kLdIn;
                    }
                } else {
#pragma unroll 4
                    for (int i = 0; i < kSeq / 2; ++i) {
                        const int r = r0 + 2 * i;
                        if (r < rows_here) dst[2 * i * kLdIn] = ext_value(x_row0 + r * x_stride, T, edge, k);
                    }
                }
            }
            cp_async_commit();
        };
        // (every call commits one cp.async group, so "all but the newest group" = the tile about to be used)
        stage(0, 0);
        if (n_tiles > 1) stage(1, 1); else cp_async_commit();
        int buf = 0;
        for (long long q = 0; q < n_tiles; ++q) {
            const long long k0 = q * kTT;
            cp_async_wait<1>();
            __syncwarp();   // tile q is in slot buf; every lane has left the slot tile q + 2 goes to
            if (q + 2 < n_tiles) stage(q + 2, buf == 0 ? kRing - 1 : buf - 1);
            else cp_async_commit();
            if (active) {
                const double* ti = ring + buf * (kSeq * kLdIn) + lane * kLdIn;
                double* to = mid_g + k0 * nth + tid;
                const int nvalid = (int)((Text - k0 < kTT) ? (Text - k0) : kTT);
                if (nvalid == kTT) {
                    // a full tile is straight-line code: the sections of consecutive samples overlap
                    double v[kTT];
#pragma unroll
                    for (int cc = 0; cc < kTT; ++cc) v[cc] = ti[cc];
#pragma unroll
                    for (int cc = 0; cc < kTT; ++cc) to[(long long)cc * nth] = step_regs<FORM, NC>(z, cr, cf.n, v[cc]);
                } else {
                    for (int cc = 0; cc < nvalid; ++cc) to[(long long)cc * nth] = step_regs<FORM, NC>(z, cr, cf.n, ti[cc]);
                }
            }
            buf = buf + 1 == kRing ? 0 : buf + 1;
        }
        __syncthreads();   // (the only CTA-wide barriers: two per group of 15,054 samples)
        if (tid == 0) grp_next = (int)gridDim.x + atomicAdd(next_group, 1);
        __syncthreads();
        grp = grp_next;
    }
}

// Backward pass: walks the tiles of mid in reverse, writes the un-padded samples of
//   y[(band * n_seq + seq) * T + k - edge].   Warps autonomous as in the forward pass: a warp streams
// its own 32 columns of every time row (256 contiguous bytes), filters them in place and writes its 32
// job rows through a transposed read of its tile (lanes along time: 128-byte row segments).
constexpr int kLdB = kSeq + 2;   // row stride of a backward tile: rows stay 16-byte aligned, transposed reads 2-way
template <int FORM, int NC>
__global__ void __launch_bounds__(kSeq * kMaxBands, 2) iir_backward_kernel(const double* __restrict__ mid, double* __restrict__ y,
                                                                           long long n_seq, long long T, int edge,
                                                                           int* __restrict__ next_group,
                                                                           const __grid_constant__ Coef cf) {
    extern __shared__ __align__(16) double iir_smem[];   // per warp: ring of kRing tiles [16 time rows][kLdB]
    const long long Text = T + 2LL * edge;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, band = tid >> 5;
    double* ring = iir_smem + band * (kRing * kTT * kLdB);
    const int r0 = lane >> 4, c = lane & 15;
    const long long n_tiles = (Text + kTT - 1) / kTT;
    const long long n_groups = (n_seq + kSeq - 1) / kSeq;
    double cr[kMaxSec * 6];
    load_coefficients<FORM>(cf, band, cr);
    // groups are handed out through a device counter (the first gridDim.x statically): CTAs that run
    // ahead of the others take more of them, which keeps the last wave short
    __shared__ int grp_next;
    for (long long grp = blockIdx.x; grp < n_groups;) {
        const long long seq0 = grp * kSeq;
        const bool active = seq0 + lane < n_seq;
        const int rows_here = (int)(n_seq - seq0 < kSeq ? n_seq - seq0 : kSeq);
        const double* mid_w = mid + grp * Text * nth + band * kSeq;   // this warp's 32 columns
        double* y_row0 = y + ((long long)band * n_seq + seq0) * T;
        __syncwarp();   // the previous group's last tile has been stored
        double z[8];
        if (active) {
            const double scale = mid_w[(Text - 1) * nth + lane];
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = __dmul_rn(cf.zi[band][k], scale);
        }
        auto stage = [&](long long tile, int slot) {   // 16 time rows of 32 doubles: lane = (row parity, 16-byte piece)
            double* dst = ring + slot * (kTT * kLdB) + r0 * kLdB + 2 * c;
            const double* src = mid_w + (tile * kTT + r0) * nth + 2 * c;
            const int rows = (int)((Text - tile * kTT < kTT) ? (Text - tile * kTT) : kTT);
#pragma unroll
            for (int i = 0; i < kTT / 2; ++i) {
                if (r0 + 2 * i < rows) cp_async_16(dst, src);
                src += 2 * (long long)nth;
                dst += 2 * kLdB;
            }
            cp_async_commit();
        };
        stage(n_tiles - 1, 0);
        if (n_tiles > 1) stage(n_tiles - 2, 1); else cp_async_commit();
        int buf = 0;
        for (long long q = 0; q < n_tiles; ++q) {
            const long long tile = n_tiles - 1 - q;
            const long long k0 = tile * kTT;
            cp_async_wait<1>();
            __syncwarp();   // tile q is in slot buf; the store of tile q - 1 has read its slot
            if (q + 2 < n_tiles) stage(tile - 2, buf == 0 ? kRing - 1 : buf - 1);
            else cp_async_commit();
            double* slot = ring + buf * (kTT * kLdB);
            if (active) {
                double* ti = slot + lane;   // column of this job; filtered in place
                const int nvalid = (int)((Text - k0 < kTT) ? (Text - k0) : kTT);
                if (nvalid == kTT) {
                    double v[kTT];
#pragma unroll
                    for (int cc = 0; cc < kTT; ++cc) v[cc] = ti[cc * kLdB];
#pragma unroll
                    for (int cc = kTT - 1; cc >= 0; --cc) ti[cc * kLdB] = step_regs<FORM, NC>(z, cr, cf.n, v[cc]);
                } else {
                    for (int cc = nvalid - 1; cc >= 0; --cc) ti[cc * kLdB] = step_regs<FORM, NC>(z, cr, cf.n, ti[cc * kLdB]);
                }
            }
            __syncwarp();   // the tile is filtered
            // ---- transposed, coalesced store: lanes along time (two job rows of 128 bytes per step)
            {
                const long long k = k0 + c;
                if (k >= edge && k < edge + T) {
                    double* dst = y_row0 + r0 * T + (k - edge);
                    const double* sv = slot + c * kLdB + r0;
#pragma unroll
                    for (int i = 0; i < kSeq / 2; ++i) {
                        if (r0 + 2 * i < rows_here) *dst = *sv;
                        dst += 2 * T;
                        sv += 2;
                    }
                }
            }
            buf = buf + 1 == kRing ? 0 : buf + 1;
        }
        __syncthreads();   // (the only CTA-wide barriers: two per group of 15,054 samples)
        if (tid == 0) grp_next = (int)gridDim.x + atomicAdd(next_group, 1);
        __syncthreads();
        grp = grp_next;
    }
}

}  // namespace iir
}  // namespace tda

extern "C" size_t tda_filtfilt_workspace_bytes(long long n_seq, int n_bands, long long T, int padlen) {
    if (n_seq < 0 || n_bands < 1 || T < 1 || padlen < 0) return 0;
    // the padded intermediate, time-major per group of 32 sequences (the last group is padded to 32)
    const long long groups = (n_seq + tda::iir::kSeq - 1) / tda::iir::kSeq;
    return (size_t)groups * tda::iir::kSeq * n_bands * (size_t)(T + 2LL * padlen) * sizeof(double) + 256;   // + work counters
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
    const long long groups = (n_seq + kSeq - 1) / kSeq;
    const int nth = kSeq * n_bands;
    const int smem_f = n_bands * kRing * kSeq * kLdIn * (int)sizeof(double);   // a ring per warp
    const int smem_b = n_bands * kRing * kTT * kLdB * (int)sizeof(double);
    // resident CTAs per SM: registers (<= 128 per thread) and the backward pass's staging
    int per_sm = 65536 / (nth * 128);
    const int by_smem = (227 * 1024) / ((smem_f > smem_b ? smem_f : smem_b) + 1024);
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    const long long maxb = (long long)sms * per_sm;
    const int grid = (int)(groups < maxb ? groups : maxb);
    cudaStream_t st = (cudaStream_t)stream;
    double* mid = (double*)ws;
    int* counters = (int*)((char*)ws + tda_filtfilt_workspace_bytes(n_seq, n_bands, T, padlen) - 256);
    if (cudaError_t e0 = cudaMemsetAsync(counters, 0, 256, st); e0 != cudaSuccess) return (int)e0;
    // sections / taps known at compile time for the filters of the pipeline (order-4 band-pass: 4 sections
    // or 9 taps; order-4 low-pass: 5 taps); any other count runs the same kernels with a run-time count
    auto launch = [&](auto fwd, auto bwd) -> int {
        cudaFuncSetAttribute(fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f);
        cudaFuncSetAttribute(bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
        {
            tda::ProfScope prof(form == 0 ? "iir_sos_forward" : "iir_ba_forward", st);
            fwd<<<grid, nth, smem_f, st>>>(x, mid, n_seq, T, x_stride, padlen, counters, cf);
            tda::count_launch();
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        {
            tda::ProfScope prof(form == 0 ? "iir_sos_backward" : "iir_ba_backward", st);
            bwd<<<grid, nth, smem_b, st>>>(mid, y, n_seq, T, padlen, counters + 1, cf);
            tda::count_launch();
        }
        return (int)cudaGetLastError();
    };
    if (form == 0 && n == 4) return launch(iir_forward_kernel<0, 4>, iir_backward_kernel<0, 4>);
    if (form == 0) return launch(iir_forward_kernel<0, 0>, iir_backward_kernel<0, 0>);
    if (n == 9) return launch(iir_forward_kernel<1, 9>, iir_backward_kernel<1, 9>);
    if (n == 5) return launch(iir_forward_kernel<1, 5>, iir_backward_kernel<1, 5>);
    return launch(iir_forward_kernel<1, 0>, iir_backward_kernel<1, 0>);
}
