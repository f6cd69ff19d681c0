// corrdist_mma.cu — the window -> correlation -> distance kernel with the Gram on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind, and FP64 is what the accuracy needs:
// d = sqrt(2(1-r)) cancels as r -> 1, corrdist.cu / SURVEY.md §7.2 H3).  Default for every shape it
// takes; measured 2.1x the register-tile FMA kernel of corrdist.cu with bit-identical distances and
// correlations (profiles/r02_staged_ab.jsonl).
//
// Same interface, same load / centring / epilogue arithmetic as corrdist_kernel (corrdist.cu), which
// it replaces per window for:
//   create_sliding_windows      /root/reference/notebooks/1_preprocesamiento.ipynb:314-364 (slicing only)
//   compute_correlation_matrix  /root/reference/notebooks/2_graph_construction.ipynb:86-97
//   correlation_to_distance     /root/reference/notebooks/2_graph_construction.ipynb:100-122
//
// Why: the first generation feeds 4x4 register tiles of DFMA from shared memory, 8 LDS.64 per 16
// DFMA, with 78 of 128 threads busy and one CTA of four warps per SM (the window takes 115 KB): 17 %
// of the FP64 pipe.  Here a CTA of eight warps owns a window, the Gram/covariance block reuses the
// window's shared memory (97 KB -> two CTAs = sixteen warps per SM), and every warp runs up to three
// upper-triangular 8x8 tiles at once: per k-step of 4 samples two conflict-free LDS.64 and one DMMA
// (256 FMAs) per tile, three independent accumulator chains.  1,323 DMMAs per 47-channel window.
//
// Fragment layout of mma.m8n8k4.f64 (PTX ISA; CuTe SM80_8x8x4_F64F64F64F64_TN, layouts SM80_8x4 /
// SM80_8x8_Row): g = lane >> 2, t = lane & 3.  A (8x4, row): a = A[g][t].  B (4x8, col): b = B[t][g].
// C/D (8x8): c0 = C[g][2t], c1 = C[g][2t + 1].  With A = rows ti*8.. of X and B = (rows tj*8.. of X)^T,
// both operands are X[row0 + g][k0 + t]: the same address pattern for a and b.
//
// Exactness that matters downstream: every Gram entry accumulates the same k-chunks in the same order
// through the same instruction, so two identical channels give c_ii = c_ij = c_jj bit for bit and
// r = 1 exactly (d = 0), as with the serial FMA loop.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace corrdist {

constexpr int kMmaThreads = 256;
constexpr int kMmaWarps = kMmaThreads / 32;
constexpr int kIlp = 3;   // tiles a warp keeps in flight

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// row stride of the staged window in doubles: a multiple of 4 that is 4 mod 8, so that the eight
// rows x four columns a fragment load touches fall on sixteen distinct 8-byte bank pairs per half-warp
__host__ __device__ inline int mma_ldw(int win) {
    int l = (win + 3) & ~3;
    if ((l & 7) != 4) l += 4;
    return l;
}

__global__ void __launch_bounds__(kMmaThreads, 2)
corrdist_mma_kernel(const double* __restrict__ x, int R, int C, long long T, long long strideR, int win, int step,
                    int W, int method, float* __restrict__ D, double* __restrict__ corr, long long strideO) {
    extern __shared__ __align__(16) double sm[];
    const int Cp = (C + 7) & ~7;        // channels padded to the 8x8 tile
    const int Kp = (win + 3) & ~3;      // samples padded to the k-step (zeros)
    const int ldw = mma_ldw(win);
    double* xs = sm;                    // Cp x ldw, later overlaid by cs
    double* cs = sm;                    // Cp x Cp covariance (written after the last read of xs)
    double* sd = sm + (size_t)Cp * ldw; // Cp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int nt = Cp / 8;
    const int ntiles = nt * (nt + 1) / 2;

    for (long long item = blockIdx.x; item < (long long)R * W; item += gridDim.x) {
        const long long rec = item / W;
        const int w = (int)(item % W);
        const double* src = x + rec * strideR + (long long)w * step;
        // ---- load (coalesced along time) and centre each channel: the arithmetic of corrdist_kernel
        for (int c = warp; c < Cp; c += kMmaWarps) {
            double* row = xs + (size_t)c * ldw;
            if (c < C) {
                const double* gsrc = src + (long long)c * T;
                double s = 0;
                for (int k = lane; k < win; k += 32) { double v = gsrc[k]; row[k] = v; s += v; }
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                const double mean = s / win;
                for (int k = lane; k < win; k += 32) row[k] -= mean;
                for (int k = win + lane; k < Kp; k += 32) row[k] = 0.0;
            } else {
                for (int k = lane; k < Kp; k += 32) row[k] = 0.0;
            }
        }
        __syncthreads();
        // ---- upper-triangular 8x8 tiles of X X^T on the FP64 tensor pipe, kIlp tiles per warp at once
        double acc[kIlp][2];
        int tile_i[kIlp], tile_j[kIlp];
        const int rounds = (ntiles + kMmaWarps * kIlp - 1) / (kMmaWarps * kIlp);
        // (one round for 47 channels: 21 tiles over 8 warps x 3; the accumulators of a round are
        //  parked in registers until xs may be overwritten, so more than one round needs cs elsewhere)
        for (int rd = 0; rd < rounds; ++rd) {
            const double* pa[kIlp];
            const double* pb[kIlp];
#pragma unroll
            for (int u = 0; u < kIlp; ++u) {
                const int tl = (rd * kIlp + u) * kMmaWarps + warp;
                int ti = 0, rem = tl < ntiles ? tl : 0;
                while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
                tile_i[u] = tl < ntiles ? ti : -1;
                tile_j[u] = ti + rem;
                pa[u] = xs + (size_t)(ti * 8 + g) * ldw + t4;
                pb[u] = xs + (size_t)((ti + rem) * 8 + g) * ldw + t4;
                acc[u][0] = 0.0;
                acc[u][1] = 0.0;
            }
            for (int k0 = 0; k0 < Kp; k0 += 4) {
#pragma unroll
                for (int u = 0; u < kIlp; ++u) {
                    if (tile_i[u] >= 0) dmma_8x8x4(acc[u][0], acc[u][1], pa[u][k0], pb[u][k0]);   // warp-uniform
                }
            }
            if (rounds > 1) {
                // general shapes: the covariance block lives behind the window instead of over it
                double* cs2 = sd + Cp;
                const double inv = 1.0 / (double)(win - 1);
#pragma unroll
                for (int u = 0; u < kIlp; ++u) {
                    if (tile_i[u] >= 0) {
                        double* o = cs2 + (size_t)(tile_i[u] * 8 + g) * Cp + tile_j[u] * 8 + 2 * t4;
                        o[0] = acc[u][0] * inv;
                        o[1] = acc[u][1] * inv;
                    }
                }
            }
        }
        __syncthreads();   // every warp has finished reading xs
        if (rounds == 1) {
            const double inv = 1.0 / (double)(win - 1);
#pragma unroll
            for (int u = 0; u < kIlp; ++u) {
                if (tile_i[u] >= 0) {
                    double* o = cs + (size_t)(tile_i[u] * 8 + g) * Cp + tile_j[u] * 8 + 2 * t4;
                    o[0] = acc[u][0] * inv;
                    o[1] = acc[u][1] * inv;
                }
            }
        }
        const double* cov = rounds == 1 ? cs : sd + Cp;
        __syncthreads();
        // sd must not alias cov: for rounds == 1 it sits behind the window, for rounds > 1 in front of cs2
        for (int c = tid; c < C; c += kMmaThreads) sd[c] = sqrt(cov[(size_t)c * Cp + c]);
        __syncthreads();
        // ---- epilogue over i <= j: the arithmetic of corrdist_kernel
        const long long oo = rec * strideO + (long long)w * C * C;
        float* Dw = D ? D + oo : nullptr;
        double* Cw = corr ? corr + oo : nullptr;
        for (int e = tid; e < C * C; e += kMmaThreads) {
            const int i = e / C, j = e % C;
            if (i > j) continue;
            double r = cov[(size_t)i * Cp + j];
            r = r / sd[i];
            r = r / sd[j];                       // numpy: c /= stddev[:, None]; c /= stddev[None, :]
            if (r != r) r = 0.0;                 // nan_to_num(nan=0.0): zero-variance channel
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

size_t corrdist_mma_smem_bytes(int C, int win) {
    const int Cp = (C + 7) & ~7;
    const int nt = Cp / 8, ntiles = nt * (nt + 1) / 2;
    const int rounds = (ntiles + kMmaWarps * kIlp - 1) / (kMmaWarps * kIlp);
    size_t doubles = (size_t)Cp * mma_ldw(win) + Cp;
    if (rounds > 1) doubles += (size_t)Cp * Cp;
    if (rounds == 1 && (size_t)Cp * Cp > (size_t)Cp * mma_ldw(win)) doubles = (size_t)Cp * Cp + Cp;   // tiny windows
    return doubles * sizeof(double);
}

// returns a cudaError_t / TDA_E_* like the C-ABI; the caller has validated the arguments
int launch_corrdist_mma(const double* x, int R, int C, long long T, long long strideR, int win, int step, long long W,
                        int method, float* D, double* corr, long long strideO, cudaStream_t stream) {
    const size_t smem = corrdist_mma_smem_bytes(C, win);
    if (smem > 227 * 1024) return TDA_E_SIZE;
    if ((size_t)((C + 7) & ~7) * ((C + 7) & ~7) > (size_t)((C + 7) & ~7) * mma_ldw(win)) return TDA_E_SIZE;   // cs must fit over xs
    cudaError_t e = cudaFuncSetAttribute(corrdist_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)((228 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    long long items = (long long)R * W;
    long long grid = (long long)sms * per_sm;
    if (grid > items) grid = items;
    tda::ProfScope prof("corrdist_mma", stream);
    corrdist_mma_kernel<<<(int)grid, kMmaThreads, smem, stream>>>(x, R, C, T, strideR, win, step, (int)W, method, D, corr,
                                                                 strideO);
    tda::count_launch();
    return (int)cudaGetLastError();
}

}  // namespace corrdist
}  // namespace tda
