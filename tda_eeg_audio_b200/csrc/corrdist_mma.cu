// corrdist_mma.cu — the window -> correlation -> distance kernel with the Gram on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind, and FP64 is what the accuracy needs:
// d = sqrt(2(1-r)) cancels as r -> 1, corrdist.cu / SURVEY.md §7.2 H3).  Default for every shape it
// takes; measured 2.1x the register-tile FMA kernel of corrdist.cu with bit-identical distances and
// correlations (profiles/r02_staged_ab.jsonl).
//
// Same interface and epilogue arithmetic as corrdist_kernel (corrdist.cu), which it replaces per window for:
//   create_sliding_windows      /root/reference/notebooks/1_preprocesamiento.ipynb:314-364 (slicing only)
//   compute_correlation_matrix  /root/reference/notebooks/2_graph_construction.ipynb:86-97
//   correlation_to_distance     /root/reference/notebooks/2_graph_construction.ipynb:100-122
//
// Two persistent CTAs of eight warps per SM, each with ONE window buffer (47 rows of 250 float64 = 94 KB)
// that the TMA engine fills: lane 0 of every warp issues a bulk asynchronous copy per channel row
// (cp.async.bulk.shared::cluster.global, 2,000 contiguous bytes each) that completes on an mbarrier.
// The copies of window k+1 are issued as soon as the tensor-pipe phase has read window k -- the
// covariance block lives in its own 10.5 KB (upper-triangular tiles, packed) -- so they stream in under
// the float64 divisions / square roots of the epilogue, and the two CTAs of an SM drift into
// complementary phases: one multiplies while the other centres, finishes or waits for HBM.  (A single
// double-buffered CTA of sixteen warps ran every warp through the same phase at the same time and left
// the tensor pipe idle two thirds of the time: profiles/r02c_corrdist_ncu.json.)  Per window: centring
// in place (a warp takes its rows four at a time), the 21 upper-triangular 8x8 tiles of X X^T with
// mma.sync.m8n8k4.f64, three tiles in flight per warp (per k-step of 4 samples two conflict-free
// LDS.64 and one DMMA per tile; 1,323 DMMAs per 47-channel window), and the epilogue over the 1,128
// pairs i <= j (a cyclic folding of the matrix gives every thread real pairs, three at a time: no
// half-idle warps in front of the divisions).
// Shapes the bulk copies cannot take (rows not 16-byte aligned, more than one round of tiles, windows
// beyond the shared memory) are served by corrdist_kernel.
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
constexpr int kIlp = 3;                     // tiles a warp keeps in flight

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- mbarrier / bulk-copy primitives (PTX; the async proxy writes shared memory, the barrier counts bytes)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// row stride of the staged window in doubles: a multiple of 4 that is 4 mod 8, so that the eight
// rows x four columns a fragment load touches fall on sixteen distinct 8-byte bank pairs per half-warp
// (and every row starts 16-byte aligned, as the bulk copies need)
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
    const int nt = Cp / 8;
    const int ntiles = nt * (nt + 1) / 2;          // <= kMmaWarps * kIlp (checked by the launcher)
    double* xs = sm;                               // the window, Cp x ldw
    double* cs = sm + (size_t)Cp * ldw;            // covariance, upper-triangular 8x8 tiles packed: ntiles x 64
    double* sd = cs + (size_t)ntiles * 64;         // Cp
    uint64_t* bar = reinterpret_cast<uint64_t*>(sd + Cp);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const long long items = (long long)R * W;
    const uint32_t row_bytes = (uint32_t)win * 8u;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the padding the copies never touch: channel rows C..Cp-1 (whole rows)
    for (int q = tid; q < (Cp - C) * ldw; q += kMmaThreads) xs[(size_t)C * ldw + q] = 0.0;
    fence_proxy_async();
    __syncthreads();

    // lane 0 of every warp issues the bulk copies of its rows (warp, warp + 8, ...): a bulk copy is a
    // warp-level instruction, and 47 of them from one thread delay that thread's warp -- and with it the
    // next barrier -- by several thousand cycles.  Thread 0 arms the barrier with the byte count; its
    // arrival is the only pending one, so the phase cannot complete before it.
    auto issue = [&](long long item) {
        if (lane != 0) return;
        const long long rec_i = item / W;
        const int w_i = (int)(item - rec_i * W);
        const double* src = x + rec_i * strideR + (long long)w_i * step;
        if (warp == 0) mbar_expect_tx(bar, row_bytes * (uint32_t)C);
        for (int c = warp; c < C; c += kMmaWarps) bulk_g2s(xs + (size_t)c * ldw, src + (long long)c * T, row_bytes, bar);
    };
    if ((long long)blockIdx.x < items) issue(blockIdx.x);

    // tile assignment of this warp (fixed for the whole kernel)
    int tile_i[kIlp], tile_j[kIlp], tile_l[kIlp];
#pragma unroll
    for (int u = 0; u < kIlp; ++u) {
        const int tl = u * kMmaWarps + warp;
        int ti = 0, rem = tl < ntiles ? tl : 0;
        while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
        tile_i[u] = tl < ntiles ? ti : -1;
        tile_j[u] = ti + rem;
        tile_l[u] = tl;
    }
    const int half = C / 2 + 1;                    // cyclic folding: offsets 0 .. floor(C/2)
    const int npairs = C * half;
    const float inv_half = 1.0f / (float)half;
    constexpr int kRowsPerWarp = 8;                // rows a warp centres (C <= 64), four at a time
    constexpr int kPairsPerThread = 3;             // pairs a thread finishes at once

    // (recording, window) of the current item, advanced without a division per window
    long long rec = (long long)blockIdx.x / W;
    int w = (int)((long long)blockIdx.x - rec * W);
    const long long step_rec = (long long)gridDim.x / W;
    const int step_w = (int)((long long)gridDim.x - step_rec * W);
    int it = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(bar, (uint32_t)(it & 1));
        // ---- centre each channel (a warp takes its rows four at a time: independent chains), zero the k-padding
        const int nfull = win >> 5, tail = win & 31;   // whole 32-sample steps, then a partial one
#pragma unroll
        for (int q0 = 0; q0 < kRowsPerWarp; q0 += 4) {
            if (warp + kMmaWarps * q0 >= C) break;
            double* rp[4];
            double s[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = warp + kMmaWarps * (q0 + q);
                rp[q] = xs + (size_t)(c < C ? c : 0) * ldw + lane;
                s[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (warp + kMmaWarps * (q0 + q) < C) {
                    const double* p = rp[q];
#pragma unroll 4
                    for (int i = 0; i < nfull; ++i) s[q] += p[32 * i];
                    if (lane < tail) s[q] += p[32 * nfull];
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(0xFFFFFFFFu, s[q], o);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] = s[q] / win;          // the mean, numpy's division
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (warp + kMmaWarps * (q0 + q) < C) {
                    double* p = rp[q];
#pragma unroll 4
                    for (int i = 0; i < nfull; ++i) p[32 * i] -= s[q];
                    if (lane < tail) p[32 * nfull] -= s[q];
                    else if (32 * nfull + lane < Kp) p[32 * nfull] = 0.0;
                }
            }
        }
        __syncthreads();
        // ---- upper-triangular 8x8 tiles of X X^T on the FP64 tensor pipe, kIlp tiles per warp at once
        double acc[kIlp][2];
        {
            const double* pa[kIlp];
            const double* pb[kIlp];
#pragma unroll
            for (int u = 0; u < kIlp; ++u) {
                const int ti = tile_i[u] < 0 ? 0 : tile_i[u], tj = tile_i[u] < 0 ? 0 : tile_j[u];
                pa[u] = xs + (size_t)(ti * 8 + g) * ldw + t4;
                pb[u] = xs + (size_t)(tj * 8 + g) * ldw + t4;
                acc[u][0] = 0.0;
                acc[u][1] = 0.0;
            }
            // (warp-uniform: the 21 tiles of a 47-channel window give five warps three tiles and three warps two)
            if (tile_i[kIlp - 1] >= 0) {
#pragma unroll 4
                for (int k0 = 0; k0 < Kp; k0 += 4) {
#pragma unroll
                    for (int u = 0; u < kIlp; ++u) dmma_8x8x4(acc[u][0], acc[u][1], pa[u][k0], pb[u][k0]);
                }
            } else if (tile_i[kIlp - 2] >= 0) {
#pragma unroll 4
                for (int k0 = 0; k0 < Kp; k0 += 4) {
#pragma unroll
                    for (int u = 0; u < kIlp - 1; ++u) dmma_8x8x4(acc[u][0], acc[u][1], pa[u][k0], pb[u][k0]);
                }
            } else if (tile_i[0] >= 0) {
#pragma unroll 4
                for (int k0 = 0; k0 < Kp; k0 += 4) dmma_8x8x4(acc[0][0], acc[0][1], pa[0][k0], pb[0][k0]);
            }
        }
        // ---- covariance tiles (their own region: the window buffer is free as soon as every warp has
        //      finished reading it, and the next window's rows start streaming in under the epilogue)
        {
            const double inv = 1.0 / (double)(win - 1);
#pragma unroll
            for (int u = 0; u < kIlp; ++u) {
                if (tile_i[u] >= 0) {
                    double* o = cs + (size_t)tile_l[u] * 64 + g * 8 + 2 * t4;
                    o[0] = acc[u][0] * inv;
                    o[1] = acc[u][1] * inv;
                }
            }
        }
        fence_proxy_async();   // generic-proxy writes to the window (centring) before the bulk copies that refill it
        __syncthreads();
        if (item + gridDim.x < items) issue(item + gridDim.x);
        for (int c = tid; c < C; c += kMmaThreads) {
            const int tc = c >> 3;
            sd[c] = sqrt(cs[(size_t)(tc * nt - tc * (tc - 1) / 2) * 64 + (c & 7) * 9]);
        }
        __syncthreads();
        // ---- epilogue over the pairs i <= j: pair (r, (r + o) mod C), o = 0 .. floor(C/2); a thread works
        //      on kPairsPerThread pairs at once (independent division / square-root chains)
        const long long oo = rec * strideO + (long long)w * C * C;
        float* Dw = D ? D + oo : nullptr;
        double* Cw = corr ? corr + oo : nullptr;
        for (int e0 = tid; e0 < npairs; e0 += kMmaThreads * kPairsPerThread) {
            int pi[kPairsPerThread], pj[kPairsPerThread];
            double r[kPairsPerThread], d[kPairsPerThread];
            bool ok[kPairsPerThread];
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) {
                const int e = e0 + q * kMmaThreads;
                const int ec = e < npairs ? e : 0;
                // ec / half without an integer division (exact: ec + 0.5 is half a unit away from every multiple)
                const int r0 = __float2int_rz(((float)ec + 0.5f) * inv_half), o = ec - r0 * half;
                int r1 = r0 + o;
                if (r1 >= C) r1 -= C;
                // even C: the antipodal pairs come twice
                ok[q] = e < npairs && !(C % 2 == 0 && o == C / 2 && r0 >= C / 2);
                pi[q] = r0 < r1 ? r0 : r1;
                pj[q] = r0 < r1 ? r1 : r0;
                const int ti = pi[q] >> 3, tj = pj[q] >> 3;
                r[q] = cs[(size_t)(ti * nt - ti * (ti - 1) / 2 + tj - ti) * 64 + (pi[q] & 7) * 8 + (pj[q] & 7)];
            }
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) r[q] = r[q] / sd[pi[q]];
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) r[q] = r[q] / sd[pj[q]];   // numpy: c /= stddev[:, None]; c /= stddev[None, :]
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) {
                if (r[q] != r[q]) r[q] = 0.0;                  // nan_to_num(nan=0.0): zero-variance channel
                else r[q] = fmin(fmax(r[q], -1.0), 1.0);       // np.clip inside corrcoef
                double t;
                if (method == 0) t = 2.0 * (1.0 - r[q]);
                else if (method == 1) t = 1.0 - fabs(r[q]);
                else if (method == 2) t = 1.0 - r[q];
                else t = 1.0 - r[q] * r[q];
                d[q] = t;
            }
            if (method == 0 || method == 3) {
#pragma unroll
                for (int q = 0; q < kPairsPerThread; ++q) d[q] = sqrt(d[q]);
            }
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) {
                if (!ok[q]) continue;
                const int i = pi[q], j = pj[q];
                if (Cw) { Cw[(size_t)i * C + j] = r[q]; Cw[(size_t)j * C + i] = r[q]; }
                if (Dw) {
                    double dd = fmax(d[q], 0.0);
                    if (i == j) dd = 0.0;
                    const float f = (float)dd;
                    Dw[(size_t)i * C + j] = f;
                    Dw[(size_t)j * C + i] = f;
                }
            }
        }
        __syncthreads();   // cs and sd are free for the next window
        rec += step_rec;
        w += step_w;
        if (w >= W) { w -= W; ++rec; }
    }
}

size_t corrdist_mma_smem_bytes(int C, int win) {
    const int Cp = (C + 7) & ~7;
    const int nt = Cp / 8, ntiles = nt * (nt + 1) / 2;
    return ((size_t)Cp * mma_ldw(win) + (size_t)ntiles * 64 + Cp) * sizeof(double) + 2 * sizeof(uint64_t);
}

// returns a cudaError_t / TDA_E_* like the C-ABI; the caller has validated the arguments.  TDA_E_SIZE:
// a shape this kernel does not take (the caller falls back to corrdist_kernel)
int launch_corrdist_mma(const double* x, int R, int C, long long T, long long strideR, int win, int step, long long W,
                        int method, float* D, double* corr, long long strideO, cudaStream_t stream) {
    const int Cp = (C + 7) & ~7;
    const int nt = Cp / 8, ntiles = nt * (nt + 1) / 2;
    if (ntiles > kMmaWarps * kIlp || C > 64) return TDA_E_SIZE;                        // one round of tiles
    const size_t smem = corrdist_mma_smem_bytes(C, win);
    if (smem > 227 * 1024) return TDA_E_SIZE;
    // bulk copies: 16-byte aligned sources and sizes
    if (((uintptr_t)x & 15) || (T & 1) || (strideR & 1) || (step & 1) || (win & 1)) return TDA_E_SIZE;
    cudaError_t e = cudaFuncSetAttribute(corrdist_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long items = (long long)R * W;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 2) per_sm = 2;
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
