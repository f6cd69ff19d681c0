// wasserstein.cu — batched EXACT 1-Wasserstein distance between persistence diagrams
// (L2 ground metric, Euclidean distance to the diagonal), float64.
//
// Replaces safe_wasserstein -> persim.wasserstein(dgm1, dgm2)
//   /root/reference/scripts/utils.py:180-191 (alias at utils.py:12), called per window by
//   /root/reference/scripts/tda_eeg_audio_comparison.py:95-96 and
//   /root/reference/scripts/matched_vs_mismatched.py:87-95 (compute_cross_wasserstein).
//
// persim builds an (M+N)x(M+N) cost matrix (points x points, points x own diagonal copy, +inf
// elsewhere, 0 diagonal-to-diagonal) and calls scipy's linear_sum_assignment.
// That optimum equals  sum_j dt_j + min over assignments of the M points of S to either a point
// of T (reduced cost c_ij - dt_j) or their own diagonal (cost ds_i): an M x (N+M) rectangular
// assignment problem whose diagonal block is implicit.  One warp solves one pair with the
// shortest-augmenting-path (Hungarian with potentials) algorithm: points, potentials and the column
// flags live in the warp's shared memory in float64 (M <= N after an optional role swap), the L2 costs
// are evaluated where they are needed, lanes stride over the columns for the relax + arg-min step.  The optimum of an LSAP is unique in value, so
// the result matches scipy's to rounding.
//
// H0 diagrams are one-dimensional: every birth is 0 and the deaths come sorted.  For such a pair (all
// births of both diagrams equal, deaths non-decreasing -- checked per pair, anything else takes the
// general solver) the ground costs have the Monge property, so an optimal matching does not cross and
// the same optimum is the end of a dynamic programme over the two sorted lists,
//   f[i][j] = min(f[i-1][j] + ds_i, f[i][j-1] + dt_j, f[i-1][j-1] + c_ij),
// one row at a time with the lanes along j: f[i][j] = P_j + min_{k <= j}(g_k - P_k), P = prefix sums of
// dt, g_k = min(f[i-1][k] + ds_i, f[i-1][k-1] + c_ik) -- a prefix minimum, i.e. a warp scan.  The c_ij
// are the very same Gram-trick values.  46 x 113 cells instead of O(M^2 N) augmentation steps:
// 19,200 EEG-vs-audio H0 pairs in 13.1 ms before (profiles/r02_wasserstein_h0_ncu.json).
//
// Shared memory is linear in the number of points of a pair (up to ~4,000 points): the solver's costs are not
// stored.  Round 2 first kept the M x N cost block of a pair in shared memory -- 93 KB for 46 EEG bars against
// 248 audio bars, two warps per SM -- and measured the on-the-fly variant (a dozen warps per SM and more) faster
// on every batch of the pipeline, with bit-identical results (the same expression); the block is gone.
//
// Two launches per batch (MODE 1, then MODE 2): the first solves every one-dimensional pair with the dynamic
// programme and marks the others in `out` with a NaN of its own; the second runs the solver on the marked pairs
// only.  A batch of H0 diagrams never reaches the solver, and a warp of the first launch is not held up behind a
// neighbour's O(M^2 N) augmentations.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace wasserstein {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr double kInf = 1e300;

template <typename TIn> struct Params {
    const TIn* bdA; const int* nA; int nA_stride, capA, limA;
    const TIn* bdB; const int* nB; int nB_stride, capB, limB;
    const int* idxA; const int* idxB;  // optional gather indices (pair k = A[idxA[k]] vs B[idxB[k]])
    long long B;
    double* out;
    int rows_cap, cols_cap;  // min / max of the two caps (+1 for the placeholder point)
};

__host__ __device__ inline size_t smem_bytes(int rows_cap, int cols_cap) {
    size_t s = 0;
    s += ((size_t)cols_cap + 2) * 8;               // one row of the 1-D programme (cols + 1 entries)
    s += (size_t)2 * (rows_cap + cols_cap) * 8;    // points S, T  (x, y)
    s += (size_t)(rows_cap + cols_cap) * 8;        // ds, dt
    s += (size_t)(rows_cap + 1) * 8;               // u
    s += (size_t)2 * (rows_cap + cols_cap + 1) * 8; // v, minv
    s += (size_t)2 * (rows_cap + cols_cap + 1) * 4; // p, way
    s += (size_t)(rows_cap + cols_cap + 4);        // column-used flags
    return (s + 15) & ~(size_t)15;
}

__device__ __forceinline__ double ground_cost(const double* S, const double* T, int i, int j) {
    // L2 cost with sklearn's Gram trick: sqrt(max(|s|^2 - 2 s.t + |t|^2, 0))
    const double sx = S[2 * i], sy = S[2 * i + 1], tx = T[2 * j], ty = T[2 * j + 1];
    const double ns = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
    const double nt = __dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty));
    const double dot = fma(sy, ty, __dmul_rn(sx, tx));
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), ns), nt);
    return sqrt(fmax(d2, 0.0));
}

// compact the finite rows of a diagram into (x, y) float64 pairs; empty -> one point (0,0)
// (TIn = float: the Rips engines' own output, exact float32 values; TIn = double: arbitrary diagrams
// of the drop-in call, which persim treats in float64)
template <typename TIn>
__device__ int load_diagram(const TIn* __restrict__ bd, int n, int cap, double* pts, int lane) {
    if (n > cap) n = cap;
    int m = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        double2 r = make_double2(0.0, 0.0);
        bool ok = false;
        if (k < n) { r.x = (double)bd[2 * k]; r.y = (double)bd[2 * k + 1]; ok = isfinite(r.x) && isfinite(r.y); }
        const uint32_t bal = __ballot_sync(kFull, ok);
        if (ok) {
            const int pos = m + __popc(bal & ((1u << lane) - 1u));
            pts[2 * pos] = r.x;
            pts[2 * pos + 1] = r.y;
        }
        m += __popc(bal);
    }
    if (m == 0) {
        if (lane == 0) { pts[0] = 0.0; pts[1] = 0.0; }
        m = 1;
    }
    __syncwarp();
    return m;
}

// "not a one-dimensional pair: left to the assignment solver" (a quiet NaN no computation produces: costs are finite)
__device__ __forceinline__ double pending_mark() { return __longlong_as_double(0x7FF8DEADBEEF0001LL); }
__device__ __forceinline__ bool is_pending(double v) { return __double_as_longlong(v) == 0x7FF8DEADBEEF0001LL; }

// MODE 1: one-dimensional pairs only, the others are marked; 2: the marked pairs only
template <typename TIn, int MODE>
__global__ void __launch_bounds__(32) wasserstein_kernel(Params<TIn> p) {
    extern __shared__ __align__(16) unsigned char wsm[];
    const int lane = threadIdx.x;
    const int RC = p.rows_cap, CC = p.cols_cap;
    double* cost = (double*)wsm;
    double* PA = cost + (size_t)CC + 2;   // points of A, then points of B right behind them
    double* dS = PA + 2 * (size_t)(RC + CC);
    double* dT = dS + RC;
    double* u = dT + CC;
    double* v = u + (RC + 1);
    double* minv = v + (RC + CC + 1);
    int* pcol = (int*)(minv + (RC + CC + 1));
    int* way = pcol + (RC + CC + 1);
    unsigned char* usedf = (unsigned char*)(way + (RC + CC + 1));   // column-used flags of the solver
    const double cs = 0.7071067811865476, sn = 0.7071067811865475;  // np.cos(pi/4), np.sin(pi/4)

    for (long long k = blockIdx.x; k < p.B; k += gridDim.x) {
        if constexpr (MODE == 2) {
            double cur = 0.0;
            if (lane == 0) cur = p.out[k];
            if (!is_pending(__shfl_sync(kFull, cur, 0))) continue;
        }
        const long long ia = p.idxA ? p.idxA[k] : k;
        const long long ib = p.idxB ? p.idxB[k] : k;
        // the diagram with fewer rows becomes S (rows of the assignment); the problem is symmetric
        int na = p.nA[ia * p.nA_stride], nb = p.nB[ib * p.nB_stride];
        double* S = PA;
        double* T;
        int M, N;
        {
            // load A then B into scratch, then order by size
            int ma = load_diagram(p.bdA + ia * (long long)p.capA * 2, na, p.limA, PA, lane);
            double* PB = PA + 2 * ma;
            int mb = load_diagram(p.bdB + ib * (long long)p.capB * 2, nb, p.limB, PB, lane);
            if (ma <= mb) { S = PA; T = PB; M = ma; N = mb; }
            else { S = PB; T = PA; M = mb; N = ma; }
        }
        // distances to the diagonal: second coordinate after rotating by 45 degrees (persim)
        for (int i = lane; i < M; i += 32) dS[i] = -S[2 * i] * sn + S[2 * i + 1] * cs;
        for (int j = lane; j < N; j += 32) dT[j] = -T[2 * j] * sn + T[2 * j + 1] * cs;
        __syncwarp();
        // ---- one-dimensional pair (all births equal, deaths sorted)?  then the dynamic programme
        if constexpr (MODE != 2) {
            const double b0 = S[0];
            bool ok = true;
            for (int i = lane; i < M; i += 32) ok &= S[2 * i] == b0 && (i == 0 || S[2 * i + 1] >= S[2 * i - 1]);
            for (int j = lane; j < N; j += 32) ok &= T[2 * j] == b0 && (j == 0 || T[2 * j + 1] >= T[2 * j - 1]);
            if (__all_sync(kFull, ok)) {
                // v -> P (prefix sums of dt), minv / u-region -> the two rows of f
                double* Pre = v;
                double* prev = minv;
                double* cur = cost;            // >= N + 1 doubles: rows_cap * cols_cap >= N, and one of slack below
                const int N1 = N + 1;
                if (lane == 0) {
                    double run = 0.0;
                    Pre[0] = 0.0;
                    for (int j = 0; j < N; ++j) { run += dT[j]; Pre[j + 1] = run; }
                }
                __syncwarp();
                for (int j = lane; j < N1; j += 32) prev[j] = Pre[j];   // f[0][j]: every point of T unmatched
                __syncwarp();
                const int bs = (N1 + 31) / 32;                 // columns per lane (a contiguous block)
                const int j0 = lane * bs, j1 = min(N1, j0 + bs);
                for (int i = 1; i <= M; ++i) {
                    const double sx = S[2 * (i - 1)], sy = S[2 * (i - 1) + 1], da = dS[i - 1];
                    const double ns = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
                    double run = kInf;                         // prefix minimum of g_k - P_k inside the block
                    for (int j = j0; j < j1; ++j) {
                        double g = prev[j] + da;
                        if (j >= 1) {
                            const double tx = T[2 * (j - 1)], ty = T[2 * (j - 1) + 1];
                            const double nt = __dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty));
                            const double dot = fma(sy, ty, __dmul_rn(sx, tx));
                            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), ns), nt);
                            g = fmin(g, prev[j - 1] + sqrt(fmax(d2, 0.0)));
                        }
                        run = fmin(run, g - Pre[j]);
                        cur[j] = run;                          // block-local prefix minimum for now
                    }
                    // exclusive prefix minimum of the block minima over the lanes
                    double carry = run;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double y = __shfl_up_sync(kFull, carry, o);
                        if (lane >= o) carry = fmin(carry, y);
                    }
                    carry = __shfl_up_sync(kFull, carry, 1);
                    if (lane == 0) carry = kInf;
                    for (int j = j0; j < j1; ++j) cur[j] = fmin(cur[j], carry) + Pre[j];
                    __syncwarp();
                    double* t = prev; prev = cur; cur = t;
                }
                if (lane == 0) p.out[k] = prev[N];
                __syncwarp();
                continue;
            }
        }
        if constexpr (MODE == 1) {
            if (lane == 0) p.out[k] = pending_mark();
            __syncwarp();
            continue;
        }
        const int Mc = N + M;  // columns: N real + M private diagonal columns
        for (int j = lane; j <= Mc; j += 32) { v[j] = 0.0; pcol[j] = 0; }
        for (int i = lane; i <= M; i += 32) u[i] = 0.0;
        __syncwarp();
        // ---- Hungarian with potentials (1-based rows/cols, column 0 is the virtual start)
        for (int i = 1; i <= M; ++i) {
            if (lane == 0) pcol[0] = i;
            for (int j = lane; j <= Mc; j += 32) minv[j] = kInf;
            for (int j = lane; j <= Mc; j += 32) usedf[j] = 0;
            __syncwarp();
            int j0 = 0;
            while (true) {
                if (lane == (j0 & 31)) usedf[j0] = 1;
                const int i0 = pcol[j0];
                const double ui0 = u[i0];
                double best = kInf;
                int bestj = -1;
                for (int j = lane; j <= Mc; j += 32) {
                    if (j == 0 || usedf[j] != 0) continue;
                    double a;
                    if (j <= N) a = ground_cost(S, T, i0 - 1, j - 1) - dT[j - 1];
                    else a = (j - N == i0) ? dS[i0 - 1] : kInf;
                    const double cur = a - ui0 - v[j];
                    double mv = minv[j];
                    if (cur < mv) { mv = cur; minv[j] = cur; way[j] = j0; }
                    if (mv < best) { best = mv; bestj = j; }
                }
                // warp arg-min (ties -> smaller column, deterministic)
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const double ob = __shfl_xor_sync(kFull, best, o);
                    const int oj = __shfl_xor_sync(kFull, bestj, o);
                    if (ob < best || (ob == best && oj >= 0 && (bestj < 0 || oj < bestj))) { best = ob; bestj = oj; }
                }
                const double delta = best;
                for (int j = lane; j <= Mc; j += 32) {
                    if (usedf[j] != 0) { u[pcol[j]] += delta; v[j] -= delta; }
                    else minv[j] -= delta;
                }
                __syncwarp();
                j0 = bestj;
                if (pcol[j0] == 0) break;
            }
            // augment along the path
            if (lane == 0) {
                int jj = j0;
                while (jj) { const int j1 = way[jj]; pcol[jj] = pcol[j1]; jj = j1; }
            }
            __syncwarp();
        }
        // ---- total = matched costs + unmatched diagonal costs (persim: sum(D[match]))
        double tot = 0.0;
        for (int j = lane + 1; j <= Mc; j += 32) {
            const int i = pcol[j];
            if (j <= N) tot += (i > 0) ? ground_cost(S, T, i - 1, j - 1) : dT[j - 1];
            else if (i > 0) tot += dS[i - 1];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
        if (lane == 0) p.out[k] = tot;
        __syncwarp();
    }
}

}  // namespace wasserstein
}  // namespace tda

namespace tda {
namespace wasserstein {
template <typename TIn>
static int launch(const TIn* bdA, const int* nA, int nA_stride, int capA, int limA, const TIn* bdB, const int* nB,
                  int nB_stride, int capB, int limB, const int* idxA, const int* idxB, long long B, double* out,
                  void* stream) {
    if (!bdA || !nA || !bdB || !nB || !out || B < 0 || capA < 0 || capB < 0 || nA_stride < 1 || nB_stride < 1)
        return TDA_E_ARG;
    if (B == 0) return 0;
    Params<TIn> p;
    p.bdA = bdA; p.nA = nA; p.nA_stride = nA_stride; p.capA = capA;
    p.bdB = bdB; p.nB = nB; p.nB_stride = nB_stride; p.capB = capB;
    p.idxA = idxA; p.idxB = idxB; p.B = B; p.out = out;
    if (limA <= 0 || limA > capA) limA = capA;
    if (limB <= 0 || limB > capB) limB = capB;
    p.limA = limA; p.limB = limB;
    const int ca = limA < 1 ? 1 : limA, cb = limB < 1 ? 1 : limB;
    p.rows_cap = ca < cb ? ca : cb;
    p.cols_cap = ca < cb ? cb : ca;
    const size_t smem = smem_bytes(p.rows_cap, p.cols_cap);
    if (smem > 227 * 1024) return TDA_E_SIZE;   // more than ~4,000 points in a pair
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    long long grid = (long long)sms * per_sm;
    if (grid > B) grid = B;
    cudaStream_t st = (cudaStream_t)stream;
    tda::ProfScope prof("wasserstein", st);
    cudaError_t e = cudaFuncSetAttribute(wasserstein_kernel<TIn, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(wasserstein_kernel<TIn, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    wasserstein_kernel<TIn, 1><<<(int)grid, 32, smem, st>>>(p);
    tda::count_launch();
    wasserstein_kernel<TIn, 2><<<(int)grid, 32, smem, st>>>(p);
    tda::count_launch();
    return (int)cudaGetLastError();
}
}  // namespace wasserstein
}  // namespace tda

extern "C" int tda_wasserstein_batched(const float* bdA, const int* nA, int nA_stride, int capA, int limA,
                                       const float* bdB, const int* nB, int nB_stride, int capB, int limB,
                                       const int* idxA, const int* idxB, long long B, double* out, void* stream) {
    return tda::wasserstein::launch<float>(bdA, nA, nA_stride, capA, limA, bdB, nB, nB_stride, capB, limB, idxA, idxB,
                                           B, out, stream);
}

extern "C" int tda_wasserstein_batched_f64(const double* bdA, const int* nA, int nA_stride, int capA, int limA,
                                           const double* bdB, const int* nB, int nB_stride, int capB, int limB,
                                           const int* idxA, const int* idxB, long long B, double* out,
                                           void* stream) {
    return tda::wasserstein::launch<double>(bdA, nA, nA_stride, capA, limA, bdB, nB, nB_stride, capB, limB, idxA,
                                            idxB, B, out, stream);
}
