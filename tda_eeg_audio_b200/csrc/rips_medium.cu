// rips_medium.cu — Vietoris–Rips H0+H1 (Z/2) for batches of medium distance matrices
// (64 < N <= 256: the Takens clouds of the audio path, 97-124 points by default, <= 248 with
// subsample=1).
//
// Replaces ripser.ripser(point_cloud / dm, maxdim=1, thresh) as called by
//   /root/reference/scripts/utils.py:123-132 (compute_audio_persistence)
//
// Same algorithm as rips_small.cu (persistent cohomology by cocycle annotation in one sweep over
// the sorted edges, see oracle/pcoh_model.py), mapped one CTA per cloud:
//   * thread v is apex v: the triangles (i, j, v) that enter with edge (i, j) are evaluated by
//     all apexes at once; G = adj[i] & adj[j] (bit rows in shared memory) is recomputed by every
//     thread, so a live edge costs ONE block barrier (a __syncthreads_or that doubles as the
//     "did any cocycle fire" reduction);
//   * CTA-wide stable LSD radix sort (per-warp contiguous segments, __match_any_sync ranking);
//   * for N <= 128 every array of a cloud lives in shared memory (~100 KB, two clouds per SM),
//     above that the edge arrays sit in an L2-resident global scratch;
//   * tie runs replay the exact simplexwise order; cycle-creating run edges whose first cofacet
//     has them as youngest edge are apparent pairs and take no slot, so degenerate inputs
//     (constant windows -> all distances equal) stay cheap;
//   * capacity tiers W=2 (64 classes) -> 8 -> 32 words per edge, device-side hand-over list.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace rips_medium {

constexpr int kMaxN = 254;   // apex ids must leave 254 / 255 free as defv sentinels
constexpr int kRows = 256;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kEssential = 0xFFFFFFFFu;

struct Params {
    const float* D;
    const int* npts;  // per item point count (nullptr => N for all)
    long long strideB;
    int ld, N, B;
    float thresh;
    float* bd0;
    long long* pr0;
    float* bd1;
    long long* pr1;
    int* counts;
    int* status;
    int cap0, cap1;
    const int* worklist;
    const int* n_work;
    int* overflow_list;
    int* n_overflow;
    unsigned char* big;      // per-CTA global scratch for the edge arrays (nullptr => shared memory)
    size_t big_stride;
    uint32_t* phi_global;    // per-CTA PHI scratch for W > 2
    size_t phi_stride;       // in words
};

__host__ __device__ inline int c2(int i) { return i * (i - 1) / 2; }
__host__ __device__ inline int c3(int i) { return i * (i - 1) * (i - 2) / 6; }
__device__ __forceinline__ int edge_q(int a, int b) { return a > b ? c2(a) + b : c2(b) + a; }
__device__ __forceinline__ int tri_index(int x, int y, int z) {
    int a = max(x, max(y, z)), c = min(x, min(y, z)), b = x + y + z - a - c;
    return c3(a) + c2(b) + c;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t float_key(float d) {
    uint32_t u = __float_as_uint(d);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}

template <int MW> __host__ __device__ inline int epad_of(int N) {
    const int nth = 32 * MW;
    return (c2(N) + nth - 1) / nth * nth;
}
template <int W> __host__ __device__ inline int recs_of() { return W <= 2 ? 192 : (W <= 8 ? 512 : 2048); }
// bytes of the "big" edge arrays: [A 4Ep][C 4Ep][B 2Ep][Dd 2Ep]
template <int MW> __host__ __device__ inline size_t big_bytes(int N) { return (size_t)epad_of<MW>(N) * 12; }
template <int MW, int W> __host__ __device__ inline size_t fixed_bytes() {
    size_t s = 0;
    s += (size_t)2 * kRows * MW * 4;   // adj, runadj (only the first N rows used)
    s += (size_t)MW * 256 * 4;         // radix histograms, one per warp
    s += (size_t)3 * recs_of<W>() * 4; // death records
    s += 32 * W * 2;                   // brank
    s += 2 * kRows;                    // comp, eld
    s += 64 + W * 4 + 64;              // wtop, broadcast buffer, misc
    return (s + 15) & ~(size_t)15;
}

template <int MW, int W> struct Cta {
    static constexpr int NTH = 32 * MW;
    // storage
    uint32_t* adj;      // [N][MW]
    uint32_t* runadj;   // [N][MW]
    uint32_t* hist;     // [MW][256]
    uint32_t* rec;      // [3][R]
    uint16_t* brank;
    uint8_t* comp;
    uint8_t* eld;
    int* wtop;          // [MW] + flags
    uint32_t* bc;       // [W] broadcast
    uint32_t* K;        // region A (sort) ; PHI (W==2) spans A and C
    uint32_t* K2;       // region C
    uint16_t* P;        // region B
    uint16_t* P2;       // region Dd (sort) ; afterwards tiebits + defv
    uint32_t* tiebits;
    uint8_t* defv;
    uint32_t* phi;
    const float* Db;
    int tid, lane, warp, N, E, Epad, R, ld;
    // uniform per-window state (replicated in every thread)
    uint32_t live[W], used[W];
    int n0, n1, ncomp, m;
    bool overflow;

    __device__ __forceinline__ float dist(int a, int b) const {
        return __ldg(Db + (size_t)min(a, b) * ld + max(a, b)) + 0.0f;
    }
    __device__ __forceinline__ bool live_any() const {
        uint32_t a = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) a |= live[w];
        return a != 0;
    }
    __device__ __forceinline__ bool tie_after(int r) const { return (tiebits[r >> 5] >> (r & 31)) & 1u; }

    // ------------------------------------------------------------------ slots (uniform)
    __device__ int alloc_slot() {
        for (int attempt = 0; attempt < 2; ++attempt) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t f = ~used[w];
                if (f) {
                    int s = __ffs(f) - 1;
                    used[w] |= 1u << s;
                    live[w] |= 1u << s;
                    return 32 * w + s;
                }
            }
            __syncthreads();
            for (int q = tid; q < E; q += NTH) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi[(size_t)q * W + w] &= live[w];
            }
            __syncthreads();
            bool room = false;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                used[w] = live[w];
                room |= (~used[w]) != 0;
            }
            if (!room) break;
        }
        overflow = true;
        return -1;
    }

    // ------------------------------------------------------------------ H0 (uniform decision)
    __device__ bool h0_step(const Params& p, int b, int i, int j) {
        if (ncomp <= 1) return false;
        const int ci = comp[i], cj = comp[j];
        if (ci == cj) return false;
        const int ei = eld[ci], ej = eld[cj];
        const float d = dist(i, j);
        if (d != 0.0f) {
            if (tid == 0 && n0 < p.cap0) {
                size_t o = ((size_t)b * p.cap0 + n0) * 2;
                p.bd0[o] = 0.0f;
                p.bd0[o + 1] = d;
                if (p.pr0) { p.pr0[o] = min(ei, ej); p.pr0[o + 1] = c2(i) + j; }
            }
            ++n0;
        }
        __syncthreads();
        for (int v = tid; v < N; v += NTH)
            if (comp[v] == ci) comp[v] = (uint8_t)cj;
        if (tid == 0) eld[cj] = (uint8_t)max(ei, ej);
        __syncthreads();
        --ncomp;
        return true;
    }

    __device__ __forceinline__ void add_adj(int i, int j, bool run) {
        if (tid == 0) {
            adj[i * MW + (j >> 5)] |= 1u << (j & 31);
            adj[j * MW + (i >> 5)] |= 1u << (i & 31);
            if (run) {
                runadj[i * MW + (j >> 5)] |= 1u << (j & 31);
                runadj[j * MW + (i >> 5)] |= 1u << (i & 31);
            }
        }
    }

    // ------------------------------------------------------------------ deaths inside a group
    // thread v holds c = coboundary masks on triangle (a, b, v); `isdef` threads carry the value of
    // an apparent edge they define (updated linearly, never a death candidate)
    __device__ void resolve(int a, int b, float dcur, uint32_t (&c)[W], bool isdef) {
        while (true) {
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) any |= c[w];
            const bool nz = any != 0 && !isdef;
            const uint32_t bal = __ballot_sync(kFull, nz);
            if (lane == 0) wtop[warp] = bal ? (32 * warp + 31 - __clz(bal)) : -1;
            __syncthreads();
            int v = -1;
#pragma unroll
            for (int w = 0; w < MW; ++w) v = max(v, wtop[w]);
            if (v < 0) { __syncthreads(); return; }
            if (tid == v) {
#pragma unroll
                for (int w = 0; w < W; ++w) bc[w] = c[w];
            }
            __syncthreads();
            uint32_t cv[W];
#pragma unroll
            for (int w = 0; w < W; ++w) cv[w] = bc[w];
            int slot = -1, age = -1;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = cv[w];
                while (bits) {
                    int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    int ag = brank[32 * w + s];
                    if (ag > age) { age = ag; slot = 32 * w + s; }
                }
            }
            const int sw = slot >> 5;
            const uint32_t sb = 1u << (slot & 31);
            const int bp = P[age];
            if (dist(bp >> 8, bp & 255) != dcur) {
                if (n1 < R) {
                    if (tid == 0) {
                        rec[n1] = (uint32_t)age;
                        rec[R + n1] = __float_as_uint(dcur);
                        rec[2 * R + n1] = (uint32_t)tri_index(a, b, v);
                    }
                    ++n1;
                } else {
                    overflow = true;
                    __syncthreads();
                    return;
                }
            }
            bool absorb = false;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (w == sw) { live[w] &= ~sb; absorb |= (cv[w] & ~sb) != 0; }
                else absorb |= cv[w] != 0;
            }
            bool has = false;
#pragma unroll
            for (int w = 0; w < W; ++w) if (w == sw) has = (c[w] & sb) != 0;
            if (has) {
#pragma unroll
                for (int w = 0; w < W; ++w) c[w] ^= cv[w];
            }
            if (absorb) {
                for (int q = tid; q < E; q += NTH) {
                    uint32_t* e = phi + (size_t)q * W;
                    if (e[sw] & sb) {
#pragma unroll
                        for (int w = 0; w < W; ++w) e[w] ^= cv[w];
                    }
                }
            }
            __syncthreads();
        }
    }

    // Evaluate the triangles (a, b, v), v in G, in descending v.  vdef >= 0: (a,b) is an apparent
    // edge defined by apex vdef (the top of G).  in_run: inside a tie run, where an apex may also
    // define one of its other two edges.  Contains exactly one unconditional block barrier.
    __device__ void group_eval(int a, int b, const uint32_t (&G)[MW], int vdef, bool in_run, float dcur) {
        const int q_ab = c2(a) + b;
        uint32_t pe[W];
        if (vdef >= 0) {
            const uint32_t* x = phi + (size_t)edge_q(a, vdef) * W;
            const uint32_t* y = phi + (size_t)edge_q(b, vdef) * W;
#pragma unroll
            for (int w = 0; w < W; ++w) pe[w] = (x[w] ^ y[w]) & live[w];
            if (tid == 0) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi[(size_t)q_ab * W + w] = pe[w];
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; ++w) pe[w] = phi[(size_t)q_ab * W + w];
        }
        const int v = tid;
        uint32_t c[W];
#pragma unroll
        for (int w = 0; w < W; ++w) c[w] = 0;
        int defq = -1;
        bool in = false;
#pragma unroll
        for (int w = 0; w < MW; ++w) if (w == (v >> 5)) in = (G[w] >> (v & 31)) & 1u;
        if (in && v != vdef) {
            const int qa = edge_q(a, v), qb = edge_q(b, v);
            const uint32_t* x = phi + (size_t)qa * W;
            const uint32_t* y = phi + (size_t)qb * W;
#pragma unroll
            for (int w = 0; w < W; ++w) c[w] = (pe[w] ^ x[w] ^ y[w]) & live[w];
            if (in_run) {
                if (defv[qa] == b) defq = qa;
                else if (defv[qb] == a) defq = qb;
            }
        }
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) any |= c[w];
        const int fired = __syncthreads_or(any != 0 && defq < 0);
        if (fired) resolve(a, b, dcur, c, defq >= 0);
        if (in_run) {
            if (defq >= 0) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi[(size_t)defq * W + w] = c[w] & live[w];
            }
            __syncthreads();
        }
    }

    // ------------------------------------------------------------------ one single (untied) edge
    __device__ void single_edge(const Params& p, int b, int r) {
        const int pij = P[r];
        const int i = pij >> 8, j = pij & 255;
        const int q = c2(i) + j;
        const bool merging = h0_step(p, b, i, j);
        uint32_t G[MW];
        uint32_t anyG = 0;
#pragma unroll
        for (int w = 0; w < MW; ++w) { G[w] = adj[i * MW + w] & adj[j * MW + w]; anyG |= G[w]; }
        add_adj(i, j, false);  // the new bits cannot appear in this edge's own G
        if (merging || (anyG && !live_any())) {
            if (tid < W) phi[(size_t)q * W + tid] = 0;
            __syncthreads();
            return;
        }
        if (!anyG) {
            const int s = alloc_slot();
            if (s < 0) return;
            if (tid == 0) brank[s] = (uint16_t)r;
            if (tid < W) phi[(size_t)q * W + tid] = (tid == (s >> 5)) ? (1u << (s & 31)) : 0u;
            __syncthreads();
            return;
        }
        int vtop = 0;
#pragma unroll
        for (int w = 0; w < MW; ++w) if (G[w]) vtop = 32 * w + 31 - __clz(G[w]);
        group_eval(i, j, G, vtop, false, dist(i, j));
    }

    // is edge (x,y) strictly earlier in the filtration than the run edge with index idx_e ?
    __device__ __forceinline__ bool earlier(int x, int y, int idx_e) const {
        if (!((runadj[x * MW + (y >> 5)] >> (y & 31)) & 1u)) return true;
        return edge_q(x, y) > idx_e;
    }

    // ------------------------------------------------------------------ a run of equal-length edges
    __device__ void tie_run(const Params& p, int b, int r, int r1) {
        const float dcur = dist(P[r] >> 8, P[r] & 255);
        for (int e = tid; e < N * MW; e += NTH) runadj[e] = 0;
        __syncthreads();
        // pass 1: all edges of the run enter, H0 decisions in rank order.  Merging edges are
        // remembered through defv = 254.
        for (int pidx = r; pidx < r1; ++pidx) {
            const int pij = P[pidx];
            const int i = pij >> 8, j = pij & 255;
            const bool merging = h0_step(p, b, i, j);
            add_adj(i, j, true);
            if (tid == 0) defv[c2(i) + j] = merging ? 254 : 255;
            if (tid < W) phi[(size_t)(c2(i) + j) * W + tid] = 0;
            __syncthreads();
        }
        // pass 2 (parallel over the run's edges): apparent pairs inside the run
        for (int pidx = r + tid; pidx < r1; pidx += NTH) {
            const int pij = P[pidx];
            const int i = pij >> 8, j = pij & 255;
            const int q = c2(i) + j;
            if (defv[q] == 254) continue;
            int vt = -1;
            for (int w = MW - 1; w >= 0 && vt < 0; --w) {
                uint32_t g = adj[i * MW + w] & adj[j * MW + w];
                if (g) vt = 32 * w + 31 - __clz(g);
            }
            if (vt >= 0 && earlier(i, vt, q) && earlier(j, vt, q)) defv[q] = (uint8_t)vt;  // vt <= 253 (kMaxN)
        }
        __syncthreads();
        // slots for the cycle-creating run edges that are not apparent, in rank order
        for (int pidx = r; pidx < r1; ++pidx) {
            const int pij = P[pidx];
            const int q = c2(pij >> 8) + (pij & 255);
            if (defv[q] != 255) continue;
            const int s = alloc_slot();
            if (s < 0) return;
            if (tid == 0) brank[s] = (uint16_t)pidx;
            if (tid < W) phi[(size_t)q * W + tid] = (tid == (s >> 5)) ? (1u << (s & 31)) : 0u;
        }
        __syncthreads();
        if (live_any()) {
            // pass 3: the run's triangles in descending index: a desc, b desc, apex c desc (c < b < a)
            for (int a = N - 1; a >= 2 && !overflow; --a) {
                // candidate b's: neighbours of a below a
                for (int bw = (a - 1) >> 5; bw >= 0 && !overflow; --bw) {
                    uint32_t bbits = adj[a * MW + bw];
                    if (bw == (a >> 5)) bbits &= (1u << (a & 31)) - 1u;
                    while (bbits && !overflow) {
                        const int bb = 32 * bw + 31 - __clz(bbits);
                        bbits &= ~(1u << (bb & 31));
                        const bool ab_in_run = (runadj[a * MW + bw] >> (bb & 31)) & 1u;
                        uint32_t G[MW];
                        uint32_t anyG = 0;
#pragma unroll
                        for (int w = 0; w < MW; ++w) {
                            uint32_t g = adj[a * MW + w] & adj[bb * MW + w];
                            if (w > (bb >> 5)) g = 0;
                            else if (w == (bb >> 5)) g &= (1u << (bb & 31)) - 1u;
                            if (!ab_in_run) g &= (runadj[a * MW + w] | runadj[bb * MW + w]);
                            G[w] = g;
                            anyG |= g;
                        }
                        if (!anyG) continue;
                        const int dv = defv[c2(a) + bb];
                        const int vdef = (dv < 254 && dv < bb && ab_in_run) ? dv : -1;
                        group_eval(a, bb, G, vdef, true, dcur);
                        if (!live_any()) goto done;
                    }
                }
            }
        }
    done:
        __syncthreads();
        for (int pidx = r + tid; pidx < r1; pidx += NTH) {
            const int pij = P[pidx];
            defv[c2(pij >> 8) + (pij & 255)] = 255;
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ CTA-wide radix sort of (K, P)
    __device__ void sort_edges() {
        uint32_t* srcK = K; uint16_t* srcP = P;
        uint32_t* dstK = K2; uint16_t* dstP = P2;
        const uint32_t lt = lanemask_lt();
        const int seg = Epad / MW;
        const int k_begin = warp * seg, k_end = k_begin + seg;
        uint32_t* myh = hist + warp * 256;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 8 * pass;
#pragma unroll
            for (int t = 0; t < 8; ++t) myh[lane + 32 * t] = 0;
            __syncwarp();
            // digit histogram of this warp's segment: shared-memory atomics (independent iterations)
            for (int k0 = k_begin; k0 < k_end; k0 += 32) atomicAdd(myh + ((srcK[k0 + lane] >> shift) & 255u), 1u);
            __syncthreads();
            if (warp == 0) {
                // per digit: exclusive prefix over warps, then exclusive scan over digits
                uint32_t tot[8], sum = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int dgt = lane * 8 + t;
                    uint32_t s = 0;
                    for (int w = 0; w < MW; ++w) { uint32_t x = hist[w * 256 + dgt]; hist[w * 256 + dgt] = s; s += x; }
                    tot[t] = s;
                    sum += s;
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t y = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += y;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int dgt = lane * 8 + t;
                    for (int w = 0; w < MW; ++w) hist[w * 256 + dgt] += run;
                    run += tot[t];
                }
            }
            __syncthreads();
            for (int k0 = k_begin; k0 < k_end; k0 += 32) {
                uint32_t key = srcK[k0 + lane];
                uint16_t pay = srcP[k0 + lane];
                uint32_t dg = (key >> shift) & 255u;
                uint32_t peers = __match_any_sync(kFull, dg);
                uint32_t pos = myh[dg] + __popc(peers & lt);
                __syncwarp();
                dstK[pos] = key;
                dstP[pos] = pay;
                if ((peers & lt) == 0) myh[dg] += __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            uint32_t* tk = srcK; srcK = dstK; dstK = tk;
            uint16_t* tp = srcP; srcP = dstP; dstP = tp;
        }
    }

    // ------------------------------------------------------------------ one cloud
    __device__ void run(const Params& p, int b) {
        N = p.npts ? p.npts[b] : p.N;
        if (N > p.N) N = p.N;
        if (N < 0) N = 0;
        E = c2(N);
        Epad = epad_of<MW>(N);
        Db = p.D + (size_t)b * p.strideB;
        overflow = false;
        n0 = n1 = 0;
        ncomp = N;
#pragma unroll
        for (int w = 0; w < W; ++w) live[w] = used[w] = 0;
        if (N < 2) {
            if (tid == 0) {
                if (N == 1 && p.cap0 > 0) {
                    size_t o = (size_t)b * p.cap0 * 2;
                    p.bd0[o] = 0.0f; p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = 0; p.pr0[o + 1] = -1; }
                }
                p.counts[2 * b] = N; p.counts[2 * b + 1] = 0; p.status[b] = 0;
            }
            return;
        }
        // ---- keys in descending edge-index order
        int valid = 0, nan_seen = 0;
        for (int k = tid; k < Epad; k += NTH) {
            uint32_t key = 0xFFFFFFFFu;
            uint16_t pay = 0;
            if (k < E) {
                const int e = E - 1 - k;
                int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);
                while (c2(i) > e) --i;
                while (c2(i + 1) <= e) ++i;
                const int j = e - c2(i);
                const float d = dist(i, j);
                const bool ok = d <= p.thresh;
                nan_seen |= (d != d);
                valid += ok;
                if (ok) key = float_key(d);
                pay = (uint16_t)((i << 8) | j);
            }
            K[k] = key;
            P[k] = pay;
        }
        for (int e = tid; e < N * MW; e += NTH) adj[e] = 0;
        for (int v = tid; v < N; v += NTH) { comp[v] = (uint8_t)v; eld[v] = (uint8_t)v; }
        __syncthreads();
        {
            // block reduction of valid / nan_seen
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                valid += __shfl_xor_sync(kFull, valid, o);
                nan_seen |= __shfl_xor_sync(kFull, nan_seen, o);
            }
            if (lane == 0) { wtop[warp] = valid; wtop[MW + warp] = nan_seen; }
            __syncthreads();
            valid = 0; nan_seen = 0;
#pragma unroll
            for (int w = 0; w < MW; ++w) { valid += wtop[w]; nan_seen |= wtop[MW + w]; }
            m = valid;
            __syncthreads();
        }
        sort_edges();
        // ---- tie bitmap (needs the sorted keys, which PHI is about to overwrite when W == 2)
        {
            uint32_t mybits[8];
            int cnt = 0;
            for (int k0 = 32 * (tid); k0 < Epad; k0 += 32 * NTH) {
                uint32_t bits = 0;
                for (int t = 0; t < 32; ++t) {
                    const int k = k0 + t;
                    if (k + 1 < m && K[k] == K[k + 1]) bits |= 1u << t;
                }
                if (cnt < 8) mybits[cnt] = bits;
                ++cnt;
            }
            __syncthreads();
            cnt = 0;
            for (int k0 = 32 * (tid); k0 < Epad; k0 += 32 * NTH) { tiebits[k0 >> 5] = mybits[cnt < 8 ? cnt : 7]; ++cnt; }
            for (int q = tid; q < E; q += NTH) defv[q] = 255;
            __syncthreads();
        }
        // ---- the sweep
        int r = 0;
        while (r < m && !overflow) {
            if (!tie_after(r)) { single_edge(p, b, r); ++r; }
            else {
                int r1 = r + 1;
                while (tie_after(r1)) ++r1;
                ++r1;
                tie_run(p, b, r, r1);
                r = r1;
            }
        }
        if (!overflow) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = live[w];
                while (bits) {
                    int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (n1 < R) {
                        if (tid == 0) {
                            rec[n1] = brank[32 * w + s];
                            rec[R + n1] = kEssential;
                            rec[2 * R + n1] = kEssential;
                        }
                        ++n1;
                    } else overflow = true;
                }
            }
        }
        __syncthreads();
        if (overflow) {
            if (p.overflow_list) {
                if (tid == 0) p.overflow_list[atomicAdd(p.n_overflow, 1)] = b;
            } else if (tid == 0) {
                p.status[b] = TDA_ST_INTERNAL;
                p.counts[2 * b] = 0;
                p.counts[2 * b + 1] = 0;
            }
            return;
        }
        // ---- H0 essentials (ascending eldest vertex)
        {
            const bool is = tid < N && eld[comp[tid]] == tid;
            const uint32_t bal = __ballot_sync(kFull, is);
            if (lane == 0) wtop[warp] = __popc(bal);
            __syncthreads();
            int base = n0, tot = 0;
#pragma unroll
            for (int w = 0; w < MW; ++w) { if (w < warp) base += wtop[w]; tot += wtop[w]; }
            if (is) {
                const int pos = base + __popc(bal & lanemask_lt());
                if (pos < p.cap0) {
                    size_t o = ((size_t)b * p.cap0 + pos) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = tid; p.pr0[o + 1] = -1; }
                }
            }
            n0 += tot;
            __syncthreads();
        }
        // ---- H1 rows, descending birth rank
        int st = nan_seen ? TDA_ST_NAN_INPUT : 0;
        for (int k = tid; k < n1; k += NTH) {
            const uint32_t br = rec[k];
            int pos = 0;
            for (int t = 0; t < n1; ++t) pos += rec[t] > br;
            if (pos < p.cap1) {
                size_t o = ((size_t)b * p.cap1 + pos) * 2;
                const uint32_t dk = rec[R + k], tr = rec[2 * R + k];
                const int pij = P[br];
                p.bd1[o] = dist(pij >> 8, pij & 255);
                p.bd1[o + 1] = (tr == kEssential) ? __int_as_float(0x7F800000) : __uint_as_float(dk);
                if (p.pr1) {
                    p.pr1[o] = c2(pij >> 8) + (pij & 255);
                    p.pr1[o + 1] = (tr == kEssential) ? -1ll : (long long)tr;
                }
            }
        }
        if (n1 > p.cap1) st |= TDA_ST_H1_TRUNCATED;
        if (tid == 0) {
            p.counts[2 * b] = n0;
            p.counts[2 * b + 1] = n1;
            p.status[b] = st;
        }
        __syncthreads();
    }
};

template <int MW, int W>
__global__ void __launch_bounds__(32 * MW) rips_medium_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cta<MW, W> s;
    s.tid = threadIdx.x;
    s.lane = threadIdx.x & 31;
    s.warp = threadIdx.x >> 5;
    s.ld = p.ld;
    s.R = recs_of<W>();
    unsigned char* base = smem_raw;
    s.adj = (uint32_t*)base;     base += (size_t)kRows * MW * 4;
    s.runadj = (uint32_t*)base;  base += (size_t)kRows * MW * 4;
    s.hist = (uint32_t*)base;    base += (size_t)MW * 256 * 4;
    s.rec = (uint32_t*)base;     base += (size_t)3 * s.R * 4;
    s.wtop = (int*)base;         base += 64;
    s.bc = (uint32_t*)base;      base += W * 4 + 64;
    s.brank = (uint16_t*)base;   base += 32 * W * 2;
    s.comp = (uint8_t*)base;     base += kRows;
    s.eld = (uint8_t*)base;      base += kRows;
    base = smem_raw + fixed_bytes<MW, W>();
    unsigned char* big = p.big ? p.big + (size_t)blockIdx.x * p.big_stride : base;
    const int EpMax = epad_of<MW>(p.N);
    s.K = (uint32_t*)big;
    s.K2 = (uint32_t*)(big + (size_t)EpMax * 4);
    s.P = (uint16_t*)(big + (size_t)EpMax * 8);
    s.P2 = (uint16_t*)(big + (size_t)EpMax * 10);
    s.tiebits = (uint32_t*)s.P2;
    s.defv = (uint8_t*)(big + (size_t)EpMax * 10 + (size_t)EpMax / 8 + 16);
    s.phi = (W == 2) ? s.K : p.phi_global + (size_t)blockIdx.x * p.phi_stride;
    const int total = p.worklist ? *p.n_work : p.B;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = p.worklist ? p.worklist[t] : t;
        s.run(p, b);
        __syncthreads();
    }
}

struct WsLayout {
    size_t counters, list1, list2, big[3], phi[2], total;
    int grid[3];
    bool big_global[3];
};

template <int MW> static WsLayout ws_layout(int B, int N, int sms) {
    WsLayout w;
    size_t o = 0;
    w.counters = o; o += 64;
    w.list1 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    w.list2 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    const size_t bb = (big_bytes<MW>(N) + 255) & ~(size_t)255;
    const size_t lim = 227 * 1024;
    // tier 1
    {
        const size_t fx = fixed_bytes<MW, 2>();
        w.big_global[0] = fx + bb > lim;
        int per_sm = w.big_global[0] ? 2 : (int)(lim / (fx + bb + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        long long g = (long long)sms * per_sm;
        w.grid[0] = (int)(g < B ? g : (B > 0 ? B : 1));
        w.big[0] = o;
        if (w.big_global[0]) o += bb * w.grid[0];
    }
    // tier 2 (W = 8) and tier 3 (W = 32): PHI always global
    const int Wt[2] = {8, 32};
    const int gt[2] = {sms, 64};
    for (int t = 0; t < 2; ++t) {
        const size_t fx = t == 0 ? fixed_bytes<MW, 8>() : fixed_bytes<MW, 32>();
        w.big_global[t + 1] = fx + bb > lim;
        w.grid[t + 1] = gt[t];
        w.big[t + 1] = o;
        if (w.big_global[t + 1]) o += bb * gt[t];
        w.phi[t] = o;
        o += ((size_t)c2(N) * Wt[t] * 4 + 255) / 256 * 256 * gt[t];
    }
    w.total = o;
    return w;
}

template <int MW, int W>
static cudaError_t launch(Params p, const WsLayout& wl, int tier, char* w8, cudaStream_t st, const char* name) {
    const size_t fx = fixed_bytes<MW, W>();
    const size_t bb = (big_bytes<MW>(p.N) + 255) & ~(size_t)255;
    size_t smem = fx;
    if (wl.big_global[tier]) { p.big = (unsigned char*)(w8 + wl.big[tier]); p.big_stride = bb; }
    else { p.big = nullptr; p.big_stride = 0; smem += bb; }
    if (W > 2) {
        p.phi_global = (uint32_t*)(w8 + wl.phi[tier - 1]);
        p.phi_stride = ((size_t)c2(p.N) * W * 4 + 255) / 256 * 256 / 4;
    }
    cudaError_t e = cudaFuncSetAttribute(rips_medium_kernel<MW, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ProfScope prof(name, st);
    rips_medium_kernel<MW, W><<<wl.grid[tier], 32 * MW, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <int MW>
static int run_all(Params p, int sms, void* ws, size_t ws_bytes, cudaStream_t st) {
    WsLayout wl = ws_layout<MW>(p.B, p.N, sms);
    if (ws_bytes < wl.total) return TDA_E_WORKSPACE;
    char* w8 = (char*)ws;
    int* counters = (int*)(w8 + wl.counters);
    cudaError_t e = cudaMemsetAsync(counters, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    p.worklist = nullptr; p.n_work = nullptr;
    p.overflow_list = (int*)(w8 + wl.list1); p.n_overflow = counters + 0;
    e = launch<MW, 2>(p, wl, 0, w8, st, "rips_medium_w2");
    if (e != cudaSuccess) return (int)e;
    p.worklist = (const int*)(w8 + wl.list1); p.n_work = counters + 0;
    p.overflow_list = (int*)(w8 + wl.list2); p.n_overflow = counters + 1;
    e = launch<MW, 8>(p, wl, 1, w8, st, "rips_medium_w8");
    if (e != cudaSuccess) return (int)e;
    p.worklist = (const int*)(w8 + wl.list2); p.n_work = counters + 1;
    p.overflow_list = nullptr; p.n_overflow = nullptr;
    e = launch<MW, 32>(p, wl, 2, w8, st, "rips_medium_w32");
    return (int)e;
}

constexpr int kSms = 148;  // B200; grids and the workspace layout are sized for it

}  // namespace rips_medium
}  // namespace tda

using namespace tda::rips_medium;

extern "C" size_t tda_rips_h01_medium_workspace_bytes(int B, int N) {
    if (B < 0 || N < 2 || N > kMaxN) return 0;
    return N <= 128 ? ws_layout<4>(B, N, kSms).total : ws_layout<8>(B, N, kSms).total;
}

extern "C" int tda_rips_h01_medium(const float* D, const int* npts, int B, int N, int ld, long long strideB,
                                   float thresh, float* bd0, long long* pr0, int cap0, float* bd1, long long* pr1,
                                   int cap1, int* counts, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (!D || !bd0 || !bd1 || !counts || !status || !ws || B < 0 || cap0 < 0 || cap1 < 0 || ld < N) return TDA_E_ARG;
    if (N < 2 || N > kMaxN) return TDA_E_SIZE;
    if (B == 0) return 0;
    Params p;
    p.D = D; p.npts = npts; p.strideB = strideB ? strideB : (long long)ld * ld; p.ld = ld; p.N = N; p.B = B;
    p.thresh = thresh;
    p.bd0 = bd0; p.pr0 = pr0; p.bd1 = bd1; p.pr1 = pr1; p.counts = counts; p.status = status;
    p.cap0 = cap0; p.cap1 = cap1;
    p.big = nullptr; p.big_stride = 0; p.phi_global = nullptr; p.phi_stride = 0;
    return N <= 128 ? run_all<4>(p, kSms, ws, ws_bytes, (cudaStream_t)stream)
                    : run_all<8>(p, kSms, ws, ws_bytes, (cudaStream_t)stream);
}
