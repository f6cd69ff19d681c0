// rips_small.cu — Vietoris–Rips H0+H1 (Z/2) for batches of small distance matrices (N <= 64).
//
// Replaces ripser.ripser(dm, maxdim=1, thresh, distance_matrix=True) as called per window by
//   /root/reference/scripts/utils.py:131,140
//   /root/reference/scripts/tda_eeg_classification_v2.py:170-175
//
// B200-first design (not a port of Ripser's heap-based column reduction):
//   * one WARP owns one window; a CTA is a bundle of independent warps, the grid is sized in
//     multiples of the SM count and every warp strides over the batch.  All state of a window
//     lives in that warp's slice of shared memory (~18 KB for N=47) — the distance matrix is
//     read from HBM exactly once, coalesced by rows, and the diagrams are written exactly once.
//   * filtration = per-warp stable LSD radix sort (4 x 8 bit, __match_any_sync ranking) of the
//     order-preserving integer image of the float32 edge lengths; initial order is descending
//     edge index so stability gives Ripser's tie-break (equal length => larger index first).
//   * H0 and H1 come out of ONE sweep over the sorted edges.  H0 is Kruskal with warp-parallel
//     relabelling.  H1 is persistent cohomology by cocycle annotation: every live 1-cocycle is a
//     bit ("slot") in a W-word mask stored per edge (PHI[edge]); adj[v] holds the neighbours of v
//     so far as a 64-bit mask.  For a cycle-creating edge (i,j) the apexes of the triangles that
//     enter with it are G = adj[i] & adj[j]; lanes take one apex each and evaluate the coboundary
//     of *all* live cocycles on that triangle with two XORs.  Apparent (zero-persistence) pairs
//     cost nothing; only the ~30 real classes of a window ever allocate a slot.  See
//     oracle/pcoh_model.py for the executable statement and its proof-by-test against the
//     definition-level reduction.
//   * tie runs (equal float32 lengths) are replayed in the exact simplexwise order (all edges of
//     the run, then the run's triangles in descending index) so that the persistence PAIRS, not
//     only the diagrams, are bit-identical to Ripser's.
//   * capacity tiers: W=2 (64 simultaneous classes, shared memory) -> W=4 -> W=64 with PHI in a
//     global scratch; a window that exceeds a tier is pushed on a device-side list and redone by
//     the next tier, no host round-trip.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace rips_small {

constexpr int kMaxN = 64;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kEssential = 0xFFFFFFFFu;

struct Params {
    const float* D;
    long long strideB;
    int ld, N, B;
    float thresh;
    float* bd0;
    long long* pr0;
    float* bd1;
    long long* pr1;
    int* counts;
    int* status;
    int cap1;
    const int* worklist;   // nullptr => every window 0..B-1
    const int* n_work;     // device counter with the length of worklist
    int* overflow_list;    // nullptr on the last tier
    int* n_overflow;
    uint32_t* phi_global;  // per-warp PHI scratch (PHI_GLOBAL tiers), E*W words per warp
    uint32_t* rec_global;  // per-warp record scratch (PHI_GLOBAL tiers), 3*R words per warp
};

__host__ __device__ inline int c2(int i) { return i * (i - 1) / 2; }
__host__ __device__ inline int c3(int i) { return i * (i - 1) * (i - 2) / 6; }

template <int W, bool PHI_GLOBAL> struct Layout {
    // all sizes in bytes, per warp
    static __host__ __device__ int epad(int N) { return (c2(N) + 31) & ~31; }
    static __host__ __device__ int recs(int N) { return PHI_GLOBAL ? c2(N) + 64 : 64 * W; }
    static __host__ __device__ size_t region1(int N) {
        size_t sortb = (size_t)epad(N) * 6;
        size_t phib = PHI_GLOBAL ? 0 : (size_t)c2(N) * W * 4;
        size_t r = sortb > phib ? sortb : phib;
        return (r + 15) & ~(size_t)15;
    }
    static __host__ __device__ size_t bytes(int N) {
        size_t s = 0;
        s += 2 * kMaxN * 8;                                   // adj, runadj
        s += (size_t)epad(N) * 4;                             // K
        s += region1(N);                                      // sort ping-pong | PHI
        s += 256 * 4;                                         // radix histogram
        s += PHI_GLOBAL ? 0 : (size_t)recs(N) * 12;           // death records
        s += (size_t)epad(N) * 2;                             // P
        s += (size_t)32 * W * 2;                              // brank
        s += 2 * kMaxN;                                       // comp, eld
        return (s + 15) & ~(size_t)15;
    }
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t float_key(float d) {
    uint32_t u = __float_as_uint(d);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t u = (k >> 31) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

// combinatorial index of triangle {i, j, v}, i > j, v distinct
__device__ __forceinline__ int tri_index(int i, int j, int v) {
    if (v > i) return c3(v) + c2(i) + j;
    if (v > j) return c3(i) + c2(v) + j;
    return c3(i) + c2(j) + v;
}
__device__ __forceinline__ int edge_q(int a, int b) { return a > b ? c2(a) + b : c2(b) + a; }

template <int W, bool PHI_GLOBAL> struct Warp {
    // ---- per-warp storage
    unsigned long long* adj;
    unsigned long long* runadj;
    uint32_t* K;
    uint32_t* R1;  // region1: sort ping-pong, then PHI (shared tiers)
    uint32_t* hist;
    uint32_t* rec;  // [3][R]: birth rank, death key, death triangle
    uint16_t* P;
    uint16_t* brank;
    uint8_t* comp;
    uint8_t* eld;
    uint32_t* phi;
    int lane, N, E, Epad, R;
    // ---- per-window uniform state
    uint32_t live[W], used[W];
    int n0, n1, ncomp, m;
    bool overflow;

    __device__ __forceinline__ bool live_any() const {
        uint32_t a = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) a |= live[w];
        return a != 0;
    }

    // ------------------------------------------------------------------ slots
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
            // every slot has been used once: scrub the dead bits out of PHI and recycle
            __syncwarp();
            for (int q = lane; q < E; q += 32) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi[(size_t)q * W + w] &= live[w];
            }
            __syncwarp();
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

    __device__ __forceinline__ void phi_store_uniform(int q, const uint32_t (&v)[W]) {
        if (lane == 0) {
#pragma unroll
            for (int w = 0; w < W; ++w) phi[(size_t)q * W + w] = v[w];
        }
    }

    // ------------------------------------------------------------------ H0 step
    // returns true when edge (i,j) merges two components (emits the H0 pair)
    __device__ bool h0_step(const Params& p, int b, int i, int j, uint32_t key) {
        if (ncomp <= 1) return false;
        int ci = comp[i], cj = comp[j];
        if (ci == cj) return false;
        int ei = eld[ci], ej = eld[cj];
        float d = key_float(key);
        if (d != 0.0f) {
            if (lane == 0) {
                size_t o = ((size_t)b * N + n0) * 2;
                p.bd0[o] = 0.0f;
                p.bd0[o + 1] = d;
                if (p.pr0) {
                    p.pr0[o] = min(ei, ej);
                    p.pr0[o + 1] = c2(i) + j;
                }
            }
            ++n0;
        }
        __syncwarp();
        for (int v = lane; v < N; v += 32)
            if (comp[v] == ci) comp[v] = (uint8_t)cj;
        if (lane == 0) eld[cj] = (uint8_t)max(ei, ej);
        __syncwarp();
        --ncomp;
        return true;
    }

    __device__ __forceinline__ void add_adj(int i, int j, bool run) {
        if (lane == 0) {
            adj[i] |= 1ull << j;
            adj[j] |= 1ull << i;
            if (run) {
                runadj[i] |= 1ull << j;
                runadj[j] |= 1ull << i;
            }
        }
    }

    // ------------------------------------------------------------------ deaths inside a group
    // c[h] = coboundary masks of the live cocycles on triangle (i, j, v = lane + 32 h)
    __device__ void resolve(int i, int j, uint32_t key, uint32_t (&c)[2][W]) {
        while (true) {
            uint32_t any0 = 0, any1 = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) { any0 |= c[0][w]; any1 |= c[1][w]; }
            uint32_t nz1 = __ballot_sync(kFull, any1 != 0);
            uint32_t nz0 = __ballot_sync(kFull, any0 != 0);
            if (!(nz0 | nz1)) return;
            int h = nz1 ? 1 : 0;
            int src = 31 - __clz(nz1 ? nz1 : nz0);
            int v = src + 32 * h;
            uint32_t cv[W];
#pragma unroll
            for (int w = 0; w < W; ++w) cv[w] = __shfl_sync(kFull, h ? c[1][w] : c[0][w], src);
            // youngest live class with coefficient 1 dies
            int slot = -1, age = -1;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = cv[w];
                while (bits) {
                    int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    int a = brank[32 * w + s];
                    if (a > age) { age = a; slot = 32 * w + s; }
                }
            }
            const int sw = slot >> 5;
            const uint32_t sb = 1u << (slot & 31);
            if (K[age] != key) {  // non-zero persistence: keep a record
                if (n1 < R) {
                    if (lane == 0) {
                        rec[n1] = (uint32_t)age;
                        rec[R + n1] = key;
                        rec[2 * R + n1] = (uint32_t)tri_index(i, j, v);
                    }
                    ++n1;
                } else {
                    overflow = true;
                    return;
                }
            }
            bool absorb = false;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (w == sw) { live[w] &= ~sb; absorb |= (cv[w] & ~sb) != 0; }
                else absorb |= cv[w] != 0;
            }
            // the other classes with coefficient 1 absorb the dying cocycle (linear update)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                bool has = false;
#pragma unroll
                for (int w = 0; w < W; ++w) if (w == sw) has = (c[hh][w] & sb) != 0;
                if (has) {
#pragma unroll
                    for (int w = 0; w < W; ++w) c[hh][w] ^= cv[w];
                }
            }
            if (absorb) {
                __syncwarp();
                for (int q = lane; q < E; q += 32) {
                    uint32_t* e = phi + (size_t)q * W;
                    if (e[sw] & sb) {
#pragma unroll
                        for (int w = 0; w < W; ++w) e[w] ^= cv[w];
                    }
                }
                __syncwarp();
            }
        }
    }

    // evaluate the triangles (i, j, v), v in G, against PHI as it stands and resolve deaths
    __device__ void group_eval(int i, int j, unsigned long long G, uint32_t key) {
        uint32_t c[2][W];
        const int q = c2(i) + j;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int v = lane + 32 * h;
            bool in = (G >> v) & 1ull;
#pragma unroll
            for (int w = 0; w < W; ++w) c[h][w] = 0;
            if (in) {
                const uint32_t* pe = phi + (size_t)q * W;
                const uint32_t* pa = phi + (size_t)edge_q(i, v) * W;
                const uint32_t* pb = phi + (size_t)edge_q(j, v) * W;
#pragma unroll
                for (int w = 0; w < W; ++w) c[h][w] = (pe[w] ^ pa[w] ^ pb[w]) & live[w];
            }
        }
        resolve(i, j, key, c);
    }

    // ------------------------------------------------------------------ one single (untied) edge
    __device__ void fast_edge(const Params& p, int b, int r, uint32_t key) {
        const int pij = P[r];
        const int i = pij >> 8, j = pij & 255;
        const int q = c2(i) + j;
        const bool merging = h0_step(p, b, i, j, key);
        const unsigned long long G = adj[i] & adj[j];
        __syncwarp();
        add_adj(i, j, false);
        uint32_t zero[W];
#pragma unroll
        for (int w = 0; w < W; ++w) zero[w] = 0;
        if (merging) { phi_store_uniform(q, zero); __syncwarp(); return; }
        if (G == 0) {  // a real class is born
            int s = alloc_slot();
            if (s < 0) return;
            if (lane == 0) brank[s] = (uint16_t)r;
#pragma unroll
            for (int w = 0; w < W; ++w) zero[w] = (w == (s >> 5)) ? (1u << (s & 31)) : 0u;
            phi_store_uniform(q, zero);
            __syncwarp();
            return;
        }
        if (!live_any()) { phi_store_uniform(q, zero); __syncwarp(); return; }
        // apparent pair (e, top triangle); extend every live cocycle over e and test the others
        const int vtop = 63 - __clzll(G);
        uint32_t c[2][W];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int v = lane + 32 * h;
            bool in = (G >> v) & 1ull;
#pragma unroll
            for (int w = 0; w < W; ++w) c[h][w] = 0;
            if (in) {
                const uint32_t* pa = phi + (size_t)edge_q(i, v) * W;
                const uint32_t* pb = phi + (size_t)edge_q(j, v) * W;
#pragma unroll
                for (int w = 0; w < W; ++w) c[h][w] = pa[w] ^ pb[w];
            }
        }
        uint32_t xtop[W];
#pragma unroll
        for (int w = 0; w < W; ++w)
            xtop[w] = __shfl_sync(kFull, (vtop >> 5) ? c[1][w] : c[0][w], vtop & 31) & live[w];
        phi_store_uniform(q, xtop);
        uint32_t anyc = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int v = lane + 32 * h;
            bool in = ((G >> v) & 1ull) && v != vtop;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                c[h][w] = in ? ((c[h][w] ^ xtop[w]) & live[w]) : 0u;
                anyc |= c[h][w];
            }
        }
        __syncwarp();
        if (__ballot_sync(kFull, anyc != 0)) resolve(i, j, key, c);
    }

    // ------------------------------------------------------------------ a run of equal-length edges
    __device__ void tie_run(const Params& p, int b, int r, int r1, uint32_t key) {
        for (int v = lane; v < N; v += 32) runadj[v] = 0;
        __syncwarp();
        for (int pidx = r; pidx < r1; ++pidx) {
            const int pij = P[pidx];
            const int i = pij >> 8, j = pij & 255;
            const int q = c2(i) + j;
            const bool merging = h0_step(p, b, i, j, key);
            uint32_t val[W];
#pragma unroll
            for (int w = 0; w < W; ++w) val[w] = 0;
            if (!merging) {
                int s = alloc_slot();
                if (s < 0) return;
                if (lane == 0) brank[s] = (uint16_t)pidx;
#pragma unroll
                for (int w = 0; w < W; ++w) val[w] = (w == (s >> 5)) ? (1u << (s & 31)) : 0u;
            }
            phi_store_uniform(q, val);
            add_adj(i, j, true);
            __syncwarp();
        }
        if (!live_any()) return;
        // the run's triangles in descending index order: (a desc, b desc, c desc), a > b > c
        for (int a = N - 1; a >= 2; --a) {
            const unsigned long long adj_a = adj[a], run_a = runadj[a];
            const unsigned long long Pa = adj_a & ((1ull << a) - 1ull);
            if (Pa == 0) continue;
            unsigned long long mk[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int bb = lane + 32 * h;
                mk[h] = 0;
                if (bb < a && ((Pa >> bb) & 1ull)) {
                    unsigned long long cand = adj_a & adj[bb] & ((1ull << bb) - 1ull);
                    if (!((run_a >> bb) & 1ull)) cand &= (run_a | runadj[bb]);
                    mk[h] = cand;
                }
            }
            for (int h = 1; h >= 0; --h) {
                uint32_t bits = __ballot_sync(kFull, mk[h] != 0);
                while (bits) {
                    int src = 31 - __clz(bits);
                    bits &= ~(1u << src);
                    unsigned long long Gm = __shfl_sync(kFull, mk[h], src);
                    group_eval(a, src + 32 * h, Gm, key);
                    if (overflow) return;
                    if (!live_any()) return;
                }
            }
        }
    }

    // ------------------------------------------------------------------ radix sort of (K, P)
    __device__ void sort_edges() {
        uint32_t* K2 = R1;
        uint16_t* P2 = (uint16_t*)(R1 + Epad);
        uint32_t* srcK = K; uint16_t* srcP = P;
        uint32_t* dstK = K2; uint16_t* dstP = P2;
        const uint32_t lt = lanemask_lt();
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 8 * pass;
#pragma unroll
            for (int t = 0; t < 8; ++t) hist[lane + 32 * t] = 0;
            __syncwarp();
            for (int k0 = 0; k0 < Epad; k0 += 32) {
                uint32_t dg = (srcK[k0 + lane] >> shift) & 255u;
                uint32_t peers = __match_any_sync(kFull, dg);
                if ((peers & lt) == 0) hist[dg] += __popc(peers);
                __syncwarp();
            }
            // exclusive scan of the 256 bins (8 per lane)
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { loc[t] = hist[lane * 8 + t]; sum += loc[t]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
            }
            uint32_t run = incl - sum;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) { hist[lane * 8 + t] = run; run += loc[t]; }
            __syncwarp();
            for (int k0 = 0; k0 < Epad; k0 += 32) {
                uint32_t key = srcK[k0 + lane];
                uint16_t pay = srcP[k0 + lane];
                uint32_t dg = (key >> shift) & 255u;
                uint32_t peers = __match_any_sync(kFull, dg);
                uint32_t pos = hist[dg] + __popc(peers & lt);
                __syncwarp();
                dstK[pos] = key;
                dstP[pos] = pay;
                if ((peers & lt) == 0) hist[dg] += __popc(peers);
                __syncwarp();
            }
            uint32_t* tk = srcK; srcK = dstK; dstK = tk;
            uint16_t* tp = srcP; srcP = dstP; dstP = tp;
        }
        // 4 passes: result is back in (K, P)
    }

    // ------------------------------------------------------------------ one window
    __device__ void run(const Params& p, int b) {
        const float* Db = p.D + (size_t)b * p.strideB;
        overflow = false;
        n0 = n1 = 0;
        ncomp = N;
#pragma unroll
        for (int w = 0; w < W; ++w) live[w] = used[w] = 0;
        // ---- keys, initial order = descending edge index
        int valid = 0, nan_seen = 0;
        for (int k = E + lane; k < Epad; k += 32) { K[k] = 0xFFFFFFFFu; P[k] = 0; }
        for (int row = 0; row < N - 1; ++row) {
            for (int i = row + 1 + lane; i < N; i += 32) {
                float d = __ldg(Db + (size_t)row * p.ld + i) + 0.0f;
                bool ok = d <= p.thresh;
                nan_seen |= (d != d);
                int k = E - 1 - (c2(i) + row);
                K[k] = ok ? float_key(d) : 0xFFFFFFFFu;
                P[k] = (uint16_t)((i << 8) | row);
                valid += ok;
            }
        }
        for (int v = lane; v < N; v += 32) {
            adj[v] = 0;
            comp[v] = (uint8_t)v;
            eld[v] = (uint8_t)v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            valid += __shfl_xor_sync(kFull, valid, o);
            nan_seen |= __shfl_xor_sync(kFull, nan_seen, o);
        }
        m = valid;
        __syncwarp();
        sort_edges();
        __syncwarp();
        // ---- the sweep
        int r = 0;
        while (r < m && !overflow) {
            const uint32_t key = K[r];
            int r1 = r + 1;
            while (r1 < m && K[r1] == key) ++r1;
            if (r1 - r == 1) fast_edge(p, b, r, key);
            else tie_run(p, b, r, r1, key);
            r = r1;
        }
        if (!overflow) {
            // cycles still alive at thresh are essential
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = live[w];
                while (bits) {
                    int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (n1 < R) {
                        if (lane == 0) {
                            rec[n1] = brank[32 * w + s];
                            rec[R + n1] = kEssential;
                            rec[2 * R + n1] = kEssential;
                        }
                        ++n1;
                    } else overflow = true;
                }
            }
        }
        if (overflow) {
            if (p.overflow_list) {
                if (lane == 0) p.overflow_list[atomicAdd(p.n_overflow, 1)] = b;
            } else if (lane == 0) {
                p.status[b] = TDA_ST_INTERNAL;
                p.counts[2 * b] = 0;
                p.counts[2 * b + 1] = 0;
            }
            return;
        }
        __syncwarp();
        // ---- H0 essentials: eldest vertex of every surviving component, ascending
        {
            int base = n0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int v = lane + 32 * h;
                bool is = v < N && eld[comp[v]] == v;
                uint32_t bal = __ballot_sync(kFull, is);
                if (is) {
                    size_t o = ((size_t)b * N + base + __popc(bal & lanemask_lt())) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = v; p.pr0[o + 1] = -1; }
                }
                base += __popc(bal);
            }
            n0 = base;
        }
        // ---- H1 rows in ripser's order: descending birth rank
        int st = nan_seen ? TDA_ST_NAN_INPUT : 0;
        for (int k = lane; k < n1; k += 32) {
            uint32_t br = rec[k];
            int pos = 0;
            for (int t = 0; t < n1; ++t) pos += rec[t] > br;
            if (pos < p.cap1) {
                size_t o = ((size_t)b * p.cap1 + pos) * 2;
                uint32_t dk = rec[R + k], tr = rec[2 * R + k];
                p.bd1[o] = key_float(K[br]);
                p.bd1[o + 1] = (tr == kEssential) ? __int_as_float(0x7F800000) : key_float(dk);
                if (p.pr1) {
                    int pij = P[br];
                    p.pr1[o] = c2(pij >> 8) + (pij & 255);
                    p.pr1[o + 1] = (tr == kEssential) ? -1ll : (long long)tr;
                }
            }
        }
        if (n1 > p.cap1) st |= TDA_ST_H1_TRUNCATED;
        if (lane == 0) {
            p.counts[2 * b] = n0;
            p.counts[2 * b + 1] = n1;
            p.status[b] = st;
        }
        __syncwarp();
    }
};

template <int W, bool PHI_GLOBAL>
__global__ void __launch_bounds__(256) rips_small_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef Layout<W, PHI_GLOBAL> L;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    const int N = p.N;
    unsigned char* base = smem_raw + (size_t)wib * L::bytes(N);
    Warp<W, PHI_GLOBAL> s;
    s.lane = lane;
    s.N = N;
    s.E = c2(N);
    s.Epad = L::epad(N);
    s.R = L::recs(N);
    s.adj = (unsigned long long*)base;           base += kMaxN * 8;
    s.runadj = (unsigned long long*)base;        base += kMaxN * 8;
    s.K = (uint32_t*)base;                       base += (size_t)s.Epad * 4;
    s.R1 = (uint32_t*)base;                      base += L::region1(N);
    s.hist = (uint32_t*)base;                    base += 256 * 4;
    if (PHI_GLOBAL) {
        s.rec = p.rec_global + (size_t)gw * 3 * s.R;
        s.phi = p.phi_global + (size_t)gw * s.E * W;
    } else {
        s.rec = (uint32_t*)base;                 base += (size_t)s.R * 12;
        s.phi = s.R1;
    }
    s.P = (uint16_t*)base;                       base += (size_t)s.Epad * 2;
    s.brank = (uint16_t*)base;                   base += 32 * W * 2;
    s.comp = (uint8_t*)base;                     base += kMaxN;
    s.eld = (uint8_t*)base;
    const int total = p.worklist ? *p.n_work : p.B;
    for (int t = gw; t < total; t += nw) {
        const int b = p.worklist ? p.worklist[t] : t;
        s.run(p, b);
        __syncwarp();
    }
}

// workspace layout: [0..15] int counters ; list1[B] ; list2[B] ; phi scratch ; rec scratch
constexpr int kLastW = 64;
constexpr int kLastGrid = 148;  // one single-warp CTA per SM on the last tier
struct WsLayout {
    size_t counters, list1, list2, phi, rec, total;
};
static WsLayout ws_layout(int B, int N) {
    WsLayout w;
    size_t o = 0;
    w.counters = o; o += 64;
    w.list1 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    w.list2 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    w.phi = o; o += (size_t)kLastGrid * c2(N) * kLastW * 4;
    w.rec = o; o += (size_t)kLastGrid * 3 * Layout<kLastW, true>::recs(N) * 4;
    w.total = o;
    return w;
}

template <int W, bool G>
static cudaError_t launch_tier(const Params& p, int warps_per_block, int grid, cudaStream_t st) {
    ProfScope prof(W == 2 ? "rips_small_w2" : (W == 4 ? "rips_small_w4" : "rips_small_w64"), st);
    size_t smem = Layout<W, G>::bytes(p.N) * warps_per_block;
    cudaError_t e = cudaFuncSetAttribute(rips_small_kernel<W, G>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rips_small_kernel<W, G><<<grid, warps_per_block * 32, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rips_small
}  // namespace tda

using namespace tda;
using namespace tda::rips_small;

extern "C" size_t tda_rips_h01_workspace_bytes(int B, int N) {
    if (B < 0 || N < 2 || N > kMaxN) return 0;
    return ws_layout(B, N).total;
}

extern "C" int tda_rips_h01_batched(const float* D, int B, int N, int ld, long long strideB, float thresh,
                                    float* bd0, long long* pr0, float* bd1, long long* pr1, int* counts,
                                    int cap1, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (!D || !bd0 || !bd1 || !counts || !status || !ws || B < 0 || cap1 < 0 || ld < N) return TDA_E_ARG;
    if (N < 2 || N > kMaxN) return TDA_E_SIZE;
    if (B == 0) return 0;
    WsLayout wl = ws_layout(B, N);
    if (ws_bytes < wl.total) return TDA_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* w8 = (char*)ws;
    int* counters = (int*)(w8 + wl.counters);
    cudaError_t e = cudaMemsetAsync(counters, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    Params p;
    p.D = D; p.strideB = strideB ? strideB : (long long)N * ld; p.ld = ld; p.N = N; p.B = B;
    p.thresh = thresh;
    p.bd0 = bd0; p.pr0 = pr0; p.bd1 = bd1; p.pr1 = pr1; p.counts = counts; p.status = status; p.cap1 = cap1;
    p.phi_global = nullptr; p.rec_global = nullptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // tier 1: W=2, shared-memory PHI, 4 warps per CTA, as many CTAs per SM as shared memory allows
    {
        p.worklist = nullptr; p.n_work = nullptr;
        p.overflow_list = (int*)(w8 + wl.list1); p.n_overflow = counters + 0;
        const int wpb = 4;
        size_t smem = Layout<2, false>::bytes(N) * wpb;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 16) per_sm = 16;
        long long need = ((long long)B + wpb - 1) / wpb;
        int grid = (int)((long long)sms * per_sm < need ? (long long)sms * per_sm : need);
        e = launch_tier<2, false>(p, wpb, grid, st);
        if (e != cudaSuccess) return (int)e;
    }
    // tier 2: W=4 on the windows tier 1 gave up on
    {
        p.worklist = (const int*)(w8 + wl.list1); p.n_work = counters + 0;
        p.overflow_list = (int*)(w8 + wl.list2); p.n_overflow = counters + 1;
        e = launch_tier<4, false>(p, 2, sms * 2, st);
        if (e != cudaSuccess) return (int)e;
    }
    // tier 3: W=64 with PHI in global scratch, handles every N<=64 input
    {
        p.worklist = (const int*)(w8 + wl.list2); p.n_work = counters + 1;
        p.overflow_list = nullptr; p.n_overflow = nullptr;
        p.phi_global = (uint32_t*)(w8 + wl.phi);
        p.rec_global = (uint32_t*)(w8 + wl.rec);
        e = launch_tier<kLastW, true>(p, 1, kLastGrid, st);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
