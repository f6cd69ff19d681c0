// rips_small.cu — Vietoris–Rips H0+H1 (Z/2) for batches of small distance matrices (N <= 64).
//
// Replaces ripser.ripser(dm, maxdim=1, thresh, distance_matrix=True) as called per window by
//   /root/reference/scripts/utils.py:131,140
//   /root/reference/scripts/tda_eeg_classification_v2.py:170-175
//
// B200-first design (not a port of Ripser's heap-based column reduction):
//   * one WARP owns one window; a CTA is a bundle of independent warps, the grid is sized in
//     multiples of the SM count and every warp strides over the batch.  All state of a window
//     lives in that warp's slice of shared memory (9.4 KB on the 47-point tier, 24 warps per SM) —
//     the distance matrix (dense, or the condensed upper triangle ripser's own C++ entry takes) is
//     read from HBM exactly once, coalesced by rows, and the diagrams are written exactly once.
//   * filtration = per-warp stable LSD radix sort (8-bit digits, __match_any_sync ranking, passes over
//     bytes shared by all keys skipped) of the order-preserving integer image of the float32 edge
//     lengths; initial order is descending edge index so stability gives Ripser's tie-break (equal
//     length => larger index first).  On the 47-point tier the keys of a pass travel through
//     registers, which is what makes the window fit 9.4 KB (see Layout).
//   * H0 = Kruskal over the sorted edges, 32 edges checked per step, warp-parallel relabelling.
//   * the sorted ranks are scattered into a rank matrix T[i][v] (u16).  "Apex v closes a triangle
//     over edge (i,j) of rank r" is then max(T[i][v], T[j][v]) < r: no adjacency to maintain, and
//     one packed-u16 min/max pass over T decides for EVERY edge in parallel whether it has an
//     apex at its own time — the edges that do not are the births of real H1 classes.
//   * H1 = persistent cohomology by cocycle annotation (oracle/pcoh_model.py is the executable
//     statement): every live 1-cocycle is a bit ("slot") of a W-word mask stored per edge,
//     PHI[rank].  The serial sweep only runs through the LIVE SPANS (from a birth until no class is
//     alive); lanes are apexes, so the coboundary of ALL live cocycles on a triangle is two XORs
//     per lane.  Apparent (zero-persistence) pairs cost nothing and take no slot.  Inside a live span
//     the sweep looks at 32 ranks at once: S[v] = OR of the cocycle masks over the edges at v, and an
//     edge whose end points carry no live bit is not visited at all.
//   * tie runs (equal float32 lengths) are replayed in the exact simplexwise order (all edges of
//     the run, then the run's triangles in descending index, apparent pairs recognised inside the
//     run) so that the persistence PAIRS, not only the diagrams, are bit-identical to Ripser's.
//   * capacity tiers: W=1 (32 simultaneous classes, PHI for the first 588 ranks; 47-point windows
//     only, where 99.93 % of the benchmark's EEG windows fit) -> W=2 (64 classes, shared memory) -> W=4 ->
//     W=64 with PHI in a global scratch; a window that exceeds a tier is pushed on a device-side
//     list and redone by the next tier, no host round-trip.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "rips_small.cuh"
#include "tda_b200.h"

namespace tda {
namespace rips_small {

// RSW > 0 ("register sort", RSW warps per CTA, two CTAs per SM): the keys of a radix pass travel
// through registers, so the sort needs one key buffer and a payload ping-pong instead of two of each,
// and the window gets a fixed budget of shared memory, 1 / (2 RSW) of an SM: [K | T] [P] [P2 + digit
// counters | PHI] [visit, brank, comp, eld].  What is left for PHI after the other arrays sets phicap():
// all 1,081 ranks at ten warps per CTA, 588 at twelve (a window with a class alive or
// born beyond that rank is redone by the next tier; no EEG-like window of the benchmark is).
template <int W, bool PHI_GLOBAL, int RSW = 0> struct Layout {
    static constexpr bool RS = RSW > 0;
    // all sizes in bytes, per warp
    static __host__ __device__ int epad(int N) { return (c2(N) + 31) & ~31; }
    static __host__ __device__ int ldt(int N) { return ((((N + 1) / 2) | 1) * 2); }  // u16 per T row, odd #words
    static __host__ __device__ int recs(int N) { return PHI_GLOBAL ? c2(N) + 64 : (W < 2 ? 96 : 48 * W); }
    static __host__ __device__ size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }
    static __host__ __device__ size_t region_a(int N) {      // sort keys (+ histogram) | rank matrix T
        size_t s1 = (size_t)epad(N) * 4 + (RS ? 0 : 512), s2 = (size_t)N * ldt(N) * 2;   // 256 u16 digit counters
        return a16(s1 > s2 ? s1 : s2);
    }
    // death records of the one-word tier live in a global scratch (written ~40 times per window): the
    // kilobyte this frees lets eight warps share a CTA, sixteen an SM
    static constexpr bool kRecGlobal = PHI_GLOBAL || W == 1;
    // visit bitmap, brank, then comp + eld (Kruskal) overlaid by S, the per-vertex summary of the sweep
    static __host__ __device__ size_t tail(int N) {
        const size_t sv = (size_t)(RS ? N : kMaxN) * 4 * W, ce = 2 * kMaxN;
        return a16((size_t)epad(N) / 8 + 32 * W * 2 + (sv > ce ? sv : ce));
    }
    // an SM has 228 KB of shared memory; every resident CTA also takes 1 KB of it
    static __host__ __device__ size_t budget() { return (size_t)(((228 * 1024) / 2 - 1024) / (RS ? RSW : 1)) & ~(size_t)15; }
    // sort ping-pong | PHI (phicap() ranks; all of them unless the region is made smaller)
    static __host__ __device__ size_t region_c(int N) {
        if (RS) return budget() - region_a(N) - a16((size_t)epad(N) * 2) - tail(N);
        size_t s1 = (size_t)epad(N) * 6, s2 = PHI_GLOBAL ? 0 : (size_t)c2(N) * W * 4;
        return a16(s1 > s2 ? s1 : s2);
    }
    static __host__ __device__ int phicap(int N) {
        if (PHI_GLOBAL) return c2(N);
        const int cap = (int)(region_c(N) / (4 * W));
        return cap < c2(N) ? cap : c2(N);
    }
    static __host__ __device__ size_t off_p(int N) { return region_a(N) + (RS ? 0 : region_c(N)); }
    static __host__ __device__ size_t off_c(int N) { return region_a(N) + (RS ? a16((size_t)epad(N) * 2) : 0); }
    static __host__ __device__ size_t off_hist(int N) { return RS ? off_c(N) + (size_t)epad(N) * 2 : (size_t)epad(N) * 4; }
    static __host__ __device__ size_t off_p2(int N) { return off_c(N) + (RS ? 0 : (size_t)epad(N) * 4); }
    static __host__ __device__ size_t off_rec(int N) { return off_p(N) + a16((size_t)epad(N) * 2); }
    static __host__ __device__ size_t off_visit(int N) {
        if (RS) return budget() - tail(N);
        return off_rec(N) + (kRecGlobal ? 0 : (size_t)recs(N) * 12);
    }
    static __host__ __device__ size_t bytes(int N) {
        if (RS) return budget();
        size_t s = region_a(N) + region_c(N);
        s += a16((size_t)epad(N) * 2);                        // P
        s += kRecGlobal ? 0 : (size_t)recs(N) * 12;           // death records
        s += (size_t)epad(N) / 8;                             // visit bitmap
        s += 32 * W * 2;                                      // brank
        s += (size_t)kMaxN * 4 * W > 2 * kMaxN ? (size_t)kMaxN * 4 * W : 2 * kMaxN;   // comp, eld | S
        return a16(s);
    }
};

// RSW (one-word 47-point tier only): see Layout
template <int W, bool PHI_GLOBAL, int NT, int RSW = 0> struct Warp {
    static constexpr bool kRS = RSW > 0;
    static_assert(RSW == 0 || (NT > 0 && !PHI_GLOBAL && W == 1), "RSW variants are specialisations with a compile-time N");
    typedef Layout<W, PHI_GLOBAL, RSW> L;
    // ---- per-warp storage: everything is an offset from `base` (compile-time when NT > 0)
    unsigned char* base;
    uint32_t* phi_g;   // PHI_GLOBAL tiers
    uint32_t* rec_g;
    uint8_t* defv_g;   // tie runs only (global scratch): defining apex of an apparent run edge
    const float* Db;
    int lane, Nrt, ld;
    __device__ __forceinline__ int n() const { return NT > 0 ? NT : Nrt; }
    __device__ __forceinline__ int e() const { return c2(n()); }
    __device__ __forceinline__ int epad() const { return L::epad(n()); }
    __device__ __forceinline__ int rcap() const { return L::recs(n()); }
    __device__ __forceinline__ int ldtv() const { return L::ldt(n()); }
    __device__ __forceinline__ int phicap() const { return L::phicap(n()); }
    // region A: sort keys + histogram, later the rank matrix T (row stride ldtv())
    __device__ __forceinline__ uint32_t* K() const { return (uint32_t*)base; }
    __device__ __forceinline__ uint16_t* hist() const { return (uint16_t*)(base + L::off_hist(n())); }
    __device__ __forceinline__ uint16_t* T() const { return (uint16_t*)base; }
    // region C: sort ping-pong, later PHI[rank][W]
    __device__ __forceinline__ uint32_t* K2() const { return (uint32_t*)(base + L::off_c(n())); }
    __device__ __forceinline__ uint16_t* P2() const { return (uint16_t*)(base + L::off_p2(n())); }
    __device__ __forceinline__ uint32_t* phi() const { return PHI_GLOBAL ? phi_g : (uint32_t*)(base + L::off_c(n())); }
    // P[rank] = j | i << 6 | flags
    __device__ __forceinline__ uint16_t* P() const { return (uint16_t*)(base + L::off_p(n())); }
    // death records [3][R]: birth rank, death rank, death triangle
    __device__ __forceinline__ uint32_t* rec() const { return L::kRecGlobal ? rec_g : (uint32_t*)(base + L::off_rec(n())); }
    __device__ __forceinline__ uint32_t* visit() const { return (uint32_t*)(base + L::off_visit(n())); }
    __device__ __forceinline__ uint16_t* brank() const { return (uint16_t*)(base + L::off_visit(n()) + (size_t)epad() / 8); }
    __device__ __forceinline__ uint8_t* comp() const { return base + L::off_visit(n()) + (size_t)epad() / 8 + 32 * W * 2; }
    __device__ __forceinline__ uint8_t* eld() const { return comp() + kMaxN; }
    // S[v][W]: superset of the OR of PHI over the edges at vertex v (over comp / eld once Kruskal is done)
    __device__ __forceinline__ uint32_t* S() const { return (uint32_t*)comp(); }
    __device__ __forceinline__ uint8_t* defv() const { return defv_g; }
    // ---- per-window uniform state
    uint32_t live[W], used[W];
    int n0, n1, ncomp, m;
    bool overflow;
    bool s_changed;   // S gained bits since the sweep last evaluated which edges a live cocycle can see

    __device__ __forceinline__ float dist(int a, int b) const {
        return __ldg(Db + (d_rowoff(min(a, b), ld, n()) + max(a, b))) + 0.0f;
    }
    __device__ __forceinline__ bool live_any() const {
        uint32_t a = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) a |= live[w];
        return a != 0;
    }

    // ------------------------------------------------------------------ slots
    __device__ __forceinline__ int alloc_slot(int upto) {
        for (int attempt = 0; attempt < 2; ++attempt) {
            // (no indexed writes inside branches: a dynamic index would push live/used, and with
            // them the whole per-warp state, into local memory)
            int slot = -1;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t f = ~used[w];
                if (slot < 0 && f) slot = 32 * w + __ffs(f) - 1;
            }
            if (slot >= 0) {
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const uint32_t bit = (w == (slot >> 5)) ? (1u << (slot & 31)) : 0u;
                    used[w] |= bit;
                    live[w] |= bit;
                }
                return slot;
            }
            // every slot has been used once: scrub the dead bits out of PHI and recycle
            __syncwarp();
            for (int q = lane; q < upto; q += 32) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi()[(size_t)q * W + w] &= live[w];
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
    __device__ __forceinline__ void birth(int r, int upto) {   // r < phicap(): see the sweep and tie_run
        const int s = alloc_slot(upto);
        if (s < 0) return;
        if (lane == 0) {
            brank()[s] = (uint16_t)r;
            const uint32_t pij = P()[r];
            uint32_t* si = S() + p_i(pij) * W;
            uint32_t* sj = S() + p_j(pij) * W;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint32_t bit = (w == (s >> 5)) ? (1u << (s & 31)) : 0u;
                phi()[(size_t)r * W + w] = bit;
                si[w] |= bit;
                sj[w] |= bit;
            }
        }
        s_changed = true;
        __syncwarp();
    }

    // ------------------------------------------------------------------ deaths inside a group
    // c[h] = coboundary masks of the live cocycles on triangle (a, b, v = lane + 32 h); lanes with
    // isdef[h] carry the value of an apparent run edge they define (updated linearly, never a death)
    // rcur = rank of the edge whose length is the death value; zero0 = first rank with that length
    // (classes born at rank >= zero0 die with zero persistence and leave no record)
    __device__ __forceinline__ void resolve(int a, int b, int rcur, int zero0, int upto, uint32_t (&c)[2][W],
                                            const bool (&isdef)[2]) {
        while (true) {
            uint32_t any0 = 0, any1 = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) { any0 |= c[0][w]; any1 |= c[1][w]; }
            const uint32_t nz1 = __ballot_sync(kFull, any1 != 0 && !isdef[1]);
            const uint32_t nz0 = __ballot_sync(kFull, any0 != 0 && !isdef[0]);
            if (!(nz0 | nz1)) return;
            const int h = nz1 ? 1 : 0;
            const int src = 31 - __clz(nz1 ? nz1 : nz0);
            const int v = src + 32 * h;
            uint32_t cv[W];
#pragma unroll
            for (int w = 0; w < W; ++w) cv[w] = __shfl_sync(kFull, h ? c[1][w] : c[0][w], src);
            // youngest live class with coefficient 1 dies
            int slot = -1, age = -1;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = cv[w];
                while (bits) {
                    const int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int ag = brank()[32 * w + s];
                    if (ag > age) { age = ag; slot = 32 * w + s; }
                }
            }
            const int sw = slot >> 5;
            const uint32_t sb = 1u << (slot & 31);
            if (age < zero0) {  // non-zero persistence: keep a record
                if (n1 < rcap()) {
                    if (lane == 0) {
                        rec()[n1] = (uint32_t)age;
                        rec()[rcap() + n1] = (uint32_t)rcur;
                        rec()[2 * rcap() + n1] = (uint32_t)tri_index(a, b, v);
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
                // (the dying cocycle vanishes on every edge older than its birth)
                for (int q = (age & ~31) + lane; q < upto; q += 32) {
                    uint32_t* e = phi() + (size_t)q * W;
                    if (e[sw] & sb) {
                        const uint32_t pq = P()[q];
                        uint32_t* si = S() + p_i(pq) * W;
                        uint32_t* sj = S() + p_j(pq) * W;
#pragma unroll
                        for (int w = 0; w < W; ++w) {
                            e[w] ^= cv[w];
                            if (cv[w]) { atomicOr(si + w, cv[w]); atomicOr(sj + w, cv[w]); }
                        }
                    }
                }
                s_changed = true;
                __syncwarp();
            }
        }
    }

    // ------------------------------------------------------------------ one single (untied) edge
    // Stage B of the sweep pipeline: everything about edge `pij` that does not depend on PHI.
    struct Edge {
        uint32_t pij;
        uint32_t ta[2], tb[2];  // ranks of (i, v) and (j, v) for the apexes v = lane, lane + 32
    };
    __device__ __forceinline__ Edge fetch_edge(uint32_t pij) const {
        Edge ed;
        ed.pij = pij;
        const int i = p_i(pij), j = p_j(pij);
        const uint16_t* Ti = T() + i * ldtv();
        const uint16_t* Tj = T() + j * ldtv();
        const int v1 = lane + 32;
        ed.ta[0] = 0xFFFFu; ed.tb[0] = 0xFFFFu;
        ed.ta[1] = 0xFFFFu; ed.tb[1] = 0xFFFFu;
        if (NT >= 32 || lane < n()) { ed.ta[0] = Ti[lane]; ed.tb[0] = Tj[lane]; }
        if (v1 < n()) { ed.ta[1] = Ti[v1]; ed.tb[1] = Tj[v1]; }
        return ed;
    }
    // Stage C: the PHI-dependent part
    __device__ __forceinline__ void process_edge(int r, const Edge& ed) {
        if (ed.pij & kMst) return;  // PHI[r] stays 0: live cocycles extend by 0 over a merging edge
        const int i = p_i(ed.pij), j = p_j(ed.pij);
        bool in[2];
        in[0] = ed.ta[0] < (uint32_t)r && ed.tb[0] < (uint32_t)r;
        in[1] = ed.ta[1] < (uint32_t)r && ed.tb[1] < (uint32_t)r;
        const uint32_t G0 = __ballot_sync(kFull, in[0]);
        const uint32_t G1 = __ballot_sync(kFull, in[1]);
        if (!(G0 | G1)) {  // no apex yet: a real class is born
            birth(r, r);
            return;
        }
        // (the sweep only comes here with a class alive: with none it jumps from birth to birth, and a
        // birth has no apex; r < phicap() is the sweep's loop bound)
        // apparent pair (e, top triangle): extend every live cocycle over e, test the other apexes
        uint32_t c[2][W];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int w = 0; w < W; ++w) c[h][w] = 0;
            if (in[h]) {
                const uint32_t* pa = phi() + (size_t)ed.ta[h] * W;
                const uint32_t* pb = phi() + (size_t)ed.tb[h] * W;
#pragma unroll
                for (int w = 0; w < W; ++w) c[h][w] = pa[w] ^ pb[w];
            }
        }
        const int vtop = G1 ? 63 - __clz(G1) : 31 - __clz(G0);
        uint32_t xtop[W];
#pragma unroll
        for (int w = 0; w < W; ++w)
            xtop[w] = __shfl_sync(kFull, (vtop >> 5) ? c[1][w] : c[0][w], vtop & 31) & live[w];
        uint32_t anyx = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) anyx |= xtop[w];
        if (anyx) {   // (PHI[r] is zero already otherwise)
            if (lane == 0) {
                uint32_t* si = S() + i * W;
                uint32_t* sj = S() + j * W;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    phi()[(size_t)r * W + w] = xtop[w];
                    si[w] |= xtop[w];
                    sj[w] |= xtop[w];
                }
            }
            s_changed = true;
        }
        uint32_t anyc = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool use = in[h] && (lane + 32 * h) != vtop;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                c[h][w] = use ? ((c[h][w] ^ xtop[w]) & live[w]) : 0u;
                anyc |= c[h][w];
            }
        }
        __syncwarp();
        if (__ballot_sync(kFull, anyc != 0)) {
            const bool nodef[2] = {false, false};
            resolve(i, j, r, r, r + 1, c, nodef);
        }
    }

    // ------------------------------------------------------------------ a run of equal-length edges
    // triangles (a, b, v), v < b < a, v in (G0, G1); vdef >= 0: (a,b) is an apparent run edge defined
    // by apex vdef (the top of G)
    __device__ __forceinline__ void run_group(int a, int b, int rab, uint32_t G0, uint32_t G1, int vdef,
                                              int r0, int r1) {
        const uint16_t* Ta = T() + a * ldtv();
        const uint16_t* Tb = T() + b * ldtv();
        uint32_t pe[W];
        if (vdef >= 0) {
            const uint32_t* x = phi() + (size_t)Ta[vdef] * W;
            const uint32_t* y = phi() + (size_t)Tb[vdef] * W;
#pragma unroll
            for (int w = 0; w < W; ++w) pe[w] = (x[w] ^ y[w]) & live[w];
            if (lane == 0) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi()[(size_t)rab * W + w] = pe[w];
            }
        } else {
#pragma unroll
            for (int w = 0; w < W; ++w) pe[w] = phi()[(size_t)rab * W + w];
        }
        uint32_t c[2][W];
        bool isdef[2] = {false, false};
        int defq[2] = {-1, -1};
        uint32_t anyc = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int v = lane + 32 * h;
            const bool in = ((h ? G1 : G0) >> lane) & 1u;
#pragma unroll
            for (int w = 0; w < W; ++w) c[h][w] = 0;
            if (in && v != vdef) {
                const int ra = Ta[v], rb = Tb[v];
                const uint32_t* x = phi() + (size_t)ra * W;
                const uint32_t* y = phi() + (size_t)rb * W;
#pragma unroll
                for (int w = 0; w < W; ++w) c[h][w] = (pe[w] ^ x[w] ^ y[w]) & live[w];
                if (ra >= r0 && defv()[ra - r0] == b) { isdef[h] = true; defq[h] = ra; }
                else if (rb >= r0 && defv()[rb - r0] == a) { isdef[h] = true; defq[h] = rb; }
                if (!isdef[h]) {
#pragma unroll
                    for (int w = 0; w < W; ++w) anyc |= c[h][w];
                }
            }
        }
        __syncwarp();
        if (__ballot_sync(kFull, anyc != 0)) resolve(a, b, r0, r0, r1, c, isdef);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (isdef[h]) {
#pragma unroll
                for (int w = 0; w < W; ++w) phi()[(size_t)defq[h] * W + w] = c[h][w] & live[w];
            }
        }
        __syncwarp();
    }

    __device__ __forceinline__ void tie_run(int r0, int r1) {
        // a run that reaches beyond this tier's PHI capacity only matters if a class is alive in it
        // or born in it
        const bool beyond = r1 > phicap();
        if (beyond && live_any()) { overflow = true; return; }
        // pass 1 (rank order): apparent pairs inside the run take no slot.  A cycle-creating run edge
        // is apparent iff its first cofacet (largest apex among the triangles present once the whole
        // run has entered) has it as youngest edge.
        for (int pr = r0; pr < r1 && !overflow; ++pr) {
            const uint32_t pij = P()[pr];
            uint8_t dv = 254;  // merging edge
            if (!(pij & kMst)) {
                const int i = p_i(pij), j = p_j(pij);
                const uint16_t* Ti = T() + i * ldtv();
                const uint16_t* Tj = T() + j * ldtv();
                const int v1 = lane + 32;
                const bool in0 = lane < n() && Ti[lane] < r1 && Tj[lane] < r1;
                const bool in1 = v1 < n() && Ti[v1] < r1 && Tj[v1] < r1;
                const uint32_t G0 = __ballot_sync(kFull, in0), G1 = __ballot_sync(kFull, in1);
                dv = 255;
                if (G0 | G1) {
                    const int vt = G1 ? 63 - __clz(G1) : 31 - __clz(G0);
                    if (Ti[vt] < pr && Tj[vt] < pr) dv = (uint8_t)vt;
                }
                if (dv == 255) {
                    if (beyond) { overflow = true; return; }
                    birth(pr, r1);
                }
            }
            if (lane == 0) defv()[pr - r0] = dv;
        }
        __syncwarp();
        if (overflow || !live_any()) return;
        // pass 2: the run's triangles in descending index: a desc, b desc, apex c desc (c < b < a)
        for (int a = n() - 1; a >= 2; --a) {
            const uint16_t* Ta = T() + a * ldtv();
            const int vb1 = lane + 32;
            const uint32_t B0 = __ballot_sync(kFull, lane < a && Ta[lane] < r1);
            const uint32_t B1 = __ballot_sync(kFull, vb1 < a && Ta[vb1] < r1);
            for (int hb = 1; hb >= 0; --hb) {
                uint32_t bits = hb ? B1 : B0;
                while (bits) {
                    const int bb = 32 * hb + 31 - __clz(bits);
                    bits &= ~(1u << (bb & 31));
                    const uint16_t* Tb = T() + bb * ldtv();
                    const int rab = Ta[bb];
                    const bool ab_run = rab >= r0;
                    bool in[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int v = lane + 32 * h;
                        in[h] = false;
                        if (v < bb) {
                            const int ra = Ta[v], rb = Tb[v];
                            in[h] = ra < r1 && rb < r1 && (ab_run || ra >= r0 || rb >= r0);
                        }
                    }
                    const uint32_t G0 = __ballot_sync(kFull, in[0]), G1 = __ballot_sync(kFull, in[1]);
                    if (!(G0 | G1)) continue;
                    int vdef = -1;
                    if (ab_run) {
                        const int dv = defv()[rab - r0];
                        if (dv < 254 && dv < bb) vdef = dv;
                    }
                    run_group(a, bb, rab, G0, G1, vdef, r0, r1);
                    if (overflow || !live_any()) return;
                }
            }
        }
    }

    // ------------------------------------------------------------------ radix sort of (K(), P())
    // `varying`: bits in which the valid keys differ.  A byte that is the same in every key needs no
    // pass (EEG distances share their top byte); the padding keys sit at the end and stay there
    __device__ __forceinline__ void sort_edges(uint32_t varying) {
        if constexpr (kRS) {
            // keys in registers (chunk t of a lane = element 32 t + lane): a pass scatters them into the
            // one key buffer and reads them back; only the payload needs a second buffer
            constexpr int NCH = (NT * (NT - 1) / 2 + 31) / 32;
            uint32_t kr[NCH];
#pragma unroll
            for (int t = 0; t < NCH; ++t) kr[t] = K()[32 * t + lane];
            uint16_t* srcP = P(); uint16_t* dstP = P2();
            uint32_t* h32 = reinterpret_cast<uint32_t*>(hist());
            const uint32_t lt = lanemask_lt();
            int done = 0;
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 8 * pass;
                if (!((varying >> shift) & 255u)) continue;
                ++done;
#pragma unroll
                for (int t = 0; t < 4; ++t) h32[lane + 32 * t] = 0;
                __syncwarp();
#pragma unroll
                for (int t = 0; t < NCH; ++t) {
                    const uint32_t dg = (kr[t] >> shift) & 255u;
                    atomicAdd(h32 + (dg >> 1), 1u << (16 * (dg & 1u)));
                }
                __syncwarp();
                uint32_t loc[8], sum = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) { loc[t] = hist()[lane * 8 + t]; sum += loc[t]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += y;
                }
                uint32_t run = incl - sum;
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; ++t) { hist()[lane * 8 + t] = (uint16_t)run; run += loc[t]; }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < NCH; ++t) {
                    const uint32_t key = kr[t];
                    const uint16_t pay = srcP[32 * t + lane];
                    const uint32_t dg = (key >> shift) & 255u;
                    const uint32_t peers = __match_any_sync(kFull, dg);
                    const uint32_t h0 = hist()[dg];
                    const uint32_t pos = h0 + __popc(peers & lt);
                    __syncwarp();
                    K()[pos] = key;
                    dstP[pos] = pay;
                    if ((peers & lt) == 0) hist()[dg] = (uint16_t)(h0 + __popc(peers));
                    __syncwarp();
                }
#pragma unroll
                for (int t = 0; t < NCH; ++t) kr[t] = K()[32 * t + lane];
                uint16_t* tp = srcP; srcP = dstP; dstP = tp;
            }
            if (done & 1) {
                for (int k = lane; k < epad(); k += 32) P()[k] = P2()[k];
                __syncwarp();
            }
            return;
        }
        uint32_t* srcK = K(); uint16_t* srcP = P();
        uint32_t* dstK = K2(); uint16_t* dstP = P2();
        const uint32_t lt = lanemask_lt();
        int done = 0;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 8 * pass;
            if (!((varying >> shift) & 255u)) continue;
            ++done;
#pragma unroll
            for (int t = 0; t < 4; ++t) reinterpret_cast<uint32_t*>(hist())[lane + 32 * t] = 0;
            __syncwarp();
            // digit histogram: shared-memory atomics, no ordering needed here (independent iterations)
            // (two 16-bit counters per word: counts stay below 2^16, so the halves never carry)
            for (int k0 = 0; k0 < epad(); k0 += 32) {
                const uint32_t dg = (srcK[k0 + lane] >> shift) & 255u;
                atomicAdd(reinterpret_cast<uint32_t*>(hist()) + (dg >> 1), 1u << (16 * (dg & 1u)));
            }
            __syncwarp();
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { loc[t] = hist()[lane * 8 + t]; sum += loc[t]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
            }
            uint32_t run = incl - sum;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) { hist()[lane * 8 + t] = (uint16_t)run; run += loc[t]; }
            __syncwarp();
            for (int k0 = 0; k0 < epad(); k0 += 32) {
                const uint32_t key = srcK[k0 + lane];
                const uint16_t pay = srcP[k0 + lane];
                const uint32_t dg = (key >> shift) & 255u;
                const uint32_t peers = __match_any_sync(kFull, dg);
                const uint32_t pos = hist()[dg] + __popc(peers & lt);
                __syncwarp();
                dstK[pos] = key;
                dstP[pos] = pay;
                if ((peers & lt) == 0) hist()[dg] = (uint16_t)(hist()[dg] + __popc(peers));
                __syncwarp();
            }
            uint32_t* tk = srcK; srcK = dstK; dstK = tk;
            uint16_t* tp = srcP; srcP = dstP; dstP = tp;
        }
        if (done & 1) {  // an odd number of passes left the result in the ping-pong buffers
            for (int k = lane; k < epad(); k += 32) { K()[k] = K2()[k]; P()[k] = P2()[k]; }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ one window
    __device__ __forceinline__ void run(const Params& p, int b) {
        Db = p.D + (size_t)b * p.strideB;
        overflow = false;
        n0 = n1 = 0;
        ncomp = n();
#pragma unroll
        for (int w = 0; w < W; ++w) live[w] = used[w] = 0;
        // ---- keys, initial order = descending edge index
        int valid = 0, nan_seen = 0;
        uint32_t k_or = 0, k_and = 0xFFFFFFFFu;
        for (int k = e() + lane; k < epad(); k += 32) { K()[k] = 0xFFFFFFFFu; P()[k] = 0; }
        for (int row = 0; row < n() - 1; ++row) {
            for (int i = row + 1 + lane; i < n(); i += 32) {
                const float d = __ldg(Db + (d_rowoff(row, ld, n()) + i)) + 0.0f;
                const bool ok = d <= p.thresh;
                nan_seen |= (d != d);
                const int k = e() - 1 - (c2(i) + row);
                const uint32_t key = ok ? float_key(d) : 0xFFFFFFFFu;
                K()[k] = key;
                P()[k] = (uint16_t)((i << 6) | row);
                valid += ok;
                k_or |= key; k_and &= key;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            valid += __shfl_xor_sync(kFull, valid, o);
            nan_seen |= __shfl_xor_sync(kFull, nan_seen, o);
            k_or |= __shfl_xor_sync(kFull, k_or, o);
            k_and &= __shfl_xor_sync(kFull, k_and, o);
        }
        m = valid;
        __syncwarp();
        // with absent edges (d > thresh, NaN) in between, every pass runs
        sort_edges(m < e() ? kFull : (k_or ^ k_and));
        __syncwarp();
        // (after the sort: in the RSW layouts the digit counters reach into these arrays)
        for (int v = lane; v < n(); v += 32) { comp()[v] = (uint8_t)v; eld()[v] = (uint8_t)v; }
        // ---- tie flags (the keys are about to be overwritten by the rank matrix)
        for (int k0 = 0; k0 < m; k0 += 32) {
            const int r = k0 + lane;
            if (r < m) {
                const uint32_t kr = K()[r];
                uint32_t f = 0;
                if (r + 1 < m && K()[r + 1] == kr) f |= kTie;
                if (r > 0 && K()[r - 1] == kr) f |= kTiePrev;
                if (f) P()[r] |= (uint16_t)f;
            }
        }
        __syncwarp();
        // ---- H0: Kruskal, 32 edges checked per step
        for (int k0 = 0; k0 < m && ncomp > 1; k0 += 32) {
            const int r = k0 + lane;
            const uint32_t pij = r < m ? P()[r] : 0u;
            bool cand = false;
            if (r < m) cand = comp()[p_i(pij)] != comp()[p_j(pij)];
            uint32_t bal = __ballot_sync(kFull, cand);
            while (bal && ncomp > 1) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const uint32_t q = __shfl_sync(kFull, pij, src);
                const int i = p_i(q), j = p_j(q);
                const int ci = comp()[i], cj = comp()[j];
                if (ci == cj) continue;
                const int ei = eld()[ci], ej = eld()[cj];
                const float d = key_float(K()[k0 + src]);
                if (d != 0.0f) {
                    if (lane == 0) {
                        const size_t o = ((size_t)b * n() + n0) * 2;
                        p.bd0[o] = 0.0f;
                        p.bd0[o + 1] = d;
                        if (p.pr0) { p.pr0[o] = min(ei, ej); p.pr0[o + 1] = c2(i) + j; }
                    }
                    ++n0;
                }
                __syncwarp();
                for (int v = lane; v < n(); v += 32)
                    if (comp()[v] == ci) comp()[v] = (uint8_t)cj;
                if (lane == 0) { eld()[cj] = (uint8_t)max(ei, ej); P()[k0 + src] |= (uint16_t)kMst; }
                __syncwarp();
                --ncomp;
            }
        }
        __syncwarp();
        // ---- H0 essentials: eldest vertex of every surviving component, ascending (written now: the
        //      component arrays are about to become the sweep's vertex summary S; a window that a later
        //      tier redoes simply writes them again)
        {
            int base = n0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int v = lane + 32 * h;
                const bool is = v < n() && eld()[comp()[v]] == v;
                const uint32_t bal = __ballot_sync(kFull, is);
                if (is) {
                    const size_t o = ((size_t)b * n() + base + __popc(bal & lanemask_lt())) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = v; p.pr0[o + 1] = -1; }
                }
                base += __popc(bal);
            }
            n0 = base;
        }
        __syncwarp();
        // ---- rank matrix T (0xFFFF = edge absent) over region A
        {
            uint32_t* T32 = reinterpret_cast<uint32_t*>(T());
            const int words = n() * ldtv() / 2;
            for (int q = lane; q < words; q += 32) T32[q] = 0xFFFFFFFFu;
            __syncwarp();
            for (int k0 = 0; k0 < m; k0 += 32) {
                const int r = k0 + lane;
                if (r < m) {
                    const uint32_t pij = P()[r];
                    const int i = p_i(pij), j = p_j(pij);
                    T()[i * ldtv() + j] = (uint16_t)r;
                    T()[j * ldtv() + i] = (uint16_t)r;
                }
            }
            __syncwarp();
        }
        // ---- which edges can give birth to a real class: no apex at their own time.  Adjacency bit masks
        //      (two words per vertex, in the free sort buffer) grow chunk by chunk: an edge whose end points
        //      already share a neighbour among the edges of EARLIER chunks has an apex, which settles nearly
        //      every edge once the graph is a few hundred edges dense; only the others take the exact test
        //      (packed u16 min over v of max(T()[i][v], T()[j][v])).  Tie-run members are always visited.
        {
            const int nw2 = ldtv() / 2;
            uint32_t* adj = K2();
            for (int q = lane; q < 2 * n(); q += 32) adj[q] = 0;
            __syncwarp();
            for (int k0 = 0; k0 < epad(); k0 += 32) {
                const int r = k0 + lane;
                const bool act = r < m;
                const uint32_t pij = act ? P()[r] : 0u;
                const int i = p_i(pij), j = p_j(pij);
                bool vis = false, exact = false;
                if (act && !(pij & kMst)) {
                    if (pij & (kTie | kTiePrev)) vis = true;
                    else exact = ((adj[2 * i] & adj[2 * j]) | (adj[2 * i + 1] & adj[2 * j + 1])) == 0;
                }
                if (__any_sync(kFull, exact)) {
                    if (exact) {
                        const uint32_t* Ti = reinterpret_cast<const uint32_t*>(T() + i * ldtv());
                        const uint32_t* Tj = reinterpret_cast<const uint32_t*>(T() + j * ldtv());
                        uint32_t mn = 0xFFFFFFFFu;
                        for (int w = 0; w < nw2; ++w) mn = __vminu2(mn, __vmaxu2(Ti[w], Tj[w]));
                        const uint32_t mm = min(mn & 0xFFFFu, mn >> 16);
                        vis = mm > (uint32_t)r;
                    }
                }
                const uint32_t bal = __ballot_sync(kFull, vis);
                if (lane == 0) visit()[k0 >> 5] = bal;
                if (act) {
                    atomicOr(adj + 2 * i + (j >> 5), 1u << (j & 31));
                    atomicOr(adj + 2 * j + (i >> 5), 1u << (i & 31));
                }
                __syncwarp();
            }
        }
        // ---- PHI := 0, S := 0
        for (int q = lane; q < min(m, phicap()) * W; q += 32) phi()[q] = 0;
        for (int q = lane; q < n() * W; q += 32) S()[q] = 0;
        s_changed = false;
        __syncwarp();
        // ---- the sweep through the live spans
        {
            // ranks below the PHI capacity of this tier; beyond it only the question "is a class alive
            // or born there" is left (then the window belongs to the next tier)
            const int mlim = min(m, phicap());
            const int nwords = epad() >> 5;
            int r = 0;
            while (r < mlim && !overflow) {
                if (!live_any()) {
                    // jump to the next rank where a class can be born
                    int wq = r >> 5;
                    uint32_t bits = visit()[wq] & (kFull << (r & 31));
                    while (!bits && ++wq < nwords) bits = visit()[wq];
                    if (!bits) { r = m; break; }
                    r = 32 * wq + __ffs(bits) - 1;
                    if (r >= mlim) break;
                }
                // ---- the 32 ranks of r's chunk at once: which of them can a live cocycle see?  A non-tied edge
                //      that gives no birth and whose end points carry no live cocycle bit (S) keeps PHI = 0 and
                //      has coboundary 0 on every triangle: it is not visited at all
                const int rb = r & ~31, rr = rb + lane;
                const uint32_t pl = rr < mlim ? P()[rr] : kMst;
                const uint32_t vbits = visit()[rb >> 5];
                auto hot_ranks = [&](int from) -> uint32_t {
                    bool hot = false;
                    if (rr >= from && !(pl & kMst)) {
                        hot = (vbits >> lane) & 1u;   // births and tie-run members
                        if (!hot) {
                            const uint32_t* si = S() + p_i(pl) * W;
                            const uint32_t* sj = S() + p_j(pl) * W;
#pragma unroll
                            for (int w = 0; w < W; ++w) hot |= ((si[w] | sj[w]) & live[w]) != 0;
                        }
                    }
                    return __ballot_sync(kFull, hot);
                };
                uint32_t bal = hot_ranks(r);
                s_changed = false;
                r = min(rb + 32, mlim);   // unless something below says otherwise
                while (bal) {
                    const int src = __ffs(bal) - 1;
                    const int re = rb + src;
                    const uint32_t pij = __shfl_sync(kFull, pl, src);
                    if (pij & (kTie | kTiePrev)) {
                        int r0 = re;
                        while (r0 > 0 && (P()[r0 - 1] & kTie)) --r0;
                        int r1 = re;
                        while (P()[r1] & kTie) ++r1;
                        ++r1;
                        tie_run(r0, r1);
                        // a run defines PHI values wholesale: every vertex counts as touched from here on
                        for (int q = lane; q < n() * W; q += 32) S()[q] = kFull;
                        __syncwarp();
                        r = r1;
                        break;
                    }
                    process_edge(re, fetch_edge(pij));
                    if (overflow) break;
                    if (!live_any()) { r = re + 1; break; }
                    if (s_changed) {   // S gained bits: later ranks of the chunk may have become visible
                        __syncwarp();
                        bal = hot_ranks(re + 1);
                        s_changed = false;
                    } else {
                        bal &= bal - 1;
                    }
                }
            }
            if (mlim < m && !overflow && r < m) {
                if (live_any()) overflow = true;
                while (r < m && !overflow) {   // nothing alive: births only
                    int wq = r >> 5;
                    uint32_t bits = visit()[wq] & (kFull << (r & 31));
                    while (!bits && ++wq < nwords) bits = visit()[wq];
                    if (!bits) break;
                    r = 32 * wq + __ffs(bits) - 1;
                    if (r >= m) break;
                    if (P()[r] & (kTie | kTiePrev)) {   // a tie run is visited whether or not it holds a birth
                        int r0 = r;
                        while (r0 > 0 && (P()[r0 - 1] & kTie)) --r0;
                        int r1 = r;
                        while (P()[r1] & kTie) ++r1;
                        ++r1;
                        tie_run(r0, r1);
                        if (live_any()) overflow = true;
                        r = r1;
                    } else {
                        overflow = true;   // a birth beyond the capacity
                    }
                }
            }
        }
        if (!overflow) {
            // cycles still alive at thresh are essential
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t bits = live[w];
                while (bits) {
                    const int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (n1 < rcap()) {
                        if (lane == 0) {
                            rec()[n1] = brank()[32 * w + s];
                            rec()[rcap() + n1] = kEssential;
                            rec()[2 * rcap() + n1] = kEssential;
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
        // ---- H1 rows in ripser's order: descending birth rank
        int st = nan_seen ? TDA_ST_NAN_INPUT : 0;
        for (int k = lane; k < n1; k += 32) {
            const uint32_t br = rec()[k];
            int pos = 0;
            for (int t = 0; t < n1; ++t) pos += rec()[t] > br;
            if (pos < p.cap1) {
                const size_t o = ((size_t)b * p.cap1 + pos) * 2;
                const uint32_t dr = rec()[rcap() + k], tr = rec()[2 * rcap() + k];
                const uint32_t pij = P()[br];
                p.bd1[o] = dist(p_i(pij), p_j(pij));
                float dth = __int_as_float(0x7F800000);
                if (tr != kEssential) { const uint32_t pd = P()[dr]; dth = dist(p_i(pd), p_j(pd)); }
                p.bd1[o + 1] = dth;
                if (p.pr1) {
                    p.pr1[o] = c2(p_i(pij)) + (p_j(pij));
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

template <int W, bool PHI_GLOBAL, int NT, int RSW = 0>
__global__ void __launch_bounds__(RSW ? 32 * RSW : 256, RSW ? 2 : 1) rips_small_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef Layout<W, PHI_GLOBAL, RSW> L;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    const int N = NT > 0 ? NT : p.N;
    Warp<W, PHI_GLOBAL, NT, RSW> s;
    s.base = smem_raw + (size_t)wib * L::bytes(N);
    s.lane = lane;
    s.Nrt = N;
    s.ld = p.ld;
    s.phi_g = PHI_GLOBAL ? p.phi_global + (size_t)gw * c2(N) * W : nullptr;
    s.rec_g = L::kRecGlobal ? p.rec_global + (size_t)gw * 3 * L::recs(N) : nullptr;
    s.defv_g = p.defv_global + (size_t)gw * L::epad(N);
    const int total = p.worklist ? *p.n_work : p.B;
    for (int t = gw; t < total; t += nw) {
        const int b = p.worklist ? p.worklist[t] : t;
        if (t + nw < total) {
            // the next window's matrix starts its way from HBM to L2 now: the key phase of a window is
            // otherwise one exposed memory round trip per warp
            const int bn = p.worklist ? p.worklist[t + nw] : t + nw;
            const char* nx = (const char*)(p.D + (size_t)bn * p.strideB);
            const int bytes = (p.ld ? (N - 1) * p.ld + N : c2(N)) * 4;
            for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
        }
        s.run(p, b);
        __syncwarp();
    }
}

// workspace layout: counters ; list1[B] ; list2[B] ; tie-run scratch ; phi scratch ; rec scratch
constexpr int kLastW = 64;
constexpr int kLastGrid = 148;   // one single-warp CTA per SM on the last tier
constexpr int kMaxWarps = 148 * 24;  // upper bound on resident warps of any tier
struct WsLayout {
    size_t counters, list1, list2, defv, phi, rec, rec0, total;
};
static WsLayout ws_layout(int B, int N) {
    WsLayout w;
    size_t o = 0;
    w.counters = o; o += 64;
    w.list1 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    w.list2 = o; o += ((size_t)B * 4 + 63) & ~(size_t)63;
    w.defv = o; o += (size_t)kMaxWarps * Layout<2, false>::epad(N);
    w.phi = o; o += (size_t)kLastGrid * c2(N) * kLastW * 4;
    w.rec = o; o += (size_t)kLastGrid * 4 * Layout<kLastW, true>::recs(N) * 4;
    w.rec0 = o; o += (size_t)kMaxWarps * 3 * Layout<1, false>::recs(N) * 4;
    w.total = o;
    return w;
}

template <int W, bool G, int NT, int RSW = 0>
static cudaError_t launch_tier(const Params& p, int warps_per_block, int grid, cudaStream_t st) {
    ProfScope prof(W == 1 ? "rips_small_w1" : (W == 2 ? "rips_small_w2" : (W == 4 ? "rips_small_w4" : "rips_small_w64")), st);
    size_t smem = Layout<W, G, RSW>::bytes(p.N) * warps_per_block;
    cudaError_t e = cudaFuncSetAttribute(rips_small_kernel<W, G, NT, RSW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rips_small_kernel<W, G, NT, RSW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    rips_small_kernel<W, G, NT, RSW><<<grid, warps_per_block * 32, smem, st>>>(p);
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
    if (!D || !bd0 || !bd1 || !counts || !status || !ws || B < 0 || cap1 < 0 || (ld != 0 && ld < N)) return TDA_E_ARG;
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
    p.D = D; p.strideB = strideB ? strideB : (ld ? (long long)N * ld : (long long)c2(N)); p.ld = ld; p.N = N; p.B = B;
    p.thresh = thresh;
    p.bd0 = bd0; p.pr0 = pr0; p.bd1 = bd1; p.pr1 = pr1; p.counts = counts; p.status = status; p.cap1 = cap1;
    p.phi_global = nullptr; p.rec_global = nullptr;
    p.defv_global = (uint8_t*)(w8 + wl.defv);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > 148) sms = 148;
    // tiers 1-2: PHI per edge, lanes = apexes
    const bool tier0 = (N == 47);
    if (tier0) {
        // 47-point windows rarely hold more than 32 classes at once: a one-word tier first (register
        // sort, twelve warps per CTA, two CTAs per SM)
        p.worklist = nullptr; p.n_work = nullptr;
        p.overflow_list = (int*)(w8 + wl.list2); p.n_overflow = counters + 2;
        p.rec_global = (uint32_t*)(w8 + wl.rec0);
        constexpr int wpb = 12;
        const int per_sm = 2;
        long long need = ((long long)B + wpb - 1) / wpb;
        int grid = (int)((long long)sms * per_sm < need ? (long long)sms * per_sm : need);
        e = launch_tier<1, false, 47, wpb>(p, wpb, grid, st);
        if (e != cudaSuccess) return (int)e;
    }
    {
        p.worklist = tier0 ? (const int*)(w8 + wl.list2) : nullptr; p.n_work = tier0 ? counters + 2 : nullptr;
        p.overflow_list = (int*)(w8 + wl.list1); p.n_overflow = counters + 0;
        int wpb = (int)(((227 * 1024) / 2 - 1024) / Layout<2, false>::bytes(N));
        if (wpb < 1) wpb = 1;
        if (wpb > 8) wpb = 8;
        size_t smem = Layout<2, false>::bytes(N) * wpb;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
        long long need = tier0 ? (long long)sms * per_sm : ((long long)B + wpb - 1) / wpb;
        int grid = (int)((long long)sms * per_sm < need ? (long long)sms * per_sm : need);
        e = (N == 47) ? launch_tier<2, false, 47>(p, wpb, grid, st) : launch_tier<2, false, 0>(p, wpb, grid, st);
        if (e != cudaSuccess) return (int)e;
    }
    {
        p.worklist = (const int*)(w8 + wl.list1); p.n_work = counters + 0;
        p.overflow_list = (int*)(w8 + wl.list2); p.n_overflow = counters + 1;
        e = launch_tier<4, false, 0>(p, 2, sms * 2, st);
        if (e != cudaSuccess) return (int)e;
    }
    // tier 3: W=64 with PHI in global scratch, handles every N<=64 input
    {
        p.worklist = (const int*)(w8 + wl.list2); p.n_work = counters + 1;
        p.overflow_list = nullptr; p.n_overflow = nullptr;
        p.phi_global = (uint32_t*)(w8 + wl.phi);
        p.rec_global = (uint32_t*)(w8 + wl.rec);
        e = launch_tier<kLastW, true, 0>(p, 1, kLastGrid, st);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
