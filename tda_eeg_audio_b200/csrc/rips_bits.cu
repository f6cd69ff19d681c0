// rips_bits.cu — ALTERNATIVE tiers 1-2 of the N <= 64 Rips engine (selected with TDA_RIPS_ENGINE=bits;
// the default tiers are rips_small.cu's, which measure ~8 % faster on 47-point EEG windows: 8.2 vs
// 7.6 M diagrams/s; kept as a second, independently written implementation under the same parity tests).
//
// Replaces ripser.ripser(dm, maxdim=1, thresh, distance_matrix=True) as called per window by
//   /root/reference/scripts/utils.py:131,140
//   /root/reference/scripts/tda_eeg_classification_v2.py:170-175
//
// One warp owns one window; everything up to the sweep is shared with rips_small.cu's design
// (per-warp stable LSD radix sort of the order-preserving integer image of the f32 lengths,
// initial order = descending edge index so stability gives Ripser's tie-break; Kruskal 32 edges
// per step; rank matrix T[i][v]; one packed-u16 min/max pass that finds every edge without an
// apex at its own time = the births).  The sweep is organised the other way round:
//   * a live 1-cocycle is a BIT MATRIX: M[v][l] holds, for class l, the bit row
//     { phi_l(v, u) : u } — two 32-bit halves, class index fastest, so 32 lanes reading their own
//     class's row of vertex v hit 32 different banks;
//   * LANE = CLASS (CPL classes per lane: 32 or 64 simultaneous classes).  For an edge e = (i, j)
//     with apex set G (two ballots over T), lane l gets the coboundary of its class on ALL
//     triangles (i, j, v), v in G, from x = (M[i][l] ^ M[j][l]) & G: phi_l(e) := x[vtop] (the
//     extension that keeps the top triangle closed, i.e. the apparent pair), c = x ^ (phi_l(e)?G:0).
//     c == 0 on every lane (the common case) means nothing dies;
//   * deaths walk the apexes downwards: the youngest class with a 1 at the top apex dies, the
//     other classes with a 1 there absorb it (their rows ^= the dying class's rows, only rows
//     where the dying class is non-zero are touched);
//   * a birth takes a free lane and clears its column; no scrubbing, no slot recycling pass;
//   * tie runs use the single exact rule of oracle/pcoh_large_model.cpp: births and definitions
//     in rank order, then repeatedly the triangle with the LARGEST index and a non-zero
//     coboundary mask among the run's triangles kills its youngest class.
// A window with more simultaneous classes than a tier holds goes to the next tier through a
// device-side list (tier 3 = rips_small.cu's global-memory kernel, which handles every input).
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "common.cuh"
#include "rips_small.cuh"
#include "tda_b200.h"

namespace tda {
namespace rips_small {

template <int CPL, int NT> struct BitLayout {
    // hi half of a row covers apexes 32..N-1: 16 bits are enough up to 48 points
    typedef typename std::conditional<(NT > 0 && NT <= 48), uint16_t, uint32_t>::type hi_t;
    static constexpr int NC = 32 * CPL;      // classes
    static constexpr int LDM = NC + 1;       // row stride of the class matrices: rows AND columns conflict-free
    static __host__ __device__ int epad(int N) { return (c2(N) + 31) & ~31; }
    static __host__ __device__ int ldt(int N) { return ((((N + 1) / 2) | 1) * 2); }
    static __host__ __device__ int recs() { return 96 * CPL; }
    static __host__ __device__ size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }
    static __host__ __device__ size_t region_a(int N) {  // sort keys + histogram | rank matrix T
        size_t s1 = (size_t)epad(N) * 4 + 1024, s2 = (size_t)N * ldt(N) * 2;
        return a16(s1 > s2 ? s1 : s2);
    }
    static __host__ __device__ size_t mlo_bytes(int N) { return a16((size_t)N * LDM * 4); }
    static __host__ __device__ size_t mhi_bytes(int N) { return a16((size_t)N * LDM * sizeof(hi_t)); }
    static __host__ __device__ size_t region_c(int N) {  // sort ping-pong | class matrices
        size_t s1 = (size_t)epad(N) * 6, s2 = mlo_bytes(N) + mhi_bytes(N);
        return a16(s1 > s2 ? s1 : s2);
    }
    static __host__ __device__ size_t off_p(int N) { return region_a(N) + region_c(N); }
    static __host__ __device__ size_t off_rec(int N) { return off_p(N) + a16((size_t)epad(N) * 2); }
    static __host__ __device__ size_t off_visit(int N) { return off_rec(N) + (size_t)recs() * 12; }
    static __host__ __device__ size_t bytes(int N) {
        return a16(off_visit(N) + (size_t)epad(N) / 8 + 2 * kMaxN);
    }
};

template <int CPL, int NT> struct BitWarp {
    typedef BitLayout<CPL, NT> L;
    typedef typename L::hi_t hi_t;
    static constexpr int NC = L::NC;
    static constexpr int LDM = L::LDM;
    unsigned char* base;
    const float* Db;
    int lane, Nrt, ld;
    __device__ __forceinline__ int n() const { return NT > 0 ? NT : Nrt; }
    __device__ __forceinline__ int e() const { return c2(n()); }
    __device__ __forceinline__ int epad() const { return L::epad(n()); }
    __device__ __forceinline__ int rcap() const { return L::recs(); }
    __device__ __forceinline__ int ldtv() const { return L::ldt(n()); }
    __device__ __forceinline__ uint32_t* K() const { return (uint32_t*)base; }
    __device__ __forceinline__ uint32_t* hist() const { return (uint32_t*)(base + (size_t)epad() * 4); }
    __device__ __forceinline__ uint16_t* T() const { return (uint16_t*)base; }
    __device__ __forceinline__ uint32_t* K2() const { return (uint32_t*)(base + L::region_a(n())); }
    __device__ __forceinline__ uint16_t* P2() const { return (uint16_t*)(base + L::region_a(n()) + (size_t)epad() * 4); }
    __device__ __forceinline__ uint32_t* Mlo() const { return (uint32_t*)(base + L::region_a(n())); }
    __device__ __forceinline__ hi_t* Mhi() const { return (hi_t*)(base + L::region_a(n()) + L::mlo_bytes(n())); }
    __device__ __forceinline__ uint16_t* P() const { return (uint16_t*)(base + L::off_p(n())); }
    __device__ __forceinline__ uint32_t* rec() const { return (uint32_t*)(base + L::off_rec(n())); }
    __device__ __forceinline__ uint32_t* visit() const { return (uint32_t*)(base + L::off_visit(n())); }
    __device__ __forceinline__ uint8_t* comp() const { return base + L::off_visit(n()) + (size_t)epad() / 8; }
    __device__ __forceinline__ uint8_t* eld() const { return comp() + kMaxN; }
    // ---- per-window state: `live` uniform, `brank` per lane (birth rank of the lane's classes)
    uint32_t live[CPL];
    int brank[CPL];
    int n0, n1, ncomp, m;
    bool overflow;

    __device__ __forceinline__ float dist(int a, int b) const {
        return __ldg(Db + (d_rowoff(min(a, b), ld, n()) + max(a, b))) + 0.0f;
    }
    __device__ __forceinline__ bool live_any() const {
        uint32_t a = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) a |= live[c];
        return a != 0;
    }
    __device__ __forceinline__ bool mine(int c) const { return (live[c] >> lane) & 1u; }

    // apexes v with both (i, v) and (j, v) of rank < lim: two ballots over the rank matrix
    __device__ __forceinline__ void apexes(int i, int j, int lim, uint32_t& G0, uint32_t& G1) const {
        const uint16_t* Ti = T() + i * ldtv();
        const uint16_t* Tj = T() + j * ldtv();
        const int v1 = lane + 32;
        bool in0 = false, in1 = false;
        if (lane < n()) in0 = Ti[lane] < lim && Tj[lane] < lim;
        if (v1 < n()) in1 = Ti[v1] < lim && Tj[v1] < lim;
        G0 = __ballot_sync(kFull, in0);
        G1 = __ballot_sync(kFull, in1);
    }
    static __device__ __forceinline__ int top_bit(uint32_t lo, uint32_t hi) {
        return hi ? 63 - __clz(hi) : (lo ? 31 - __clz(lo) : -1);
    }
    static __device__ __forceinline__ uint32_t bit_at(uint32_t lo, uint32_t hi, int v) {
        return ((v < 32 ? lo >> v : hi >> (v - 32)) & 1u);
    }
    // class row access: element (vertex v, class cl = lane + 32 c)
    __device__ __forceinline__ int midx(int v, int c) const { return v * LDM + 32 * c + lane; }

    // ------------------------------------------------------------------ birth
    __device__ __forceinline__ void birth(int r, int i, int j) {
        int slot = -1;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t f = ~live[c];
            if (slot < 0 && f) slot = 32 * c + __ffs(f) - 1;
        }
        if (slot < 0) { overflow = true; return; }
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t bit = (c == (slot >> 5)) ? (1u << (slot & 31)) : 0u;
            live[c] |= bit;
            if (bit && lane == (slot & 31)) brank[c] = r;
        }
        // clear the class's column and set phi(i, j) = 1
        for (int v = lane; v < n(); v += 32) {
            uint32_t lo = 0, hi = 0;
            if (v == i) { if (j < 32) lo = 1u << j; else hi = 1u << (j - 32); }
            if (v == j) { if (i < 32) lo = 1u << i; else hi = 1u << (i - 32); }
            Mlo()[v * LDM + slot] = lo;
            Mhi()[v * LDM + slot] = (hi_t)hi;
        }
        __syncwarp();
    }

    // coboundary masks of this lane's classes on the triangles (i, j, v), v in (G0, G1);
    // pe[c] = phi(i, j) of the class
    __device__ __forceinline__ void coboundary(int i, int j, uint32_t G0, uint32_t G1, uint32_t (&c0)[CPL],
                                               uint32_t (&c1)[CPL]) const {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            c0[c] = 0; c1[c] = 0;
            if (mine(c)) {
                const uint32_t il = Mlo()[midx(i, c)], jl = Mlo()[midx(j, c)];
                const uint32_t ih = Mhi()[midx(i, c)], jh = Mhi()[midx(j, c)];
                const uint32_t pe = bit_at(il, ih, j);
                c0[c] = ((il ^ jl) & G0) ^ (pe ? G0 : 0u);
                c1[c] = ((ih ^ jh) & G1) ^ (pe ? G1 : 0u);
            }
        }
    }

    // ------------------------------------------------------------------ one death
    // triangle (i, j, v) of an edge of rank rdeath: the youngest class with a 1 there dies, the
    // others absorb it.  c0/c1 (this lane's masks for edge (i, j)) are updated linearly.
    __device__ __forceinline__ void kill_at(int i, int j, int v, int rdeath, int zero0, uint32_t (&c0)[CPL],
                                            uint32_t (&c1)[CPL]) {
        uint32_t memb[CPL];
        int br = -1;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const bool is = bit_at(c0[c], c1[c], v) != 0;
            memb[c] = __ballot_sync(kFull, is);
            if (is) br = max(br, brank[c]);
        }
        const int brmax = __reduce_max_sync(kFull, br);
        int dl = 0, dc = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const uint32_t who = __ballot_sync(kFull, ((memb[c] >> lane) & 1u) && brank[c] == brmax);
            if (who) { dl = __ffs(who) - 1; dc = c; }
        }
        if (brmax < zero0) {  // non-zero persistence: keep a record
            if (n1 >= rcap()) { overflow = true; return; }
            if (lane == 0) {
                rec()[n1] = (uint32_t)brmax;
                rec()[rcap() + n1] = (uint32_t)rdeath;
                rec()[2 * rcap() + n1] = (uint32_t)tri_index(i, j, v);
            }
            ++n1;
        }
        uint32_t others = 0;
        uint32_t d0 = 0, d1 = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            if (c == dc) {
                memb[c] &= ~(1u << dl);
                live[c] &= ~(1u << dl);
                d0 = __shfl_sync(kFull, c0[c], dl);
                d1 = __shfl_sync(kFull, c1[c], dl);
                if (lane == dl) { c0[c] = 0; c1[c] = 0; }
            }
            others |= memb[c];
        }
        if (others) {
            // rows of the absorbing classes ^= rows of the dying class; lanes = vertices here
            const int dcol = 32 * dc + dl;
            uint32_t dlo[2] = {0, 0}, dhi[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int u = lane + 32 * h;
                if (u < n()) { dlo[h] = Mlo()[u * LDM + dcol]; dhi[h] = Mhi()[u * LDM + dcol]; }
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                uint32_t bits = memb[c];
                while (bits) {
                    const int col = 32 * c + __ffs(bits) - 1;
                    bits &= bits - 1;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = lane + 32 * h;
                        if (u < n() && (dlo[h] | dhi[h])) {
                            Mlo()[u * LDM + col] ^= dlo[h];
                            Mhi()[u * LDM + col] = (hi_t)(Mhi()[u * LDM + col] ^ dhi[h]);
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if ((memb[c] >> lane) & 1u) { c0[c] ^= d0; c1[c] ^= d1; }
            }
            __syncwarp();
        }
    }

    // ------------------------------------------------------------------ one untied edge
    __device__ __forceinline__ void single_edge(int r, uint32_t pij) {
        const int i = p_i(pij), j = p_j(pij);
        // the class rows do not depend on the apex set: fetch them together with the rank rows
        uint32_t il[CPL], jl[CPL], ih[CPL], jh[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            il[c] = Mlo()[midx(i, c)]; jl[c] = Mlo()[midx(j, c)];
            ih[c] = Mhi()[midx(i, c)]; jh[c] = Mhi()[midx(j, c)];
        }
        uint32_t G0, G1;
        apexes(i, j, r, G0, G1);
        if (!(G0 | G1)) { birth(r, i, j); return; }
        if (!live_any()) return;
        const int vtop = top_bit(G0, G1);
        uint32_t c0[CPL], c1[CPL];
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            c0[c] = 0; c1[c] = 0;
            if (mine(c)) {
                const uint32_t x0 = (il[c] ^ jl[c]) & G0, x1 = (ih[c] ^ jh[c]) & G1;
                if (bit_at(x0, x1, vtop)) {   // phi(e) := 1 keeps the top triangle closed
                    c0[c] = x0 ^ G0; c1[c] = x1 ^ G1;
                    if (j < 32) Mlo()[midx(i, c)] = il[c] | (1u << j); else Mhi()[midx(i, c)] = (hi_t)(ih[c] | (1u << (j - 32)));
                    if (i < 32) Mlo()[midx(j, c)] = jl[c] | (1u << i); else Mhi()[midx(j, c)] = (hi_t)(jh[c] | (1u << (i - 32)));
                } else {
                    c0[c] = x0; c1[c] = x1;
                }
                any |= c0[c] | c1[c];
            }
        }
        if (!__ballot_sync(kFull, any != 0)) return;
        // deaths, apexes downwards (the triangle index is monotone in the apex)
        while (!overflow) {
            int t = -1;
#pragma unroll
            for (int c = 0; c < CPL; ++c) t = max(t, top_bit(c0[c], c1[c]));
            const int v = __reduce_max_sync(kFull, t);
            if (v < 0) break;
            kill_at(i, j, v, r, r, c0, c1);
        }
        __syncwarp();
    }

    // ------------------------------------------------------------------ a run of equal-length edges
    __device__ __forceinline__ void tie_run(int r0, int r1) {
        // A. rank order: births take a lane; an apparent edge (its first cofacet, the largest apex
        //    among the triangles present once the whole run has entered, has it as youngest edge)
        //    gets phi(e) := phi(i, vt) ^ phi(j, vt) for every live class
        for (int pr = r0; pr < r1 && !overflow; ++pr) {
            const uint32_t pij = P()[pr];
            if (pij & kMst) continue;
            const int i = p_i(pij), j = p_j(pij);
            uint32_t G0, G1;
            apexes(i, j, r1, G0, G1);
            int vt = top_bit(G0, G1);
            if (vt >= 0 && !(T()[i * ldtv() + vt] < pr && T()[j * ldtv() + vt] < pr)) vt = -1;
            if (vt < 0) { birth(pr, i, j); continue; }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (mine(c)) {
                    const uint32_t il = Mlo()[midx(i, c)], jl = Mlo()[midx(j, c)];
                    const uint32_t ih = Mhi()[midx(i, c)], jh = Mhi()[midx(j, c)];
                    if (bit_at(il ^ jl, ih ^ jh, vt)) {
                        if (j < 32) Mlo()[midx(i, c)] = il | (1u << j); else Mhi()[midx(i, c)] = (hi_t)(ih | (1u << (j - 32)));
                        if (i < 32) Mlo()[midx(j, c)] = jl | (1u << i); else Mhi()[midx(j, c)] = (hi_t)(jh | (1u << (i - 32)));
                    }
                }
            }
            __syncwarp();
        }
        // B. deaths: among the triangles whose youngest edge is in the run, largest index first
        while (!overflow && live_any()) {
            uint32_t best = 0;
            int bpr = -1, bv = -1;
            for (int pr = r0; pr < r1; ++pr) {
                const uint32_t pij = P()[pr];
                if (pij & kMst) continue;
                const int i = p_i(pij), j = p_j(pij);
                uint32_t G0, G1;
                apexes(i, j, pr, G0, G1);
                if (!(G0 | G1)) continue;
                uint32_t c0[CPL], c1[CPL];
                coboundary(i, j, G0, G1, c0, c1);
                int t = -1;
#pragma unroll
                for (int c = 0; c < CPL; ++c) t = max(t, top_bit(c0[c], c1[c]));
                const int v = __reduce_max_sync(kFull, t);
                if (v >= 0) {
                    const uint32_t tri = (uint32_t)tri_index(i, j, v) + 1u;
                    if (tri > best) { best = tri; bpr = pr; bv = v; }
                }
            }
            if (!best) break;
            const uint32_t pij = P()[bpr];
            const int i = p_i(pij), j = p_j(pij);
            uint32_t G0, G1;
            apexes(i, j, bpr, G0, G1);
            uint32_t c0[CPL], c1[CPL];
            coboundary(i, j, G0, G1, c0, c1);
            kill_at(i, j, bv, bpr, r0, c0, c1);
        }
        __syncwarp();
    }

    // ------------------------------------------------------------------ radix sort of (K(), P())
    __device__ __forceinline__ void sort_edges(uint32_t varying) {
        uint32_t* srcK = K(); uint16_t* srcP = P();
        uint32_t* dstK = K2(); uint16_t* dstP = P2();
        const uint32_t lt = lanemask_lt();
        int done = 0;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 8 * pass;
            if (!((varying >> shift) & 255u)) continue;  // this byte is the same in every key
            ++done;
#pragma unroll
            for (int t = 0; t < 8; ++t) hist()[lane + 32 * t] = 0;
            __syncwarp();
            // digit histogram: shared-memory atomics, no ordering needed here (independent iterations)
            for (int k0 = 0; k0 < epad(); k0 += 32) atomicAdd(hist() + ((srcK[k0 + lane] >> shift) & 255u), 1u);
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
            for (int t = 0; t < 8; ++t) { hist()[lane * 8 + t] = run; run += loc[t]; }
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
                if ((peers & lt) == 0) hist()[dg] += __popc(peers);
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
        for (int c = 0; c < CPL; ++c) { live[c] = 0; brank[c] = -1; }
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
        for (int v = lane; v < n(); v += 32) { comp()[v] = (uint8_t)v; eld()[v] = (uint8_t)v; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            valid += __shfl_xor_sync(kFull, valid, o);
            nan_seen |= __shfl_xor_sync(kFull, nan_seen, o);
            k_or |= __shfl_xor_sync(kFull, k_or, o);
            k_and &= __shfl_xor_sync(kFull, k_and, o);
        }
        m = valid;
        __syncwarp();
        // bytes that are the same in every key need no pass (the padding keys sit at the end and stay
        // there); with absent edges (d > thresh, NaN) in between, every pass runs
        sort_edges(m < e() ? kFull : (k_or ^ k_and));
        __syncwarp();
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
        // ---- which edges can give birth to a real class: no apex at their own time
        //      (packed u16 min over v of max(T[i][v], T[j][v]); tie-run members are always visited)
        {
            const int nw2 = ldtv() / 2;
            for (int k0 = 0; k0 < epad(); k0 += 32) {
                const int r = k0 + lane;
                bool vis = false;
                if (r < m) {
                    const uint32_t pij = P()[r];
                    if (!(pij & kMst)) {
                        if (pij & (kTie | kTiePrev)) vis = true;
                        else {
                            const uint32_t* Ti = reinterpret_cast<const uint32_t*>(T() + (p_i(pij)) * ldtv());
                            const uint32_t* Tj = reinterpret_cast<const uint32_t*>(T() + (p_j(pij)) * ldtv());
                            uint32_t mn = 0xFFFFFFFFu;
                            for (int w = 0; w < nw2; ++w) mn = __vminu2(mn, __vmaxu2(Ti[w], Tj[w]));
                            vis = min(mn & 0xFFFFu, mn >> 16) > (uint32_t)r;
                        }
                    }
                }
                const uint32_t bal = __ballot_sync(kFull, vis);
                if (lane == 0) visit()[k0 >> 5] = bal;
            }
        }
        __syncwarp();
        // ---- the sweep through the live spans
        {
            int r = 0;
            uint32_t pnext = P()[0];
            while (r < m && !overflow) {
                if (!live_any()) {
                    // jump to the next rank where a class can be born
                    int wq = r >> 5;
                    uint32_t bits = visit()[wq] & (kFull << (r & 31));
                    const int nwords = epad() >> 5;
                    while (!bits && ++wq < nwords) bits = visit()[wq];
                    if (!bits) break;
                    const int rn = 32 * wq + __ffs(bits) - 1;
                    if (rn >= m) break;
                    if (rn != r) { r = rn; pnext = P()[r]; }
                }
                const uint32_t pij = pnext;
                pnext = P()[r + 1];   // r + 1 <= epad() - 1 or the padding entry: always readable
                if (pij & (kTie | kTiePrev)) {
                    int r0 = r;
                    while (r0 > 0 && (P()[r0 - 1] & kTie)) --r0;
                    int r1 = r;
                    while (P()[r1] & kTie) ++r1;
                    ++r1;
                    tie_run(r0, r1);
                    r = r1;
                    pnext = P()[r < epad() ? r : epad() - 1];
                } else {
                    if (!(pij & kMst)) single_edge(r, pij);
                    ++r;
                }
            }
        }
        if (!overflow) {
            // cycles still alive at thresh are essential
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                uint32_t bits = live[c];
                while (bits) {
                    const int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int br = __shfl_sync(kFull, brank[c], s);
                    if (n1 < rcap()) {
                        if (lane == 0) {
                            rec()[n1] = (uint32_t)br;
                            rec()[rcap() + n1] = kEssential;
                            rec()[2 * rcap() + n1] = kEssential;
                        }
                        ++n1;
                    } else overflow = true;
                }
            }
        }
        if (overflow) {
            if (lane == 0) p.overflow_list[atomicAdd(p.n_overflow, 1)] = b;
            return;
        }
        __syncwarp();
        // ---- H0 essentials: eldest vertex of every surviving component, ascending
        {
            int base0 = n0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int v = lane + 32 * h;
                const bool is = v < n() && eld()[comp()[v]] == v;
                const uint32_t bal = __ballot_sync(kFull, is);
                if (is) {
                    const size_t o = ((size_t)b * n() + base0 + __popc(bal & lanemask_lt())) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = v; p.pr0[o + 1] = -1; }
                }
                base0 += __popc(bal);
            }
            n0 = base0;
        }
        // ---- H1 rows in ripser's order: descending birth rank
        int st = nan_seen ? TDA_ST_NAN_INPUT : 0;
        for (int k = lane; k < n1; k += 32) {
            const uint32_t br = rec()[k];
            int pos = 0;
            for (int t = 0; t < n1; ++t) pos += rec()[t] > br;
            if (pos < p.cap1) {
                const size_t o = ((size_t)b * p.cap1 + pos) * 2;
                const uint32_t dr = rec()[rcap() + k], tr = rec()[2 * rcap() + k];
                const uint32_t pb = P()[br];
                p.bd1[o] = dist(p_i(pb), p_j(pb));
                float dth = __int_as_float(0x7F800000);
                if (tr != kEssential) { const uint32_t pd = P()[dr]; dth = dist(p_i(pd), p_j(pd)); }
                p.bd1[o + 1] = dth;
                if (p.pr1) {
                    p.pr1[o] = c2(p_i(pb)) + (p_j(pb));
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

template <int CPL, int NT>
__global__ void __launch_bounds__(256, 1) rips_bits_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef BitLayout<CPL, NT> L;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    const int N = NT > 0 ? NT : p.N;
    BitWarp<CPL, NT> s;
    s.base = smem_raw + (size_t)wib * L::bytes(N);
    s.lane = lane;
    s.Nrt = N;
    s.ld = p.ld;
    const int total = p.worklist ? *p.n_work : p.B;
    for (int t = gw; t < total; t += nw) {
        const int b = p.worklist ? p.worklist[t] : t;
        s.run(p, b);
        __syncwarp();
    }
}

template <int CPL, int NT>
static cudaError_t launch_one(const Params& p, int sms, cudaStream_t st, bool whole_batch) {
    typedef BitLayout<CPL, NT> L;
    const size_t per_warp = L::bytes(p.N);
    // two CTAs per SM, each with as many warps as shared memory allows (<= 8)
    int wpb = (int)(((227 * 1024) / 2 - 1024) / per_warp);
    if (wpb < 1) wpb = 1;
    if (wpb > 8) wpb = 8;
    int per_sm = (int)((227 * 1024) / (per_warp * wpb + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    long long grid = (long long)sms * per_sm;
    if (whole_batch) {
        const long long need = ((long long)p.B + wpb - 1) / wpb;
        if (need < grid) grid = need;
    }
    const size_t smem = per_warp * wpb;
    cudaError_t e = cudaFuncSetAttribute(rips_bits_kernel<CPL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ProfScope prof(CPL == 1 ? "rips_small_w2" : "rips_small_w4", st);   // tier names kept for the bench
    rips_bits_kernel<CPL, NT><<<(int)grid, wpb * 32, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_bits_tier(const Params& p, int cpl, int sms, cudaStream_t st) {
    if (cpl == 1) {
        return p.N == 47 ? launch_one<1, 47>(p, sms, st, true) : launch_one<1, 0>(p, sms, st, true);
    }
    return launch_one<2, 0>(p, sms, st, false);
}

size_t bits_tier_warp_bytes(int N, int cpl) {
    if (cpl == 1) return N == 47 ? BitLayout<1, 47>::bytes(N) : BitLayout<1, 0>::bytes(N);
    return BitLayout<2, 0>::bytes(N);
}

}  // namespace rips_small
}  // namespace tda
