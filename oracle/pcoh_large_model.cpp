// TEST INFRASTRUCTURE ONLY — executable CPU model of the algorithm that
// tda_eeg_audio_b200/csrc/rips_large.cu runs for big clouds (64 < N <= 2048).  Never linked into
// or called from the product path; tests/ check it against oracle/rips_cpu.cpp (the restatement
// of ripser, /root/reference/scripts/utils.py:131) so that the *algorithm* is validated on CPU
// before the kernel is compared with the oracle on the GPU.
//
// Algorithm = persistent cohomology by cocycle annotation (see oracle/pcoh_model.py), organised
// so that almost everything is data-parallel and the serial sweep only touches edges that a live
// cocycle can see:
//   1. edges sorted by (length asc, index desc) -> rank r; rank matrix T[i][j] = r (INF if absent)
//   2. Kruskal -> MST flags + H0 pairs
//   3. classification of every non-MST edge e (parallel in the kernel), run-aware:
//        top cofacet = largest apex v with T[i][v], T[j][v] < r1 (r1 = end of e's tie run);
//        apparent (zero-persistence pair, defv = v) iff that triangle has e as youngest edge,
//        otherwise (or no cofacet) e gives BIRTH to a class
//   4. sweep over tie runs [r0, r1) (a single edge is a run of one).  S[v] = OR of PHI over the
//      edges at v (support summary).  A run edge is ACTIVE iff it is a birth or
//      (S[i] | S[j]) & live != 0; inactive edges have PHI = 0 and all coboundaries 0 -> skipped.
//        A. in rank order: births take a slot (PHI = unit), active apparent edges get
//           PHI[e] := PHI[i,defv] ^ PHI[j,defv]
//        B. death loop: over the triangles (e, z), z in G_e = {z : T[i][z], T[j][z] < rank(e)} of
//           the active edges find the one with the LARGEST index whose coboundary mask
//           c = PHI[e]^PHI[i,z]^PHI[j,z] (live bits) is non-zero; its youngest class dies there, the
//           other classes in c absorb it (PHI[q] ^= c wherever the dying bit is set); repeat.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {
typedef int64_t idx_t;
inline idx_t c2(idx_t i) { return i * (i - 1) / 2; }
inline idx_t c3(idx_t i) { return i * (i - 1) * (i - 2) / 6; }
inline idx_t tri_index(int x, int y, int z) {
    int a = std::max(x, std::max(y, z)), c = std::min(x, std::min(y, z)), b = x + y + z - a - c;
    return c3(a) + c2(b) + c;
}
constexpr int WW = 16;  // 64-bit words per mask -> 1024 simultaneous classes in the model
struct Mask {
    uint64_t w[WW];
    Mask() { memset(w, 0, sizeof w); }
    bool any() const { uint64_t a = 0; for (int k = 0; k < WW; ++k) a |= w[k]; return a != 0; }
    Mask operator&(const Mask& o) const { Mask r; for (int k = 0; k < WW; ++k) r.w[k] = w[k] & o.w[k]; return r; }
    Mask operator^(const Mask& o) const { Mask r; for (int k = 0; k < WW; ++k) r.w[k] = w[k] ^ o.w[k]; return r; }
    Mask& operator|=(const Mask& o) { for (int k = 0; k < WW; ++k) w[k] |= o.w[k]; return *this; }
    Mask& operator^=(const Mask& o) { for (int k = 0; k < WW; ++k) w[k] ^= o.w[k]; return *this; }
    Mask& operator&=(const Mask& o) { for (int k = 0; k < WW; ++k) w[k] &= o.w[k]; return *this; }
    bool test(int s) const { return (w[s >> 6] >> (s & 63)) & 1; }
    void set(int s) { w[s >> 6] |= 1ull << (s & 63); }
    void clr(int s) { w[s >> 6] &= ~(1ull << (s & 63)); }
    int count() const { int c = 0; for (int k = 0; k < WW; ++k) c += __builtin_popcountll(w[k]); return c; }
};
struct Edge { float d; idx_t idx; int i, j; };
}  // namespace

extern "C" int model_large_rips_h01(const float* D, int n, int ld, float thresh, float* bd0, int64_t* pr0,
                                    float* bd1, int64_t* pr1, int cap1, int* counts, long long* stats) {
    const int INF = std::numeric_limits<int>::max();
    std::vector<Edge> E;
    for (int i = 1; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            float v = D[(size_t)j * ld + i] + 0.0f;
            if (v <= thresh) E.push_back(Edge{v, c2(i) + j, i, j});
        }
    std::sort(E.begin(), E.end(), [](const Edge& a, const Edge& b) { return a.d < b.d || (a.d == b.d && a.idx > b.idx); });
    const int m = (int)E.size();
    std::vector<int> T((size_t)n * n, INF);
    for (int r = 0; r < m; ++r) { T[(size_t)E[r].i * n + E[r].j] = r; T[(size_t)E[r].j * n + E[r].i] = r; }
    // ---- Kruskal
    std::vector<int> comp(n), eld(n);
    for (int v = 0; v < n; ++v) comp[v] = eld[v] = v;
    std::vector<uint8_t> mst(m, 0);
    int n0 = 0;
    for (int r = 0; r < m; ++r) {
        int a = comp[E[r].i], b = comp[E[r].j];
        if (a == b) continue;
        mst[r] = 1;
        int ea = eld[a], eb = eld[b];
        if (E[r].d != 0.0f) {
            bd0[2 * n0] = 0; bd0[2 * n0 + 1] = E[r].d; pr0[2 * n0] = std::min(ea, eb); pr0[2 * n0 + 1] = E[r].idx; ++n0;
        }
        for (int v = 0; v < n; ++v) if (comp[v] == a) comp[v] = b;
        eld[b] = std::max(ea, eb);
    }
    {
        std::vector<int> ess;
        for (int v = 0; v < n; ++v) if (eld[comp[v]] == v) ess.push_back(v);
        std::sort(ess.begin(), ess.end());
        for (int v : ess) { bd0[2 * n0] = 0; bd0[2 * n0 + 1] = std::numeric_limits<float>::infinity(); pr0[2 * n0] = v; pr0[2 * n0 + 1] = -1; ++n0; }
    }
    counts[0] = n0;
    // ---- run ends and classification
    std::vector<int> run_end(m);
    for (int r = m - 1; r >= 0; --r) run_end[r] = (r + 1 < m && E[r + 1].d == E[r].d) ? run_end[r + 1] : r + 1;
    std::vector<int> defv(m, -1);  // -1 = birth (for non-MST edges)
    for (int r = 0; r < m; ++r) {
        if (mst[r]) continue;
        const int r1 = run_end[r];
        const int* Ti = &T[(size_t)E[r].i * n];
        const int* Tj = &T[(size_t)E[r].j * n];
        for (int v = n - 1; v >= 0; --v)
            if (Ti[v] < r1 && Tj[v] < r1) {
                if (Ti[v] < r && Tj[v] < r) defv[r] = v;
                break;
            }
    }
    // ---- sweep
    std::vector<Mask> PHI(m), S(n);
    Mask live, used;
    std::vector<int> brank(64 * WW, -1);
    struct Rec { int birth_rank, death_rank; idx_t tri; };
    std::vector<Rec> recs;
    long long st_runs = 0, st_active = 0, st_deaths = 0, st_absorb = 0, st_maxlive = 0, st_births = 0, st_loop = 0,
              st_scrubs = 0, st_active_edges_tested = 0;
    bool overflow = false;
    int r0 = 0;
    while (r0 < m && !overflow) {
        const int r1 = run_end[r0];
        // quick skip test (what the kernel's scan does)
        bool any_flag = false;
        for (int r = r0; r < r1 && !any_flag; ++r) {
            if (mst[r]) continue;
            if (defv[r] < 0) any_flag = true;
            else { Mask s = S[E[r].i]; s |= S[E[r].j]; if ((s & live).any()) any_flag = true; }
        }
        if (!any_flag) { r0 = r1; continue; }
        ++st_runs;
        std::vector<int> active;
        // A. births and definitions, rank order
        for (int r = r0; r < r1 && !overflow; ++r) {
            if (mst[r]) continue;
            const int x = E[r].i, y = E[r].j;
            if (defv[r] < 0) {
                int s = -1;
                for (int attempt = 0; attempt < 2 && s < 0; ++attempt) {
                    for (int q = 0; q < 64 * WW; ++q) if (!used.test(q)) { s = q; break; }
                    if (s < 0) {
                        ++st_scrubs;
                        for (int q = 0; q < m; ++q) PHI[q] &= live;
                        for (int v = 0; v < n; ++v) S[v] &= live;
                        used = live;
                    }
                }
                if (s < 0) { overflow = true; break; }
                used.set(s); live.set(s);
                brank[s] = r;
                PHI[r] = Mask(); PHI[r].set(s);
                S[x].set(s); S[y].set(s);
                active.push_back(r);
                ++st_births;
                st_maxlive = std::max<long long>(st_maxlive, live.count());
            } else {
                Mask s = S[x]; s |= S[y];
                if (!(s & live).any()) continue;
                const int vt = defv[r];
                PHI[r] = (PHI[T[(size_t)x * n + vt]] ^ PHI[T[(size_t)y * n + vt]]) & live;
                S[x] |= PHI[r]; S[y] |= PHI[r];
                active.push_back(r);
            }
        }
        st_active += (long long)active.size();
        // B. death loop
        while (!overflow) {
            ++st_loop;
            idx_t best = -1; Mask bc; int brun = -1;
            for (int r : active) {
                const int x = E[r].i, y = E[r].j;
                const int* Tx = &T[(size_t)x * n];
                const int* Ty = &T[(size_t)y * n];
                ++st_active_edges_tested;
                for (int z = 0; z < n; ++z) {
                    if (Tx[z] < r && Ty[z] < r) {
                        Mask c = (PHI[r] ^ PHI[Tx[z]] ^ PHI[Ty[z]]) & live;
                        if (c.any()) {
                            idx_t t = tri_index(x, y, z);
                            if (t > best) { best = t; bc = c; brun = r; }
                        }
                    }
                }
            }
            if (best < 0) break;
            ++st_deaths;
            int slot = -1, age = -1;
            for (int q = 0; q < 64 * WW; ++q) if (bc.test(q) && brank[q] > age) { age = brank[q]; slot = q; }
            if (age < r0) recs.push_back(Rec{age, brun, best});
            live.clr(slot);
            Mask rest = bc; rest.clr(slot);
            if (rest.any()) {
                ++st_absorb;
                for (int q = 0; q < r1; ++q)
                    if (PHI[q].test(slot)) { PHI[q] ^= bc; S[E[q].i] |= bc; S[E[q].j] |= bc; }
            }
        }
        r0 = r1;
    }
    if (overflow) return -2;
    for (int q = 0; q < 64 * WW; ++q) if (live.test(q)) recs.push_back(Rec{brank[q], -1, -1});
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.birth_rank > b.birth_rank; });
    int n1 = 0;
    for (const Rec& rc : recs) {
        if (n1 < cap1) {
            bd1[2 * n1] = E[rc.birth_rank].d;
            bd1[2 * n1 + 1] = rc.death_rank < 0 ? std::numeric_limits<float>::infinity() : E[rc.death_rank].d;
            pr1[2 * n1] = E[rc.birth_rank].idx;
            pr1[2 * n1 + 1] = rc.tri;
        }
        ++n1;
    }
    counts[1] = n1;
    if (stats) {
        stats[0] = m; stats[1] = st_runs; stats[2] = st_active; stats[3] = st_births; stats[4] = st_deaths;
        stats[5] = st_absorb; stats[6] = st_maxlive; stats[7] = st_loop; stats[8] = st_scrubs;
        stats[9] = st_active_edges_tested;
    }
    return 0;
}
