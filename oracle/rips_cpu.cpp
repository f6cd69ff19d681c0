// TEST INFRASTRUCTURE ONLY — CPU oracle / CPU baseline.  Never linked into, imported by or
// called from the product path (tda_eeg_audio_b200/); only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load liboracle_rips.so.
//
// parity status: "parity unpinned".  The reference's Rips arithmetic lives in the third-party
// package `ripser>=0.6` (/root/reference/requirements.txt:5; call sites
// /root/reference/scripts/utils.py:131,140 and
// /root/reference/scripts/tda_eeg_classification_v2.py:170-175), which is not in this image and
// for which the reference holds no golden vectors.  This file restates Ripser's
// published algorithm and is pinned instead against the definition-level
// boundary-matrix reduction in oracle/rips_naive.py (tests/test_oracle_rips.py).
//
// Algorithm (persistent cohomology, Z/2, dims 0 and 1), float32 values, int64 indices:
//   * edges (i>j) with d <= thresh, index C(i,2)+j, sorted by (d ascending, index descending);
//   * H0 by union-find over the sorted edges (elder rule recorded for the vertex of the pair);
//   * the non-merging edges, in reverse order, are the H1 columns; each column's coboundary
//     (triangles e+v, index C(a,3)+C(b,2)+c, diameter = max edge) is reduced against earlier
//     columns; pivot = cofacet that is earliest in the filtration (smallest diameter, then
//     largest index); "emergent pair" shortcut: the first enumerated cofacet of equal diameter
//     that nobody owns yet is the pivot; zero-persistence pairs are not emitted;
//   * clearing is implicit: merging edges are never H1 columns.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef int64_t idx_t;

struct Entry {  // a simplex in a working column
    float diam;
    idx_t idx;
};
// priority: earliest in filtration first  (smaller diam, then larger index)
struct LaterInFiltration {
    bool operator()(const Entry& a, const Entry& b) const {
        return a.diam > b.diam || (a.diam == b.diam && a.idx < b.idx);
    }
};
typedef std::priority_queue<Entry, std::vector<Entry>, LaterInFiltration> Heap;

struct Scratch {
    std::vector<float> tri;  // lower-triangular distances, row i (i>j): tri[C(i,2)+j]
    std::vector<Entry> edges;
    std::vector<int> parent, eldest;
    std::vector<Entry> columns;
    std::vector<int32_t> pivot_flat;
    std::unordered_map<idx_t, int32_t> pivot_map;
    std::vector<std::vector<int32_t>> added;  // reduction columns: which earlier columns were added
    std::vector<Entry> cof;
    long long n_additions = 0, n_emergent = 0, n_columns = 0;
};

inline idx_t c2(idx_t i) { return i * (i - 1) / 2; }
inline idx_t c3(idx_t i) { return i * (i - 1) * (i - 2) / 6; }

struct Engine {
    int n;
    float thresh;
    Scratch& s;
    bool use_flat;
    Engine(int n_, float t, Scratch& s_) : n(n_), thresh(t), s(s_) {
        use_flat = c3(n) <= (idx_t)1 << 24;
    }
    inline float dist(int a, int b) const { return a > b ? s.tri[c2(a) + b] : s.tri[c2(b) + a]; }
    inline void edge_vertices(idx_t e, int& i, int& j) const {
        // largest i with C(i,2) <= e
        idx_t ii = (idx_t)((1.0 + std::sqrt(1.0 + 8.0 * (double)e)) / 2.0);
        while (c2(ii) > e) --ii;
        while (c2(ii + 1) <= e) ++ii;
        i = (int)ii;
        j = (int)(e - c2(ii));
    }
    int find(int x) {
        while (s.parent[x] != x) {
            s.parent[x] = s.parent[s.parent[x]];
            x = s.parent[x];
        }
        return x;
    }
    int32_t owner(idx_t t) const {
        if (use_flat) return s.pivot_flat[t];
        auto it = s.pivot_map.find(t);
        return it == s.pivot_map.end() ? -1 : it->second;
    }
    void set_owner(idx_t t, int32_t c) {
        if (use_flat) s.pivot_flat[t] = c; else s.pivot_map[t] = c;
    }
    // enumerate cofacets of edge e (diameter de) in DESCENDING index order
    template <class F> void cofacets(idx_t e, float de, F&& f) const {
        int i, j;
        edge_vertices(e, i, j);
        for (int v = n - 1; v >= 0; --v) {
            if (v == i || v == j) continue;
            float a = dist(v, i), b = dist(v, j);
            if (!(a <= thresh) || !(b <= thresh)) continue;
            float dm = std::max(de, std::max(a, b));
            idx_t t;
            if (v > i) t = c3(v) + c2(i) + j;
            else if (v > j) t = c3(i) + c2(v) + j;
            else t = c3(i) + c2(j) + v;
            if (f(Entry{dm, t})) return;
        }
    }
    static bool pop_pivot(Heap& h, Entry& out) {
        while (!h.empty()) {
            Entry p = h.top();
            h.pop();
            if (!h.empty() && h.top().idx == p.idx) { h.pop(); continue; }  // cancels mod 2
            out = p;
            return true;
        }
        return false;
    }
};

void run_one(const float* D, int n, int ld, float thresh, Scratch& s,
             float* bd0, idx_t* pr0, float* bd1, idx_t* pr1, int cap1, int* counts, int* status) {
    Engine g(n, thresh, s);
    const float INF = std::numeric_limits<float>::infinity();
    s.tri.resize((size_t)c2(n));
    for (int i = 1; i < n; ++i)
        for (int j = 0; j < i; ++j) s.tri[c2(i) + j] = D[(size_t)j * ld + i];  // upper triangle
    s.edges.clear();
    for (int i = 1; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            float v = s.tri[c2(i) + j];
            if (v <= thresh) s.edges.push_back(Entry{v, c2(i) + j});
        }
    std::sort(s.edges.begin(), s.edges.end(), [](const Entry& a, const Entry& b) {
        return a.diam < b.diam || (a.diam == b.diam && a.idx > b.idx);
    });
    // ---- H0 ----
    s.parent.resize(n);
    s.eldest.resize(n);
    for (int v = 0; v < n; ++v) s.parent[v] = v, s.eldest[v] = v;
    s.columns.clear();
    int n0 = 0;
    for (const Entry& e : s.edges) {
        int i, j;
        g.edge_vertices(e.idx, i, j);
        int a = g.find(i), b = g.find(j);
        if (a == b) { s.columns.push_back(e); continue; }
        int ea = s.eldest[a], eb = s.eldest[b];
        int dying = std::min(ea, eb);  // larger vertex index == earlier in the filtration == elder
        if (e.diam != 0.0f) {
            bd0[2 * n0] = 0.0f; bd0[2 * n0 + 1] = e.diam;
            pr0[2 * n0] = dying; pr0[2 * n0 + 1] = e.idx;
            ++n0;
        }
        s.parent[a] = b;
        s.eldest[b] = std::max(ea, eb);
    }
    // essential classes, ascending eldest vertex
    {
        std::vector<int> ess;
        for (int v = 0; v < n; ++v) if (g.find(v) == v) ess.push_back(s.eldest[v]);
        std::sort(ess.begin(), ess.end());
        for (int v : ess) {
            bd0[2 * n0] = 0.0f; bd0[2 * n0 + 1] = INF;
            pr0[2 * n0] = v; pr0[2 * n0 + 1] = -1;
            ++n0;
        }
    }
    counts[0] = n0;
    // ---- H1 ----
    std::reverse(s.columns.begin(), s.columns.end());
    const int ncol = (int)s.columns.size();
    if (g.use_flat) s.pivot_flat.assign((size_t)c3(n) + 1, -1); else s.pivot_map.clear();
    if ((int)s.added.size() < ncol) s.added.resize(ncol);
    int n1 = 0;
    int st = 0;
    s.n_columns += ncol;
    for (int c = 0; c < ncol; ++c) {
        const Entry col = s.columns[c];
        s.added[c].clear();
        Heap work;
        Entry pivot{0, -1};
        bool have_pivot = false, emergent_open = true;
        // initial coboundary with emergent-pair shortcut
        s.cof.clear();
        g.cofacets(col.idx, col.diam, [&](const Entry& t) {
            s.cof.push_back(t);
            if (emergent_open && t.diam == col.diam) {
                if (g.owner(t.idx) < 0) { pivot = t; have_pivot = true; return true; }
                emergent_open = false;
            }
            return false;
        });
        if (have_pivot) {
            ++s.n_emergent;
        } else {
            for (const Entry& t : s.cof) work.push(t);
            have_pivot = Engine::pop_pivot(work, pivot);
            while (have_pivot) {
                int32_t o = g.owner(pivot.idx);
                if (o < 0) break;
                // add column o: its own coboundary plus those of everything it absorbed
                work.push(pivot);  // put the pivot back, the added column cancels it
                auto add_cob = [&](const Entry& e) {
                    g.cofacets(e.idx, e.diam, [&](const Entry& t) { work.push(t); return false; });
                };
                add_cob(s.columns[o]);
                s.added[c].push_back(o);
                ++s.n_additions;
                for (int32_t q : s.added[o]) { add_cob(s.columns[q]); s.added[c].push_back(q); ++s.n_additions; }
                have_pivot = Engine::pop_pivot(work, pivot);
            }
            // reduce the stored list mod 2 (an index appearing twice cancels)
            if (!s.added[c].empty()) {
                std::sort(s.added[c].begin(), s.added[c].end());
                std::vector<int32_t> keep;
                for (size_t k = 0; k < s.added[c].size();) {
                    size_t m = k;
                    while (m < s.added[c].size() && s.added[c][m] == s.added[c][k]) ++m;
                    if ((m - k) & 1) keep.push_back(s.added[c][k]);
                    k = m;
                }
                s.added[c].swap(keep);
            }
        }
        if (have_pivot) {
            g.set_owner(pivot.idx, c);
            if (pivot.diam > col.diam) {
                if (n1 < cap1) {
                    bd1[2 * n1] = col.diam; bd1[2 * n1 + 1] = pivot.diam;
                    pr1[2 * n1] = col.idx; pr1[2 * n1 + 1] = pivot.idx;
                } else st = 1;
                ++n1;
            }
        } else {
            if (n1 < cap1) {
                bd1[2 * n1] = col.diam; bd1[2 * n1 + 1] = INF;
                pr1[2 * n1] = col.idx; pr1[2 * n1 + 1] = -1;
            } else st = 1;
            ++n1;
        }
    }
    counts[1] = n1;
    if (status) *status = st;
}

}  // namespace

extern "C" {

// D: (B, n, ld) float32 row-major, only [i][j], i<j is read.  Outputs are padded per item:
// bd0 (B, n, 2) f32, pr0 (B, n, 2) i64, bd1 (B, cap1, 2) f32, pr1 (B, cap1, 2) i64,
// counts (B, 2) i32 (true counts even when H1 overflows cap1), status (B) i32 (1 = overflow).
// stats (optional, 3 x int64): columns, emergent pairs, column additions.
int oracle_rips_h01_batched(const float* D, int B, int n, int ld, float thresh, float* bd0, int64_t* pr0,
                            float* bd1, int64_t* pr1, int* counts, int cap1, int* status, int nthreads,
                            long long* stats) {
    if (B < 0 || n < 1 || ld < n || cap1 < 0) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    long long tot[3] = {0, 0, 0};
#pragma omp parallel num_threads(nthreads)
    {
        Scratch s;
#pragma omp for schedule(dynamic, 16)
        for (int b = 0; b < B; ++b) {
            run_one(D + (size_t)b * n * ld, n, ld, thresh, s, bd0 + (size_t)b * n * 2, pr0 + (size_t)b * n * 2,
                    bd1 + (size_t)b * cap1 * 2, pr1 + (size_t)b * cap1 * 2, cap1, counts + 2 * (size_t)b,
                    status ? status + b : nullptr);
        }
#pragma omp critical
        {
            tot[0] += s.n_columns; tot[1] += s.n_emergent; tot[2] += s.n_additions;
        }
    }
    if (stats) { stats[0] = tot[0]; stats[1] = tot[1]; stats[2] = tot[2]; }
    return 0;
}

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
