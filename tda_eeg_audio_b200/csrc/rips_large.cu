// rips_large.cu — Vietoris–Rips H0+H1 (Z/2) for batches of big clouds (N <= 2048): the
// scaling-stress Takens clouds of 1,000-2,000 points (BASELINE.json configs[4]) and, because the
// algorithm does far less serial work than a CTA-per-cloud sweep over every edge, any batch
// of clouds above the 64-point engine.
//
// Replaces ripser.ripser(point_cloud / dm, maxdim=1, thresh) as called by
//   /root/reference/scripts/utils.py:123-132 (compute_audio_persistence)
//
// Not a port of Ripser.  Ripser walks the 2M columns of a 2,000-point cloud one after the other
// (coboundary enumeration + apparent/emergent pair test per column, heap reduction for the rest).
// Here (oracle/pcoh_large_model.cpp is the executable statement of the algorithm, validated on
// CPU) the work is split into grid-wide data-parallel phases over ALL clouds of a chunk and a
// short serial sweep per cloud:
//   K1 rank      the order-preserving integer image of every f32 edge length, laid out in descending edge index,
//                then a stable LSD radix sort of (key, (i, j)) -- 8-bit digits, lanes with the same digit found by
//                ballots -- whose stability yields Ripser's tie-break (equal length => larger index first); sorted
//                position r -> P[r] = (i, j, tie flag), rank matrix T[i][j] = r.  Up to 256 points (the audio
//                path): one CTA per cloud, the whole array in shared memory, in-place passes through registers,
//                16-bit ranks.  Above: a grid-wide sort of a group of clouds at a time, as many as keep their
//                sort arrays and rank matrices in L2 (keys / per-tile histograms / staged stable scatter / finish).
//   K3 kruskal   MST flags + the H0 pairs (elder rule for the vertex): up to 256 points inside K4's walk of the
//                edge list, above one CTA per cloud
//   K4 classify  the first cofacet of every non-MST edge (largest apex v with T[i][v], T[j][v] inside the edge's
//                tie run or before it).  If that triangle has the edge as its youngest edge the two form an
//                apparent zero-persistence pair (defv = v) -- >99 % of all columns end here -- otherwise the edge
//                gives BIRTH to an H1 class.  Up to 256 points: one warp per cloud, 32 ranks per step, the
//                adjacency of the earlier ranks as bit rows in shared memory.  Above: one thread per edge over the
//                whole grid walking two rank rows, the early edges through adjacency bit rows at six early ranks.
//   K5 sweep     one CTA per cloud, persistent cohomology by cocycle annotation restricted to the
//                edges a live cocycle can see: S[v] = OR of the cocycle masks over the edges at v;
//                an edge with (S[i] | S[j]) & live == 0 has PHI = 0 and coboundary 0 on all of
//                its triangles and is skipped by a scan of the chunk's ranks.  The
//                visited tie runs (a single edge is a run of one) are handled exactly:
//                  A. rank order: births take a slot, visible apparent edges get
//                     PHI[e] := PHI[i,defv] ^ PHI[j,defv]   (non-zero values live in a compact
//                     store addressed through Q[i][j], a u16 matrix read by rows like T)
//                  B. death loop, thread = apex: among the triangles (e, z), T[i][z], T[j][z] <
//                     rank(e), of the visited edges the one with the LARGEST index and non-zero
//                     coboundary mask kills its youngest class; the other classes of the mask
//                     absorb it (PHI ^= mask wherever the dying bit is set); repeat.
//   Capacity tiers: 64 simultaneous classes first for clouds up to 256 points (W = 2 words), 256
//   (W = 8; 512 above 1,024 points), then 1,024 (W = 32), through device-side hand-over lists.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
namespace rips_large {

constexpr int kMaxN = 2048;
constexpr int kSmCount = 148;   // B200
constexpr uint32_t kInf = 0xFFFFFFFFu;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kEssential = 0xFFFFFFFFu;
// P[r] = j | i << 11 | flags
constexpr uint32_t kMst = 1u << 22;      // merging edge (H0 death)
constexpr uint32_t kTieNext = 1u << 23;  // the next edge in the order has the same length
constexpr uint32_t kBirth = 1u << 24;    // gives birth to an H1 class
constexpr int kCapPMax = 65534;          // compact PHI entries per cloud (Q is u16, 0 = none)
constexpr int kCapRMax = 65536;          // death records per cloud
constexpr int kLevels = 6;               // early adjacency snapshots of a big cloud: after m/128, m/64, ... m/4 edges
constexpr int kLevWords = 2048 / 32;     // words per bit row (kMaxN vertices)

__host__ __device__ inline long long c2(long long i) { return i * (i - 1) / 2; }
__host__ __device__ inline long long c3(long long i) { return i * (i - 1) * (i - 2) / 6; }
__device__ __forceinline__ int p_i(uint32_t p) { return (p >> 11) & 2047; }
__device__ __forceinline__ int p_j(uint32_t p) { return p & 2047; }
__device__ __forceinline__ uint32_t float_key(float d) {
    uint32_t u = __float_as_uint(d);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
// (32-bit throughout: C(a,2) (a - 2) < 2^32 for a <= 2047 and is divisible by 3.  For a fixed edge (x, y) the
// index grows with the apex z, so "largest index" = "largest apex" and one evaluation per edge is enough)
__device__ __forceinline__ uint32_t tri_index(int x, int y, int z) {
    const uint32_t a = (uint32_t)max(x, max(y, z)), c = (uint32_t)min(x, min(y, z)), b = (uint32_t)(x + y + z) - a - c;
    const uint32_t a2 = a * (a - 1u) / 2u;
    return a2 * (a - 2u) / 3u + b * (b - 1u) / 2u + c;
}

struct Params {
    const float* D;
    const int* npts;
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
    // chunk of clouds [c0, c0 + C)
    int c0, C;
    int ldT, ib;
    long long Emax;
    uint32_t* sortbuf; // [C][4][Emax]: keys / payloads, ping and pong
    uint32_t* skey;   // [C][Emax] sorted keys (classification of tie runs)
    uint32_t* P;      // [C][Emax]
    void* T;          // [C][N][ldT] rank matrix, uint16_t (N <= 256) or uint32_t; all ones = edge absent
    uint16_t* Q;      // [C][N][ldT]
    uint16_t* defv;   // [C][Emax]
    int* m;           // [C] number of edges <= thresh
    int* nanflag;     // [C]
    // per-CTA sweep scratch
    int capP, capR;   // compact PHI entries / death records per cloud
    uint32_t* phic;   // [grid][capP][W]
    uint32_t* pcr;    // [grid][capP] rank of the edge of a compact entry
    uint32_t* act;    // [grid][Emax]
    uint32_t* rec;    // [grid][3][capR]
    uint32_t* sglob;  // [grid][N][W] (tiers that keep S in global memory)
    const int* worklist;
    const int* n_work;
    int* overflow_list;
    int* n_overflow;
    // sweep CTAs take their clouds from a device-side queue: a cloud's sweep is serial and its length varies
    // several-fold between clouds, so a static assignment leaves most SMs idle behind the heaviest one
    int* queue;         // next position of the current sweep launch
    const int* order;   // position -> cloud, heaviest first (nullptr: positions are clouds)
    int* nbirth;        // [C] number of H1 births of every cloud (classify), the weight behind `order`
    uint2* left;           // (cloud, rank) of the tie-run members classify_bits_kernel leaves to the rank-row walk
    int* n_left;
    // big clouds: adjacency bit rows at six early ranks (classification of the early edges)
    const uint32_t* lev;   // [C][kLevels][N][kLevWords]; nullptr: none
    const int* levR;       // [C][2 kLevels]: the level ranks, then the same extended to the end of their tie runs
};

__device__ __forceinline__ int cloud_n(const Params& p, int b) {
    int n = p.npts ? p.npts[b] : p.N;
    return n < 0 ? 0 : (n > p.N ? p.N : n);
}

// ------------------------------------------------------------------------------------ K1 rank
template <typename TT> struct RankOf;
template <> struct RankOf<uint16_t> { static constexpr uint32_t kAbsent = 0xFFFFu; };
template <> struct RankOf<uint32_t> { static constexpr uint32_t kAbsent = 0xFFFFFFFFu; };

// K1 for big clouds (257-2,048 points): a grid-wide sort of a GROUP of clouds at a time.  The sort arrays of a
// 2,000-point cloud are 32 MB (keys and payloads, ping and pong): a CTA per cloud on every SM streams 4.7 GB of
// them through DRAM four times, and scatters the ranks into 148 rank matrices at once, one 32-byte sector per
// 4-byte write (measured: 94 GB of DRAM traffic for 148 clouds, 3x the algorithmic bytes, long-scoreboard stall
// 54 cycles per issue).  Instead the whole grid works on as many clouds as fit the 126 MB L2 (three at 2,000
// points): keys -> four passes of (per-tile digit histograms, stable scatter of staged tiles) -> P,
// sorted keys and the rank matrix, every array of the group L2-resident from its first write to its last read.
// A tile is 8,192 elements of one cloud: every warp ranks its 256 elements (a row of 16-bit digit counters per
// warp, lanes with the same digit found by eight ballots), the tile is staged in shared memory ordered by digit
// and written out as runs (a digit's elements of one tile are contiguous in the destination).
constexpr int kRankThreads = 1024;
constexpr int kRankWarps = kRankThreads / 32;
constexpr int kRankPerThread = 8;
constexpr int kRankTile = kRankThreads * kRankPerThread;
constexpr size_t kRankSmem = (size_t)kRankTile * 8 + (size_t)kRankWarps * 256 * 2 + 2 * 256 * 4 + 64 * 4;

struct SortGroup {
    int g0, G;          // clouds [g0, g0 + G) of the chunk
    int ntiles;         // tiles per cloud (of the largest possible cloud)
    uint32_t* hist;     // [G][ntiles][256]: digit counts of every tile
};

// keys in descending edge index: blockIdx.y = cloud of the group, the CTAs of a cloud stride over the 32 x 32 blocks
// of the upper triangle.  A block is read by rows (lanes along i: 128-byte segments of D) and, transposed through
// shared memory, written by columns (lanes along j: the position E - 1 - (C(i,2) + j) is consecutive in j), so
// both sides move full sectors; lanes along i on the write side cost one sector per 4-byte store.
__global__ void __launch_bounds__(kRankThreads) keys_big_kernel(Params p, SortGroup sg) {
    __shared__ float tile[32][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = sg.g0 + blockIdx.y;
    const int b = p.c0 + c;
    const int n = cloud_n(p, b);
    const int E = (int)c2(n);
    uint32_t* kA = p.sortbuf + (size_t)c * 4 * p.Emax;
    uint32_t* pA = kA + p.Emax;
    const float* Db = p.D + (size_t)b * p.strideB;
    const int nb = (n + 31) >> 5, nblk = nb * (nb + 1) / 2;
    int valid = 0, nan_seen = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        // blk -> (bi, bj), bj <= bi: block row of the larger index, block column of the smaller
        int bi = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
        while (bi * (bi + 1) / 2 > blk) --bi;
        const int bj = blk - bi * (bi + 1) / 2;
        {
            const int j = 32 * bj + warp, i = 32 * bi + lane;
            float v = 0.0f;
            if (j < i && i < n) v = __ldg(Db + (size_t)j * p.ld + i);
            tile[warp][lane] = v;
        }
        __syncthreads();
        {
            const int i = 32 * bi + warp, j = 32 * bj + lane;
            if (j < i && i < n) {
                const float d = tile[lane][warp] + 0.0f;
                nan_seen |= (d != d);
                const bool ok = d <= p.thresh;
                const int pos = E - 1 - (i * (i - 1) / 2 + j);
                kA[pos] = ok ? float_key(d) : kInf;
                pA[pos] = (uint32_t)j | ((uint32_t)i << 11);
                valid += ok;
            }
        }
        __syncthreads();
    }
    valid = __reduce_add_sync(kFull, valid);
    nan_seen = __reduce_or_sync(kFull, nan_seen);
    if (lane == 0) {
        if (valid) atomicAdd(p.m + c, valid);
        if (nan_seen) atomicOr(p.nanflag + c, 1);
    }
}

// digit counts of one tile (blockIdx.x) of one cloud (blockIdx.y)
__global__ void __launch_bounds__(kRankThreads) hist_big_kernel(Params p, SortGroup sg, int pass) {
    __shared__ uint32_t h[256];
    const int tid = threadIdx.x;
    const int c = sg.g0 + blockIdx.y;
    const int E = (int)c2(cloud_n(p, p.c0 + c));
    const uint32_t* srcK = p.sortbuf + (size_t)c * 4 * p.Emax + ((pass & 1) ? 2 * p.Emax : 0);
    const int sh = 8 * pass;
    const int t0 = blockIdx.x * kRankTile;
    if (tid < 256) h[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRankPerThread; ++i) {
        const int k = t0 + i * kRankThreads + tid;
        if (k < E) atomicAdd(&h[(srcK[k] >> sh) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 256) sg.hist[((size_t)blockIdx.y * sg.ntiles + blockIdx.x) * 256 + tid] = h[tid];
}

// stable scatter of one tile
__global__ void __launch_bounds__(kRankThreads, 2) scatter_big_kernel(Params p, SortGroup sg, int pass) {
    constexpr int NTH = kRankThreads, NW = kRankWarps;
    extern __shared__ __align__(16) unsigned char rk_raw[];
    uint2* stage = reinterpret_cast<uint2*>(rk_raw);                                   // [kRankTile] (key, payload)
    uint16_t* wcnt = reinterpret_cast<uint16_t*>(rk_raw + (size_t)kRankTile * 8);       // [NW][256]
    uint32_t* gofs = reinterpret_cast<uint32_t*>(wcnt + NW * 256);                      // [256] global - local position
    uint32_t* spare = gofs + 256;                                                       // [256]
    uint32_t* red = spare + 256;                                                        // [64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int c = sg.g0 + blockIdx.y;
    const int E = (int)c2(cloud_n(p, p.c0 + c));
    const int t0 = blockIdx.x * kRankTile;
    if (t0 >= E) return;
    const int nT = E - t0 < kRankTile ? E - t0 : kRankTile;
    uint32_t* base = p.sortbuf + (size_t)c * 4 * p.Emax;
    const uint32_t* srcK = base + ((pass & 1) ? 2 * p.Emax : 0);
    const uint32_t* srcP = srcK + p.Emax;
    uint32_t* dstK = base + ((pass & 1) ? 0 : 2 * p.Emax);
    uint32_t* dstP = dstK + p.Emax;
    const int sh = 8 * pass;
    // ---- where this tile's digits go: the elements of all lower digits of the cloud + those of the same digit in
    //      earlier tiles.  A few threads per digit sum the tile counts (independent loads from L2, 62 k words per
    //      cloud); no scan kernel between the histogram and the scatter
    {
        uint32_t* part = reinterpret_cast<uint32_t*>(stage);   // [2][<= 4][256], before the tile is staged
        constexpr int TPD = NTH / 256;   // threads per digit
        const int d = tid & 255, q = tid >> 8;
        const uint32_t* hc = sg.hist + (size_t)blockIdx.y * sg.ntiles * 256;
        uint32_t tot = 0, pre = 0;
        for (int t = q; t < sg.ntiles; t += TPD) {
            const uint32_t v = hc[(size_t)t * 256 + d];
            tot += v;
            if (t < (int)blockIdx.x) pre += v;
        }
        part[q * 256 + d] = tot;
        part[1024 + q * 256 + d] = pre;
        __syncthreads();
        if (tid < 256) {
            tot = 0; pre = 0;
#pragma unroll
            for (int k = 0; k < TPD; ++k) { tot += part[k * 256 + tid]; pre += part[1024 + k * 256 + tid]; }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) red[warp] = incl;
            asm volatile("bar.sync 1, 256;");
            uint32_t dbase = incl - tot;
            for (int w = 0; w < warp; ++w) dbase += red[w];
            spare[tid] = dbase + pre;
        }
        __syncthreads();
    }
    // ---- this warp's 256 elements of the tile, their digits counted into its counter row
    uint32_t* wc32 = reinterpret_cast<uint32_t*>(wcnt);
    for (int q = tid; q < NW * 128; q += NTH) wc32[q] = 0;
    __syncthreads();
    uint32_t kr[kRankPerThread], pr[kRankPerThread];
    uint32_t* row32 = wc32 + warp * 128;
#pragma unroll
    for (int i = 0; i < kRankPerThread; ++i) {
        const int k = warp * (32 * kRankPerThread) + 32 * i + lane;
        const bool act = k < nT;
        kr[i] = act ? srcK[t0 + k] : 0u;
        pr[i] = act ? srcP[t0 + k] : 0u;
        if (act) {
            const uint32_t dg = (kr[i] >> sh) & 255u;
            atomicAdd(row32 + (dg >> 1), 1u << (16 * (dg & 1u)));
        }
    }
    __syncthreads();
    // ---- local slot of (digit, warp): exclusive scan in digit-major order
    if (tid < 256) {
        uint32_t run = 0;
        for (int w = 0; w < NW; ++w) { const uint32_t t = wcnt[w * 256 + tid]; wcnt[w * 256 + tid] = (uint16_t)run; run += t; }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) red[32 + warp] = incl;
        asm volatile("bar.sync 1, 256;");       // the eight scanning warps only
        uint32_t tstart = 0;
        for (int w = 0; w < warp; ++w) tstart += red[32 + w];
        tstart += incl - run;
        for (int w = 0; w < NW; ++w) wcnt[w * 256 + tid] += (uint16_t)tstart;
        gofs[tid] = spare[tid] - tstart;
    }
    __syncthreads();
    // ---- stage the tile in digit order (stable: chunks in order, lanes ranked inside a chunk)
    uint16_t* rowc = wcnt + warp * 256;
#pragma unroll
    for (int i = 0; i < kRankPerThread; ++i) {
        const int k = warp * (32 * kRankPerThread) + 32 * i + lane;
        const bool act = k < nT;
        const uint32_t dg = (kr[i] >> sh) & 255u;
        uint32_t peers = __ballot_sync(kFull, act);   // (eight ballots instead of __match_any_sync, see rank_small_kernel)
#pragma unroll
        for (int bt = 0; bt < 8; ++bt) {
            const uint32_t bal = __ballot_sync(kFull, (dg >> bt) & 1u);
            peers &= ((dg >> bt) & 1u) ? bal : ~bal;
        }
        if (act) stage[rowc[dg] + __popc(peers & lt)] = make_uint2(kr[i], pr[i]);
        __syncwarp();
        if (act && (peers & lt) == 0) rowc[dg] += (uint16_t)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // ---- write the tile out: a digit's elements are a contiguous run in the destination
    for (int lp = tid; lp < nT; lp += NTH) {
        const uint2 e = stage[lp];
        const uint32_t g = gofs[(e.x >> sh) & 255u] + lp;
        dstK[g] = e.x;
        dstP[g] = e.y;
    }
}

// sorted position -> P, sorted keys, rank matrix (every entry of a row is written here: present edges by rank,
// absent ones, the diagonal and the padding as "absent" -- full sectors, the matrices of the group stay in L2)
__global__ void __launch_bounds__(kRankThreads) finish_big_kernel(Params p, SortGroup sg) {
    const int tid = threadIdx.x;
    const int c = sg.g0 + blockIdx.y;
    const int n = cloud_n(p, p.c0 + c);
    const int E = (int)c2(n);
    const int m = n >= 2 ? p.m[c] : 0;
    const uint32_t* srcK = p.sortbuf + (size_t)c * 4 * p.Emax;   // four passes: back in the first pair of arrays
    const uint32_t* srcP = srcK + p.Emax;
    uint32_t* Pc = p.P + (size_t)c * p.Emax;
    uint32_t* sk = p.skey + (size_t)c * p.Emax;
    uint32_t* Tc = reinterpret_cast<uint32_t*>(p.T) + (size_t)c * p.N * p.ldT;
    for (int r = blockIdx.x * kRankThreads + tid; r < E; r += gridDim.x * kRankThreads) {
        const uint32_t key = srcK[r], pay = srcP[r];
        const int i = (int)(pay >> 11), j = (int)(pay & 2047u);
        const uint32_t val = r < m ? (uint32_t)r : 0xFFFFFFFFu;
        if (r < m) {
            const bool tie = r + 1 < m && srcK[r + 1] == key;
            Pc[r] = pay | (tie ? kTieNext : 0u);
            sk[r] = key;
        }
        Tc[(size_t)i * p.ldT + j] = val;
        Tc[(size_t)j * p.ldT + i] = val;
    }
    // diagonal and padding
    const int padw = p.ldT - n + 1;   // per row: the diagonal entry + the columns n .. ldT - 1
    for (int q = blockIdx.x * kRankThreads + tid; q < n * padw; q += gridDim.x * kRankThreads) {
        const int i = q / padw, k = q - i * padw;
        Tc[(size_t)i * p.ldT + (k == 0 ? i : n + k - 1)] = 0xFFFFFFFFu;
    }
}

// Adjacency bit rows of a big cloud after its first m/128, m/64, ... m/4 edges (each rank extended to the end of
// its tie run), from the rank matrix while it is still L2-resident.  The classification walks two rank rows per
// edge from the largest apex down until it finds a cofacet: late edges find one within a step or two, but an
// early edge -- the graph is sparse, there is often none -- walks the whole of both rows (8 KB each), and those
// few per cent of the edges were most of the classification's traffic.  With the rows of the first level above
// its rank an early edge looks at the handful of common neighbours instead.
__global__ void __launch_bounds__(256) levels_big_kernel(Params p, SortGroup sg, uint32_t* lev, int* levR) {
    __shared__ int R[kLevels];
    const int c = sg.g0 + blockIdx.y;
    const int n = cloud_n(p, p.c0 + c);
    const int m = n >= 2 ? p.m[c] : 0;
    const uint32_t* sk = p.skey + (size_t)c * p.Emax;
    if (threadIdx.x < kLevels) {
        int r = m >> (7 - threadIdx.x);
        const int r_lev = r;
        while (r > 0 && r < m && sk[r] == sk[r - 1]) ++r;   // to the end of the tie run that holds rank r - 1
        R[threadIdx.x] = r;
        if (blockIdx.x == 0) { levR[c * 2 * kLevels + threadIdx.x] = r_lev; levR[c * 2 * kLevels + kLevels + threadIdx.x] = r; }
    }
    __syncthreads();
    uint32_t rl[kLevels];
#pragma unroll
    for (int k = 0; k < kLevels; ++k) rl[k] = (uint32_t)R[k];
    const uint32_t* Tc = reinterpret_cast<const uint32_t*>(p.T) + (size_t)c * p.N * p.ldT;
    uint32_t* L = lev + (size_t)c * kLevels * p.N * kLevWords;
    const int nwt = p.ldT >> 5;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * kLevWords; idx += gridDim.x * blockDim.x) {
        const int v = idx / kLevWords, w = idx - v * kLevWords;
        uint32_t bits[kLevels];
#pragma unroll
        for (int k = 0; k < kLevels; ++k) bits[k] = 0;
        if (w < nwt) {
            const uint4* t4 = reinterpret_cast<const uint4*>(Tc + (size_t)v * p.ldT + 32 * w);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 t = t4[q];
                const uint32_t tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int k = 0; k < kLevels; ++k) bits[k] |= (uint32_t)(tt[e] < rl[k]) << (4 * q + e);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kLevels; ++k) L[(size_t)(k * p.N + v) * kLevWords + w] = bits[k];
    }
}

// K1 for clouds up to 256 points (the audio path: Takens clouds of 97-248 points): the (key, (i, j)) array
// of a cloud -- at most 32,640 elements of 32 + 16 bits -- stays in shared memory for the whole sort.  A pass is
// in place: every warp owns a contiguous segment, its lanes hold the segment's elements in registers, and once
// everybody has read (barrier) the elements are scattered to their stable positions (digit base + the counts of
// the earlier warps + the earlier chunks of the own warp + the lower lanes with the same digit); the 16-bit
// payloads follow in a second round through the saved positions, so that never more than EPT + EPT / 2 data
// registers are live.  Nothing but D is read from and nothing but P, the sorted keys, T and Q is written to
// global memory; T is assembled in the key array's space and leaves as 16-byte rows.
// digits of a pass: 256 (8 bits), or -- NINE, where the shared memory allows -- 512 when only 27 key bits vary: three
// passes of 9 bits instead of four of 8.  Counter rows of kDig + 2 u16: an odd number of words, rows start in different banks
template <int NTH, int EPT, bool NINE> struct RankSmall {
    static constexpr int NW = NTH / 32;
    static constexpr int kCap = NTH * EPT;
    static constexpr int kDigMax = NINE ? 512 : 256;
    static constexpr int kCntStride = kDigMax + 2;
    static constexpr size_t kKeyBytes = (size_t)kCap * 4;
    static constexpr size_t kSmem = kKeyBytes + (size_t)kCap * 2 + (size_t)NW * kCntStride * 2 + kDigMax * 2 + kDigMax * 2 + 4 * NW * 4 + 64;
};

template <int NTH, int EPT, int MINB, bool NINE>
__global__ void __launch_bounds__(NTH, MINB) rank_small_kernel(Params p) {
    using RS = RankSmall<NTH, EPT, NINE>;
    constexpr int NW = RS::NW;
    constexpr int kDigMax = RS::kDigMax, kCntStride = RS::kCntStride;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    uint32_t* K = reinterpret_cast<uint32_t*>(rs_raw);                                  // [kCap] keys; later T
    uint16_t* Pp = reinterpret_cast<uint16_t*>(rs_raw + RS::kKeyBytes);                 // [kCap] i << 8 | j
    uint16_t* cnt = Pp + RS::kCap;                                                      // [NW][kCntStride]
    uint16_t* tot = cnt + NW * kCntStride;                                              // [kDigMax] elements per digit
    uint16_t* dbase = tot + kDigMax;                                                    // [kDigMax] first position per digit
    uint32_t* red = reinterpret_cast<uint32_t*>(dbase + kDigMax);                       // [4 NW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int c = blockIdx.x; c < p.C; c += gridDim.x) {
        const int b = p.c0 + c;
        const int n = cloud_n(p, b);
        const int E = (int)c2(n);
        const float* Db = p.D + (size_t)b * p.strideB;
        // ---- keys in descending edge index (a warp per matrix row, lanes along the row: coalesced reads)
        int valid = 0, nan_seen = 0;
        uint32_t k_or = 0, k_and = 0xFFFFFFFFu;
        for (int j = warp; j < n - 1; j += NW) {
            const float* row = Db + (size_t)j * p.ld;
            for (int i = j + 1 + lane; i < n; i += 32) {
                const float d = __ldg(row + i) + 0.0f;
                nan_seen |= (d != d);
                const bool ok = d <= p.thresh;
                const uint32_t key = ok ? float_key(d) : kInf;
                const int pos = E - 1 - (i * (i - 1) / 2 + j);
                K[pos] = key;
                Pp[pos] = (uint16_t)((i << 8) | j);
                if (ok) { ++valid; k_or |= key; k_and &= key; }
            }
        }
        valid = __reduce_add_sync(kFull, valid);
        nan_seen = __reduce_or_sync(kFull, nan_seen);
        k_or = __reduce_or_sync(kFull, k_or);
        k_and = __reduce_and_sync(kFull, k_and);
        if (lane == 0) { red[warp] = (uint32_t)valid; red[NW + warp] = k_or; red[2 * NW + warp] = k_and; red[3 * NW + warp] = (uint32_t)nan_seen; }
        __syncthreads();
        int m = 0;
        uint32_t vor = 0, vand = 0xFFFFFFFFu, any_nan = 0;
        for (int w = 0; w < NW; ++w) { m += (int)red[w]; vor |= red[NW + w]; vand &= red[2 * NW + w]; any_nan |= red[3 * NW + w]; }
        // with absent edges (d > thresh, NaN) in between every pass runs: they have to travel to the end
        const uint32_t varying = (m < E) ? 0xFFFFFFFFu : (vor ^ vand);
        const int nch = (E + NTH - 1) / NTH;          // chunks of 32 elements per warp
        const int seg0 = warp * nch * 32;             // first element of this warp's segment
        uint32_t* row32 = reinterpret_cast<uint32_t*>(cnt + warp * kCntStride);
        // distances of a normalised cloud lie in [2^-15, 2): the five top bits of every key are the same and the 27
        // others sort in THREE passes of 9-bit digits (512 counters per warp); otherwise four passes of 8
        const int db = (NINE && !(varying >> 27)) ? 9 : 8;
        const uint32_t dmask = (1u << db) - 1u;
        const int ndig = 1 << db;
        for (int pass = 0; pass * db < 32; ++pass) {
            const int sh = db * pass;
            if (!((varying >> sh) & dmask)) continue;
            {
                uint32_t* c32 = reinterpret_cast<uint32_t*>(cnt);
                for (int q = tid; q < NW * kCntStride / 2; q += NTH) c32[q] = 0;
            }
            __syncthreads();
            // ---- this warp's segment into registers, its digits counted into the warp's counter row
            uint32_t kr[EPT];
#pragma unroll
            for (int t = 0; t < EPT; ++t) {
                const int k = seg0 + 32 * t + lane;
                const bool act = t < nch && k < E;
                kr[t] = act ? K[k] : 0u;
                if (act) {
                    const uint32_t dg = (kr[t] >> sh) & dmask;
                    atomicAdd(row32 + (dg >> 1), 1u << (16 * (dg & 1u)));
                }
            }
            __syncthreads();
            // ---- per digit: exclusive scan over the warps (lane = warp row), total per digit
            const int dpw = ndig / NW;   // digits scanned per warp
            for (int dd = 0; dd < dpw; ++dd) {
                const int d = warp * dpw + dd;
                const uint32_t own = lane < NW ? cnt[lane * kCntStride + d] : 0u;
                uint32_t incl = own;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += y;
                }
                if (lane < NW) cnt[lane * kCntStride + d] = (uint16_t)(incl - own);
                if (lane == 31) tot[d] = (uint16_t)incl;
            }
            __syncthreads();
            if (warp == 0) {   // first position of every digit: 8 or 16 digits per lane
                const int dpl = ndig >> 5;
                uint32_t v[kDigMax / 32], run = 0;
#pragma unroll
                for (int q = 0; q < kDigMax / 32; ++q) { v[q] = q < dpl ? tot[dpl * lane + q] : 0u; run += v[q]; }
                uint32_t incl = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += y;
                }
                uint32_t base = incl - run;
#pragma unroll
                for (int q = 0; q < kDigMax / 32; ++q) { if (q < dpl) dbase[dpl * lane + q] = (uint16_t)base; base += v[q]; }
            }
            __syncthreads();
            // ---- keys to their positions (stable: chunks in order, lanes ranked inside a chunk)
            uint32_t ps[EPT / 2];
#pragma unroll
            for (int t = 0; t < EPT; ++t) {
                const int k = seg0 + 32 * t + lane;
                const bool act = t < nch && k < E;
                // lanes with the same digit: a ballot per digit bit.  __match_any_sync takes a step per distinct
                // value, ~30 of them among 32 random digits, and was half of this kernel's stall samples
                const uint32_t dg = (kr[t] >> sh) & dmask;
                uint32_t peers = __ballot_sync(kFull, act);
#pragma unroll
                for (int bt = 0; bt < 9; ++bt) {
                    const uint32_t bal = __ballot_sync(kFull, (dg >> bt) & 1u);
                    peers &= ((dg >> bt) & 1u) ? bal : ~bal;
                }
                // the first lane of a digit group takes the group's slots from the warp's counter (one atomic on the
                // packed pair of 16-bit counters: no carry, a cloud has fewer than 2^16 edges) and hands the old value
                // on -- no read / barrier / write chain from one chunk to the next
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (act && lane == leader) old = atomicAdd(row32 + (dg >> 1), (uint32_t)__popc(peers) << (16 * (dg & 1u)));
                old = __shfl_sync(kFull, old, leader);
                uint32_t pos = 0;
                if (act) {
                    pos = dbase[dg] + ((old >> (16 * (dg & 1u))) & 0xFFFFu) + __popc(peers & lt);
                    K[pos] = kr[t];
                }
                if (t & 1) ps[t >> 1] |= pos << 16; else ps[t >> 1] = pos;
            }
            // ---- the payloads follow through the saved positions
            uint32_t pv[EPT / 2];
#pragma unroll
            for (int t = 0; t < EPT; ++t) {
                const int k = seg0 + 32 * t + lane;
                const bool act = t < nch && k < E;
                const uint32_t v = act ? (uint32_t)Pp[k] : 0u;
                if (t & 1) pv[t >> 1] |= v << 16; else pv[t >> 1] = v;
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < EPT; ++t) {
                const int k = seg0 + 32 * t + lane;
                const bool act = t < nch && k < E;
                if (act) Pp[(ps[t >> 1] >> (16 * (t & 1))) & 0xFFFFu] = (uint16_t)((pv[t >> 1] >> (16 * (t & 1))) & 0xFFFFu);
            }
            __syncthreads();
        }
        __syncthreads();
        // ---- sorted position -> P and the sorted keys
        uint32_t* Pc = p.P + (size_t)c * p.Emax;
        uint32_t* sk = p.skey + (size_t)c * p.Emax;
        for (int r = tid; r < m; r += NTH) {
            const uint32_t key = K[r], pay = Pp[r];
            const bool tie = r + 1 < m && K[r + 1] == key;
            Pc[r] = (pay & 255u) | ((pay >> 8) << 11) | (tie ? kTieNext : 0u);
            sk[r] = key;
        }
        __syncthreads();
        // ---- the rank matrix, assembled in the key array's space, leaves as 16-byte rows; Q = none
        uint16_t* Ts = reinterpret_cast<uint16_t*>(K);
        const int ldT = p.ldT;
        const int nt4 = n * ldT * 2 / 16;
        {
            uint4* t4 = reinterpret_cast<uint4*>(Ts);
            for (int q = tid; q < nt4; q += NTH) t4[q] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
        __syncthreads();
        for (int r = tid; r < m; r += NTH) {
            const uint32_t pay = Pp[r];
            const int i = (int)(pay >> 8), j = (int)(pay & 255u);
            Ts[i * ldT + j] = (uint16_t)r;
            Ts[j * ldT + i] = (uint16_t)r;
        }
        __syncthreads();
        {
            uint4* tg = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.T) + (size_t)c * p.N * ldT);
            uint4* qg = reinterpret_cast<uint4*>(p.Q + (size_t)c * p.N * ldT);
            const uint4* t4 = reinterpret_cast<const uint4*>(Ts);
            for (int q = tid; q < nt4; q += NTH) { tg[q] = t4[q]; qg[q] = make_uint4(0u, 0u, 0u, 0u); }
        }
        if (tid == 0) { p.m[c] = n >= 2 ? m : 0; p.nanflag[c] = any_nan ? 1 : 0; }
        __syncthreads();   // the arrays are rewritten for the next cloud
    }
}

// block-wide minimum of an int (every thread gets the result); sm[] holds >= 33 ints
template <int NTH> __device__ __forceinline__ int block_min(int v, int* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_min_sync(kFull, v);
    if (NTH == 32) return v;   // a one-warp CTA: the warp reduction is the block reduction
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    int x = lane < NTH / 32 ? sm[lane] : 0x7FFFFFFF;
    x = __reduce_min_sync(kFull, x);
    __syncthreads();
    return x;
}
template <int NTH> __device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_max_sync(kFull, v);
    if (NTH == 32) return v;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    uint32_t x = lane < NTH / 32 ? sm[lane] : 0u;
    x = __reduce_max_sync(kFull, x);
    __syncthreads();
    return x;
}

// ------------------------------------------------------------------------------------ K3 kruskal
template <int NTH> __global__ void __launch_bounds__(NTH) kruskal_kernel(Params p) {
    __shared__ uint16_t comp[kMaxN];
    __shared__ uint16_t eld[kMaxN];
    __shared__ int red[34];
    __shared__ uint32_t pub;
    const int tid = threadIdx.x;
    for (int c = blockIdx.x; c < p.C; c += gridDim.x) {
        const int b = p.c0 + c;
        const int n = cloud_n(p, b);
        const int m = n >= 2 ? p.m[c] : 0;
        uint32_t* Pc = p.P + (size_t)c * p.Emax;
        const float* Db = p.D + (size_t)b * p.strideB;
        for (int v = tid; v < n; v += NTH) { comp[v] = (uint16_t)v; eld[v] = (uint16_t)v; }
        __syncthreads();
        int ncomp = n, n0 = 0;
        for (int r = 0; r < m && ncomp > 1; r += NTH) {
            const int rr = r + tid;
            uint32_t pe = 0;
            bool cand = false;
            if (rr < m) {
                pe = Pc[rr];
                cand = comp[p_i(pe)] != comp[p_j(pe)];
            }
            while (ncomp > 1) {
                const int first = block_min<NTH>(cand ? rr : 0x7FFFFFFF, red);
                if (first == 0x7FFFFFFF) break;
                if (rr == first) { pub = pe; Pc[rr] = pe | kMst; cand = false; }
                __syncthreads();
                const uint32_t q = pub;
                const int i = p_i(q), j = p_j(q);
                const int ci = comp[i], cj = comp[j];
                const int ei = eld[ci], ej = eld[cj];
                const float d = Db[(size_t)j * p.ld + i] + 0.0f;
                if (d != 0.0f) {
                    if (tid == 0 && n0 < p.cap0) {
                        const size_t o = ((size_t)b * p.cap0 + n0) * 2;
                        p.bd0[o] = 0.0f;
                        p.bd0[o + 1] = d;
                        if (p.pr0) { p.pr0[o] = min(ei, ej); p.pr0[o + 1] = c2(i) + j; }
                    }
                    ++n0;
                }
                __syncthreads();
                for (int v = tid; v < n; v += NTH)
                    if (comp[v] == ci) comp[v] = (uint16_t)cj;
                if (tid == 0) eld[cj] = (uint16_t)max(ei, ej);
                __syncthreads();
                --ncomp;
                if (cand) cand = comp[p_i(pe)] != comp[p_j(pe)];
            }
        }
        // essential classes: eldest vertex of every surviving component, ascending
        for (int v0 = 0; v0 < n; v0 += NTH) {
            const int v = v0 + tid;
            const bool is = v < n && eld[comp[v]] == v;
            const uint32_t bal = __ballot_sync(kFull, is);
            if ((tid & 31) == 0) red[tid >> 5] = __popc(bal);
            __syncthreads();
            int base = n0, tot = 0;
            for (int w = 0; w < NTH / 32; ++w) { if (w < (tid >> 5)) base += red[w]; tot += red[w]; }
            if (is) {
                const int pos = base + __popc(bal & ((1u << (tid & 31)) - 1u));
                if (pos < p.cap0) {
                    const size_t o = ((size_t)b * p.cap0 + pos) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = v; p.pr0[o + 1] = -1; }
                }
            }
            n0 += tot;
            __syncthreads();
        }
        if (tid == 0) p.counts[2 * b] = n0;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ K4 classify
// one THREAD per edge: walk the apexes downwards, four ranks per load, until the first cofacet.
// Lanes of a warp hold consecutive ranks, i.e. edges of nearly equal length whose neighbourhoods
// are equally dense, so their scans have similar lengths (~2 sqrt(N) on average: most edges are
// late and find an apex within the first few candidates).
// ONLY_TIED: the members of tie runs only (the audio path classifies everything else from adjacency bit rows,
// classify_bits_kernel below)
template <typename TT, bool ONLY_TIED>
__global__ void __launch_bounds__(256) classify_kernel(Params p) {
    constexpr int VPL = 16 / (int)sizeof(TT);   // ranks per 16-byte load
    constexpr uint32_t kAbsent = RankOf<TT>::kAbsent;
    // ONLY_TIED: the (cloud, rank) items classify_bits_kernel left over, from its list; else every edge of the chunk
    const uint32_t emax = (uint32_t)p.Emax;
    const uint32_t total = ONLY_TIED ? (uint32_t)*p.n_left : (uint32_t)(p.Emax * p.C);   // (< 2^31: make_plan)
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
        int c, r;
        if (ONLY_TIED) { const uint2 it = p.left[w]; c = (int)it.x; r = (int)it.y; }
        else { c = (int)(w / emax); r = (int)(w - (uint32_t)c * emax); }
        if (r >= p.m[c]) continue;
        uint32_t* Pc = p.P + (size_t)c * p.Emax;
        const uint32_t q = Pc[r];
        if (q & kMst) continue;
        const int n = cloud_n(p, p.c0 + c);
        const TT* Tc = reinterpret_cast<const TT*>(p.T) + (size_t)c * p.N * p.ldT;
        const uint32_t* keys = p.skey + (size_t)c * p.Emax;
        const int i = p_i(q), j = p_j(q);
        const bool tied = (q & kTieNext) || (r > 0 && (Pc[r - 1] & kTieNext));

        const uint32_t k32r = tied ? keys[r] : 0u;
        const uint4* Ti = reinterpret_cast<const uint4*>(Tc + (size_t)i * p.ldT);
        const uint4* Tj = reinterpret_cast<const uint4*>(Tc + (size_t)j * p.ldT);
        int dv = -1;
        bool found = false;
        if constexpr (sizeof(TT) == 4 && !ONLY_TIED) {
            // an early edge of a big cloud: candidates = common neighbours in the bit rows of the first level above
            // its rank (a superset of its apexes, tie run included), largest first, each checked on the ranks
            int lv = -1;
            if (p.lev) {
                const int* R = p.levR + c * 2 * kLevels;
#pragma unroll
                for (int k = kLevels - 1; k >= 0; --k) if (r < R[k]) lv = k;
            }
            if (lv >= 0) {
                const uint32_t* Lc = p.lev + ((size_t)c * kLevels + lv) * p.N * kLevWords;
                const uint4* ai = reinterpret_cast<const uint4*>(Lc + (size_t)i * kLevWords);
                const uint4* aj = reinterpret_cast<const uint4*>(Lc + (size_t)j * kLevWords);
                const uint32_t* Ti32 = reinterpret_cast<const uint32_t*>(Tc + (size_t)i * p.ldT);
                const uint32_t* Tj32 = reinterpret_cast<const uint32_t*>(Tc + (size_t)j * p.ldT);
                for (int w4 = (n - 1) >> 7; w4 >= 0 && !found; --w4) {
                    const uint4 a = __ldg(ai + w4), b = __ldg(aj + w4);
                    const uint32_t cm[4] = {a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w};
#pragma unroll
                    for (int k = 3; k >= 0; --k) {
                        uint32_t word = cm[k];
                        while (word && !found) {
                            const int bit = 31 - __clz(word);
                            word &= ~(1u << bit);
                            const int v = 32 * (4 * w4 + k) + bit;
                            const uint32_t ta = __ldg(Ti32 + v), tb = __ldg(Tj32 + v);
                            bool ina = ta < (uint32_t)r, inb = tb < (uint32_t)r;
                            const bool strict = ina && inb;
                            if (tied) {
                                if (!ina && ta != kAbsent && ta > (uint32_t)r) ina = keys[ta] == k32r;
                                if (!inb && tb != kAbsent && tb > (uint32_t)r) inb = keys[tb] == k32r;
                            }
                            if (ina && inb) { found = true; dv = strict ? v : -1; }
                        }
                    }
                }
                found = true;   // (no candidate left: a birth)
            }
        }
        for (int vb = (n - 1) / VPL; vb >= 0 && !found; --vb) {
            const uint4 a4 = __ldg(Ti + vb), b4 = __ldg(Tj + vb);
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int k = VPL - 1; k >= 0; --k) {
                if (found || VPL * vb + k >= n) continue;
                const uint32_t ta = sizeof(TT) == 4 ? aw[k] : ((aw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
                const uint32_t tb = sizeof(TT) == 4 ? bw[k] : ((bw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
                bool ina = ta < (uint32_t)r, inb = tb < (uint32_t)r;
                const bool strict = ina && inb;
                if (tied) {
                    // an edge of the same tie run that comes later in the order is present too
                    if (!ina && ta != kAbsent && ta > (uint32_t)r) ina = keys[ta] == k32r;
                    if (!inb && tb != kAbsent && tb > (uint32_t)r) inb = keys[tb] == k32r;
                }
                if (ina && inb) {
                    found = true;
                    dv = strict ? VPL * vb + k : -1;   // the first cofacet must have the edge as youngest edge
                }
            }
        }
        if (dv < 0) { Pc[r] = q | kBirth; atomicAdd(p.nbirth + c, 1); }
        else p.defv[(size_t)c * p.Emax + r] = (uint16_t)dv;
    }
}

// K4 for clouds up to 256 points (the audio path): one WARP per cloud walks the sorted edge list 32 ranks at a
// time and keeps the adjacency of all earlier ranks as bit rows in shared memory, so "the largest apex whose two
// other edges are earlier" is the top bit of adj[i] & adj[j] -- all apexes of an edge in a handful of word
// operations, 32 edges per step -- instead of a walk down two rank rows per edge (which cost as much as the sort).
// Edges of the same step that share a vertex with the lane's edge are looked at one by one (touched[v] = lanes
// of the step with an edge at v); members of tie runs are left to classify_kernel<ONLY_TIED>.
template <int NWORDS>   // 32-bit words per adjacency row: 4 up to 128 points, 8 up to 256
__global__ void __launch_bounds__(256) classify_bits_kernel(Params p) {
    extern __shared__ __align__(16) uint32_t cb_raw[];
    constexpr int NV = 32 * NWORDS;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    constexpr int kPerWarp = NV * NWORDS + NV + NV / 2;            // words: bit rows, vertex lists, comp + eld bytes
    uint32_t* adj = cb_raw + (size_t)wib * kPerWarp;               // [NV][NWORDS]
    uint32_t* touched = adj + NV * NWORDS;                         // [NV]
    uint8_t* comp = reinterpret_cast<uint8_t*>(touched + NV);      // [NV] Kruskal: component of a vertex
    uint8_t* eld = comp + NV;                                      // [NV] eldest vertex of a component
    const uint32_t lt = (1u << lane) - 1u;
    for (int c = blockIdx.x * wpb + wib; c < p.C; c += gridDim.x * wpb) {
        const int b = p.c0 + c;
        const int n = cloud_n(p, b);
        const int m = n >= 2 ? p.m[c] : 0;
        uint32_t* Pc = p.P + (size_t)c * p.Emax;
        uint16_t* dvc = p.defv + (size_t)c * p.Emax;
        const float* Db = p.D + (size_t)b * p.strideB;
        for (int q = lane; q < NV * NWORDS + NV; q += 32) adj[q] = 0;
        for (int v = lane; v < NV; v += 32) { comp[v] = (uint8_t)v; eld[v] = (uint8_t)v; }
        __syncwarp();
        int ncomp = n, n0 = 0;   // K3 (Kruskal, H0 pairs by the elder rule) rides on the same walk
        int births = 0;
        uint32_t prev_tie = 0;   // "rank c0 - 1 has the same length as rank c0"
        uint32_t qn = lane < m ? __ldg(Pc + lane) : 0u;
        for (int c0 = 0; c0 < m; c0 += 32) {
            const int r = c0 + lane;
            const bool act = r < m;
            uint32_t q = qn;
            qn = r + 32 < m ? __ldg(Pc + r + 32) : 0u;   // (the flags this kernel sets are not read back)
            const int i = p_i(q), j = p_j(q);
            // ---- merging edges of this step, in rank order
            if (ncomp > 1) {
                uint32_t bal = __ballot_sync(kFull, act && comp[i] != comp[j]);
                uint32_t mst = 0;
                while (bal && ncomp > 1) {
                    const int src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const uint32_t qs = __shfl_sync(kFull, q, src);
                    const int is = p_i(qs), js = p_j(qs);
                    const int ci = comp[is], cj = comp[js];
                    if (ci == cj) continue;
                    const int ei = eld[ci], ej = eld[cj];
                    const float d = __ldg(Db + (size_t)js * p.ld + is) + 0.0f;
                    if (d != 0.0f) {
                        if (lane == 0 && n0 < p.cap0) {
                            const size_t o = ((size_t)b * p.cap0 + n0) * 2;
                            p.bd0[o] = 0.0f;
                            p.bd0[o + 1] = d;
                            if (p.pr0) { p.pr0[o] = min(ei, ej); p.pr0[o + 1] = c2(is) + js; }
                        }
                        ++n0;
                    }
                    __syncwarp();
                    for (int v = lane; v < n; v += 32)
                        if (comp[v] == ci) comp[v] = (uint8_t)cj;
                    if (lane == 0) eld[cj] = (uint8_t)max(ei, ej);
                    __syncwarp();
                    mst |= 1u << src;
                    --ncomp;
                }
                if ((mst >> lane) & 1u) { q |= kMst; Pc[r] = q; }
            }
            // ---- tie runs inside the step.  The first cofacet of a run member is decided with the whole run present
            //      (lanes below `bl`, the end of its run), and it is an apparent pair only if both other edges of
            //      that triangle are strictly earlier (lanes below its own).  A run that reaches into the previous
            //      or the next step is left to classify_kernel<ONLY_TIED>.
            const uint32_t tn = __ballot_sync(kFull, act && (q & kTieNext));
            const uint32_t nx = ~(tn >> lane);
            const int bl = lane + (nx ? __ffs(nx) : 33);                                   // (exclusive); > 32: into the next step
            const int nbelow = lane ? __clz(~((tn & lt) << (32 - lane))) : 0;              // run members below this lane
            const bool cross = bl > 32 || (nbelow == lane && prev_tie);
            prev_tie = tn >> 31;
            const uint32_t pm = (bl >= 32 ? kFull : ((1u << bl) - 1u)) & ~(1u << lane);     // present by the end of the run
            if (act) { atomicOr(touched + i, 1u << lane); atomicOr(touched + j, 1u << lane); }
            __syncwarp();
            const bool want = act && !(q & kMst) && !cross;
            {   // members of a run that crosses a step boundary go on the list of the rank-row walk
                const bool push = act && !(q & kMst) && cross;
                const uint32_t pb = __ballot_sync(kFull, push);
                if (pb) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(p.n_left, __popc(pb));
                    base = __shfl_sync(kFull, base, 0);
                    if (push) p.left[base + __popc(pb & lt)] = make_uint2((uint32_t)c, (uint32_t)r);
                }
            }
            int vc = -1;
            bool strict = true;   // both other edges of the triangle (i, j, vc) are strictly earlier than this edge
            uint32_t tl = 0;
            if (want) {
                const uint4* ai = reinterpret_cast<const uint4*>(adj + i * NWORDS);
                const uint4* aj = reinterpret_cast<const uint4*>(adj + j * NWORDS);
#pragma unroll
                for (int w4 = NWORDS / 4 - 1; w4 >= 0; --w4) {
                    const uint4 a = ai[w4], b = aj[w4];
                    const uint32_t cm[4] = {a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w};
#pragma unroll
                    for (int k = 3; k >= 0; --k)
                        if (vc < 0 && cm[k]) vc = 32 * (4 * w4 + k) + 31 - __clz(cm[k]);
                }
                tl = (touched[i] | touched[j]) & pm;
            }
            while (__any_sync(kFull, tl != 0)) {
                const int l2 = tl ? __ffs(tl) - 1 : lane;
                const uint32_t q2 = __shfl_sync(kFull, q, l2);
                if (tl) {
                    tl &= tl - 1;
                    const int a = p_i(q2), b2 = p_j(q2);
                    const bool at_a = (a == i) || (a == j);
                    const int s = at_a ? a : b2, v = at_a ? b2 : a;  // shared vertex, candidate apex
                    const int o = (s == i) ? j : i;                  // the end point of this edge that is not shared
                    if (v > vc) {
                        const bool pre = (adj[o * NWORDS + (v >> 5)] >> (v & 31)) & 1u;
                        const uint32_t l3 = touched[o] & touched[v] & pm;   // edge (o, v) in this step, inside the run's reach
                        if (pre || l3) {
                            vc = v;
                            strict = l2 < lane && (pre || (l3 & lt) != 0);
                        }
                    }
                }
            }
            __syncwarp();
            if (act) { touched[i] = 0; touched[j] = 0; }
            if (want) {
                if (vc < 0 || !strict) { Pc[r] = q | kBirth; ++births; }
                else dvc[r] = (uint16_t)vc;
            }
            if (act) {
                atomicOr(adj + i * NWORDS + (j >> 5), 1u << (j & 31));
                atomicOr(adj + j * NWORDS + (i >> 5), 1u << (i & 31));
            }
            __syncwarp();
        }
        births = __reduce_add_sync(kFull, births);
        if (lane == 0 && births) atomicAdd(p.nbirth + c, births);
        // ---- H0 essential classes: eldest vertex of every surviving component, ascending
        __syncwarp();
        for (int v0 = 0; v0 < n; v0 += 32) {
            const int v = v0 + lane;
            const bool is = v < n && eld[comp[v]] == v;
            const uint32_t bal = __ballot_sync(kFull, is);
            if (is) {
                const int pos = n0 + __popc(bal & lt);
                if (pos < p.cap0) {
                    const size_t o = ((size_t)b * p.cap0 + pos) * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = v; p.pr0[o + 1] = -1; }
                }
            }
            n0 += __popc(bal);
        }
        if (lane == 0) p.counts[2 * b] = n0;
        __syncwarp();
    }
}

// longest-processing-time-first order of the sweep queue: position of a cloud = number of clouds with more
// births (ties: smaller index first).  The number of visited edges of a sweep grows with the births.
__global__ void __launch_bounds__(256) order_kernel(Params p, int* order) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    const int mine = p.nbirth[c];
    int pos = 0;
    for (int o = 0; o < p.C; ++o) {
        const int nb = __ldg(p.nbirth + o);
        pos += (nb > mine) || (nb == mine && o < c);
    }
    order[pos] = c;
}

// ------------------------------------------------------------------------------------ K5 sweep
template <int NTH, int APT, int W, bool SG, typename TT> struct Sweep {
    static constexpr uint32_t kAbsent = RankOf<TT>::kAbsent;
    // shared state
    uint32_t* S;      // [n][W]  (shared or global)
    uint32_t* hot;    // [kMaxN / 32]
    uint32_t* live;   // [W]
    uint32_t* used;   // [W]
    uint32_t* brank;  // [32 W]
    uint32_t* bstart; // [32 W] number of compact entries when the class of a slot was born
    uint32_t* cv;     // [W] broadcast of the firing mask
    uint32_t* peval;  // [W] value of the edge handled by the fast path
    uint32_t* pub;    // [4] published edge words
    uint32_t* red;    // [34]
    int* ctl;         // [8]: npc, na, n1, overflow, resume, winner rank, fast-path code
    // per-CTA global scratch
    uint32_t* phic;
    uint32_t* pcr;
    uint32_t* act;
    uint32_t* rec;
    // per-cloud
    const uint32_t* P;
    const TT* T;
    uint16_t* Q;
    const uint16_t* defv;
    const float* Db;
    int n, m, ldT, ld, tid, lane, warp, capP, capR;

    // a one-warp CTA needs no block barrier
    __device__ __forceinline__ void bar() const {
        if constexpr (NTH == 32) __syncwarp(); else __syncthreads();
    }
    __device__ __forceinline__ float dist(int a, int b) const {
        return Db[(size_t)min(a, b) * ld + max(a, b)] + 0.0f;
    }
    __device__ __forceinline__ bool hotbit(int v) const { return (hot[v >> 5] >> (v & 31)) & 1u; }
    __device__ __forceinline__ bool live_any() const {
        uint32_t a = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) a |= live[w];
        return a != 0;
    }
    __device__ void rebuild_hot() {
        for (int v0 = 0; v0 < ((n + 31) & ~31); v0 += NTH) {
            const int v = v0 + tid;
            bool h = false;
            if (v < n) {
                uint32_t a = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) a |= S[(size_t)v * W + w] & live[w];
                h = a != 0;
            }
            const uint32_t bal = __ballot_sync(kFull, h);
            if (lane == 0 && v < ((n + 31) & ~31)) hot[v >> 5] = bal;
        }
    }
    // warp 0 only: append a compact PHI entry for edge (x, y) of rank pr; val = this lane's word
    __device__ __forceinline__ bool append(int pr, int x, int y, uint32_t val) {
        const int idx = ctl[0];
        if (idx >= capP) { ctl[3] = 1; return false; }
        if (lane < W) {
            phic[(size_t)idx * W + lane] = val;
            S[(size_t)x * W + lane] |= val;
            S[(size_t)y * W + lane] |= val;
        }
        if (lane == 0) {
            pcr[idx] = (uint32_t)pr;
            Q[(size_t)x * ldT + y] = (uint16_t)(idx + 1);
            Q[(size_t)y * ldT + x] = (uint16_t)(idx + 1);
            hot[x >> 5] |= 1u << (x & 31);
            hot[y >> 5] |= 1u << (y & 31);
            ctl[0] = idx + 1;
        }
        __syncwarp();
        return true;
    }
    // warp 0 only: birth / definition of one edge.  0 = no live cocycle sees it, 1 = visited,
    // 2 = birth without a free slot (nothing changed; the CTA scrubs and calls again).
    // Leaves the edge's value in peval[].
    // REC: the edge goes on the list of visited edges of a tie run (a single edge needs no list)
    template <bool REC = true> __device__ int edge_a(int pr, uint32_t pe, int dv) {
        const int x = p_i(pe), y = p_j(pe);
        uint32_t val = 0;
        if (pe & kBirth) {
            const uint32_t f = lane < W ? ~used[lane] : 0u;
            const uint32_t bal = __ballot_sync(kFull, f != 0);
            if (!bal) return 2;
            const int sw = __ffs(bal) - 1;
            const uint32_t fw = __shfl_sync(kFull, f, sw);
            const int sb = __ffs(fw) - 1;
            val = (lane == sw) ? (1u << sb) : 0u;
            if (lane == sw) { used[lane] |= 1u << sb; live[lane] |= 1u << sb; }
            if (lane == 0) { brank[32 * sw + sb] = (uint32_t)pr; bstart[32 * sw + sb] = (uint32_t)ctl[0]; }
            __syncwarp();
        } else {
            const uint32_t t = lane < W ? (S[(size_t)x * W + lane] | S[(size_t)y * W + lane]) & live[lane] : 0u;
            if (!__ballot_sync(kFull, t != 0)) return 0;
            const int qa = Q[(size_t)x * ldT + dv], qb = Q[(size_t)y * ldT + dv];
            if (lane < W) {
                if (qa) val ^= phic[(size_t)(qa - 1) * W + lane];
                if (qb) val ^= phic[(size_t)(qb - 1) * W + lane];
                val &= live[lane];
            }
        }
        if (lane < W) peval[lane] = val;
        if (__ballot_sync(kFull, val != 0)) {
            if (!append(pr, x, y, val)) return 0;
        }
        if constexpr (REC) {
            if (lane == 0) { act[ctl[1]] = (uint32_t)pr; ctl[1] = ctl[1] + 1; }
            __syncwarp();
        }
        return 1;
    }
    // warp 0 only.  Returns the rank to resume from after a scrub, or r1 when done.
    __device__ int step_a(int ra, int r1) {
        for (int pr = ra; pr < r1; ++pr) {
            const uint32_t pe = __ldg(P + pr);
            if (pe & kMst) continue;
            const int dv = (pe & kBirth) ? 0 : (int)defv[pr];
            const int code = edge_a(pr, pe, dv);
            if (code == 2) return pr;
            if (ctl[3]) return r1;
        }
        return r1;
    }
    __device__ void scrub() {
        const int npc = ctl[0];
        for (int q = tid; q < npc * W; q += NTH) phic[q] &= live[q % W];
        for (int q = tid; q < n * W; q += NTH) S[q] &= live[q % W];
        bar();
        if (tid < W) used[tid] = live[tid];
        bar();
    }
    // o ^= compact entry idx (a mask of W words = 8 W bytes, aligned: 16-byte loads)
    __device__ __forceinline__ void xor_entry(int idx, uint32_t (&o)[W]) const {
        if constexpr (W % 4 == 0) {
            const uint4* e = reinterpret_cast<const uint4*>(phic + (size_t)idx * W);
#pragma unroll
            for (int w = 0; w < W / 4; ++w) {
                const uint4 t = e[w];
                o[4 * w] ^= t.x; o[4 * w + 1] ^= t.y; o[4 * w + 2] ^= t.z; o[4 * w + 3] ^= t.w;
            }
        } else {
            const uint2* e = reinterpret_cast<const uint2*>(phic + (size_t)idx * W);
#pragma unroll
            for (int w = 0; w < W / 2; ++w) {
                const uint2 t = e[w];
                o[2 * w] ^= t.x; o[2 * w + 1] ^= t.y;
            }
        }
    }
    // coboundary mask of the live cocycles on triangle (x, y, z), given the value pe[] on (x, y) and the
    // compact indices qa, qb of (x, z), (y, z)
    __device__ __forceinline__ bool cob(const uint32_t (&pe)[W], int qa, int qb, uint32_t (&c)[W]) const {
#pragma unroll
        for (int w = 0; w < W; ++w) c[w] = pe[w];
        if (qa) xor_entry(qa - 1, c);
        if (qb) xor_entry(qb - 1, c);
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) { c[w] &= live[w]; any |= c[w]; }
        return any != 0;
    }
    // value of edge (x, y), masked by the live classes; false if it is zero
    __device__ __forceinline__ bool load_pe(int x, int y, uint32_t (&pe)[W]) const {
        const int q = Q[(size_t)x * ldT + y];
#pragma unroll
        for (int w = 0; w < W; ++w) pe[w] = 0u;
        if (!q) return false;
        xor_entry(q - 1, pe);
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) { pe[w] &= live[w]; any |= pe[w]; }
        return any != 0;
    }
    // 16-bit ranks: this thread's APT consecutive apexes of the rank rows of x and y and of their index rows, two per
    // word (one 8- or 16-byte load per row; the padding of a row is "absent" / 0, rank_small_kernel)
    static constexpr bool kVec = sizeof(TT) == 2 && (APT == 4 || APT == 8);
    static constexpr int NPK = kVec ? APT / 2 : 1;
    __device__ __forceinline__ void load_rows_packed(int x, int y, uint32_t (&a)[NPK], uint32_t (&b)[NPK],
                                                     uint32_t (&qx)[NPK], uint32_t (&qy)[NPK]) const {
        const TT* Tx = T + (size_t)x * ldT;
        const TT* Ty = T + (size_t)y * ldT;
        const uint16_t* Qx = Q + (size_t)x * ldT;
        const uint16_t* Qy = Q + (size_t)y * ldT;
#pragma unroll
        for (int k = 0; k < NPK; ++k) { a[k] = 0xFFFFFFFFu; b[k] = 0xFFFFFFFFu; qx[k] = 0u; qy[k] = 0u; }
        if (APT * tid < ldT) {
            if constexpr (APT == 4) {
                // (streaming loads: the ~1 GB of rank / index tables of the resident clouds should not push the compact
                // stores of the CTAs -- a few KB each, re-used cloud after cloud -- out of L2)
                const uint2 ta2 = __ldcs(reinterpret_cast<const uint2*>(Tx) + tid), tb2 = __ldcs(reinterpret_cast<const uint2*>(Ty) + tid);
                const uint2 qx2 = __ldcs(reinterpret_cast<const uint2*>(Qx) + tid), qy2 = __ldcs(reinterpret_cast<const uint2*>(Qy) + tid);
                a[0] = ta2.x; a[1] = ta2.y; b[0] = tb2.x; b[1] = tb2.y;
                qx[0] = qx2.x; qx[1] = qx2.y; qy[0] = qy2.x; qy[1] = qy2.y;
            } else if constexpr (APT == 8) {
                const uint4 ta4 = __ldcs(reinterpret_cast<const uint4*>(Tx) + tid), tb4 = __ldcs(reinterpret_cast<const uint4*>(Ty) + tid);
                const uint4 qx4 = __ldcs(reinterpret_cast<const uint4*>(Qx) + tid), qy4 = __ldcs(reinterpret_cast<const uint4*>(Qy) + tid);
                a[0] = ta4.x; a[1] = ta4.y; a[2] = ta4.z; a[3] = ta4.w; b[0] = tb4.x; b[1] = tb4.y; b[2] = tb4.z; b[3] = tb4.w;
                qx[0] = qx4.x; qx[1] = qx4.y; qx[2] = qx4.z; qx[3] = qx4.w; qy[0] = qy4.x; qy[1] = qy4.y; qy[2] = qy4.z; qy[3] = qy4.w;
            }
        }
    }
    // death loop over the visited edges act[0 .. na); (spr, spq): rank and P word when the run is one edge
    __device__ void step_b(int r0, int r1, int spr = -1, uint32_t spq = 0u) {
        const int na = ctl[1];
        while (true) {
            uint32_t best = 0, best_pq = 0;
            int best_pr = -1, best_z = -1;
            for (int a = 0; a < na; ++a) {
                const int pr = spr >= 0 ? spr : (int)act[a];
                const uint32_t pq = spr >= 0 ? spq : __ldg(P + pr);
                const int x = p_i(pq), y = p_j(pq);
                const TT* Tx = T + (size_t)x * ldT;
                const TT* Ty = T + (size_t)y * ldT;
                const uint16_t* Qx = Q + (size_t)x * ldT;
                const uint16_t* Qy = Q + (size_t)y * ldT;
                // every load of the scan that does not depend on another one is issued first: the scan is a
                // chain of global round trips otherwise (rank -> index -> compact entry, per apex)
                const int qe = Qx[y];
                uint32_t ta[APT], tb[APT], qq[APT];
                uint32_t ra[NPK], rb[NPK], qx[NPK], qy[NPK];
                if constexpr (kVec) {
                    load_rows_packed(x, y, ra, rb, qx, qy);
                } else {
#pragma unroll
                    for (int h = 0; h < APT; ++h) {
                        const int z = tid + h * NTH;
                        ta[h] = kAbsent; tb[h] = kAbsent; qq[h] = 0u;
                        if (z < n) { ta[h] = __ldg(Tx + z); tb[h] = __ldg(Ty + z); qq[h] = (uint32_t)Qx[z] | ((uint32_t)Qy[z] << 16); }
                    }
                }
                uint32_t pe[W];
#pragma unroll
                for (int w = 0; w < W; ++w) pe[w] = 0u;
                if (qe) xor_entry(qe - 1, pe);
                uint32_t pany = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) { pe[w] &= live[w]; pany |= pe[w]; }
                // this thread's apexes from the largest down: the first triangle with a non-zero mask is its
                // candidate for this edge (the index grows with the apex)
                int zhit = -1;
                if constexpr (kVec) {
                    const uint32_t pr2 = (uint32_t)pr * 0x10001u;
#pragma unroll
                    for (int k = NPK - 1; k >= 0; --k) {
                        const uint32_t vm = __vcmpltu2(__vmaxu2(ra[k], rb[k]), pr2);
#pragma unroll
                        for (int hf = 1; hf >= 0; --hf) {
                            if (zhit < 0 && ((vm >> (16 * hf)) & 0xFFFFu)) {
                                const int qa = (int)((qx[k] >> (16 * hf)) & 0xFFFFu), qb = (int)((qy[k] >> (16 * hf)) & 0xFFFFu);
                                if (qa | qb) {
                                    uint32_t c[W];
                                    if (cob(pe, qa, qb, c)) zhit = APT * tid + 2 * k + hf;
                                } else if (pany) zhit = APT * tid + 2 * k + hf;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int h = APT - 1; h >= 0; --h) {
                        if (zhit < 0 && ta[h] < (uint32_t)pr && tb[h] < (uint32_t)pr) {
                            if (qq[h]) {
                                uint32_t c[W];
                                if (cob(pe, (int)(qq[h] & 0xFFFFu), (int)(qq[h] >> 16), c)) zhit = tid + h * NTH;
                            } else if (pany) zhit = tid + h * NTH;   // both other edges carry 0: the mask is the edge's own value
                        }
                    }
                }
                if (zhit >= 0) {
                    const uint32_t t = tri_index(x, y, zhit) + 1u;
                    if (t > best) { best = t; best_pr = pr; best_pq = pq; best_z = zhit; }
                }
            }
            const uint32_t top = block_max_u32<NTH>(best, red);
            if (top == 0) return;
            if (best == top) {   // the winner publishes its mask (recomputed: keeping it would cost W registers)
                uint32_t pe[W], c[W];
                const int x = p_i(best_pq), y = p_j(best_pq);
                load_pe(x, y, pe);
                cob(pe, Q[(size_t)x * ldT + best_z], Q[(size_t)y * ldT + best_z], c);
#pragma unroll
                for (int w = 0; w < W; ++w) cv[w] = c[w];
                ctl[5] = best_pr;
            }
            bar();
            // youngest class of the mask dies
            int slot = -1, age = -1;
            bool others = false;
            for (int w = 0; w < W; ++w) {
                uint32_t bits = cv[w];
                while (bits) {
                    const int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int ag = (int)brank[32 * w + s];
                    if (slot >= 0) others = true;
                    if (ag > age) { age = ag; slot = 32 * w + s; }
                }
            }
            const int sw = slot >> 5;
            const uint32_t sb = 1u << (slot & 31);
            bar();
            if (tid == 0) {
                if (age < r0) {  // non-zero persistence
                    const int k = ctl[2];
                    if (k < capR) {
                        rec[k] = (uint32_t)age;
                        rec[capR + k] = (uint32_t)ctl[5];
                        rec[2 * capR + k] = top - 1u;
                    } else ctl[3] = 1;
                    ctl[2] = k + 1;
                }
                live[sw] &= ~sb;
            }
            if (others) {
                // the dying cocycle vanishes on every edge older than its birth, and the compact entries are in
                // rank order: only the entries from its birth on can carry its bit (it is the YOUNGEST class of
                // the mask, so this is usually a short tail of the store)
                const int npc = ctl[0];
                const int q0 = (int)bstart[slot];
                for (int qb = q0 + tid; qb < npc; qb += 4 * NTH) {
                    uint32_t wd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int q = qb + u * NTH;
                        wd[u] = q < npc ? phic[(size_t)q * W + sw] : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (wd[u] & sb) {
                            const int q = qb + u * NTH;
                            uint32_t* e = phic + (size_t)q * W;
                            const uint32_t pq = __ldg(P + pcr[q]);
                            const int x = p_i(pq), y = p_j(pq);
#pragma unroll
                            for (int w = 0; w < W; ++w) {
                                const uint32_t v = cv[w];
                                if (v) {
                                    e[w] ^= v;
                                    // (the supports of cocycles are fans around a few vertices: nearly every
                                    // update finds the bits set already, and an atomic on a word that
                                    // hundreds of threads hit serialises)
                                    uint32_t* sx = S + (size_t)x * W + w;
                                    uint32_t* sy = S + (size_t)y * W + w;
                                    if ((*(volatile uint32_t*)sx & v) != v) atomicOr(sx, v);
                                    if ((*(volatile uint32_t*)sy & v) != v) atomicOr(sy, v);
                                }
                            }
                        }
                    }
                }
            }
            bar();
            rebuild_hot();
            bar();
            if (ctl[3]) return;
        }
    }

    // a tie run [r0, r1): A (births / definitions in rank order, warp 0, scrubs when slots run out), B
    __device__ void generic_run(int r0, int r1) {
        if (tid == 0) ctl[1] = 0;
        bar();
        int ra = r0;
        while (true) {
            if (warp == 0) {
                const int nxt = step_a(ra, r1);
                if (lane == 0) ctl[4] = nxt;
            }
            bar();
            ra = ctl[4];
            if (ra >= r1 || ctl[3]) break;
            scrub();  // slots exhausted at rank ra: recycle the dead ones
            bool room = false;
            for (int w = 0; w < W; ++w) room |= (~used[w]) != 0;
            if (!room) { if (tid == 0) ctl[3] = 1; bar(); break; }
        }
        if (ctl[3]) return;
        if (ctl[1] > 0) step_b(r0, r1);
    }
    // an untied edge: the apex rows are fetched while warp 0 does step A, one pass decides whether
    // anything dies (almost never), only then the general death loop runs
    __device__ void fast_single(int pr, uint32_t pe, int dv) {
        const int x = p_i(pe), y = p_j(pe);
        const TT* Tx = T + (size_t)x * ldT;
        const TT* Ty = T + (size_t)y * ldT;
        const uint16_t* Qx = Q + (size_t)x * ldT;
        const uint16_t* Qy = Q + (size_t)y * ldT;
        uint32_t ta[APT], tb[APT];
        int qa[APT], qb[APT];
        // apex of (thread, h): four CONSECUTIVE apexes per thread when the ranks are 16-bit (clouds up to
        // 256 points: NTH * 4 covers the padded row), so each of the four rows is one 8-byte load per thread;
        // the padding of a row is "absent" in T and 0 in Q (rank_small_kernel), no bounds test needed
        uint32_t a[NPK], b[NPK], qx[NPK], qy[NPK];   // kVec: two 16-bit ranks / indices per word, kept packed
        if constexpr (kVec) {
            load_rows_packed(x, y, a, b, qx, qy);
        } else {
#pragma unroll
            for (int h = 0; h < APT; ++h) {
                const int z = tid + h * NTH;
                ta[h] = kAbsent; tb[h] = kAbsent; qa[h] = 0; qb[h] = 0;
                if (z < n) { ta[h] = __ldg(Tx + z); tb[h] = __ldg(Ty + z); qa[h] = Qx[z]; qb[h] = Qy[z]; }
            }
        }
        if (warp == 0) {
            if (lane == 0) ctl[1] = 1;   // one visited edge (the death loop takes it from its arguments)
            const int code = edge_a<false>(pr, pe, dv);
            if (lane == 0) ctl[6] = code;
        }
        bar();
        const int code = ctl[6];
        if (code == 0 || ctl[3]) return;
        if (code == 2) { generic_run(pr, pr + 1); return; }
        // does any triangle over the edge carry a non-zero mask?  (peval is masked by live already)
        uint32_t pv[W], pany = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) { pv[w] = peval[w]; pany |= pv[w]; }
        uint32_t best = 0;
        if constexpr (kVec) {
            // two apexes per word: "both other edges earlier" and "an index entry on one of them" as packed
            // 16-bit compares (absent = 0xFFFF is never below a rank)
            const uint32_t pr2 = (uint32_t)pr * 0x10001u;
#pragma unroll
            for (int k = 0; k < NPK; ++k) {
                const uint32_t vm = __vcmpltu2(__vmaxu2(a[k], b[k]), pr2);
                const uint32_t nzm = __vcmpne2(qx[k] | qy[k], 0u);
                if (pany && (vm & ~nzm)) best = 1u;   // the other edges carry 0: the mask is the edge's own value
                const uint32_t need = vm & nzm;
                if (need & 0xFFFFu) {
                    uint32_t c[W];
                    if (cob(pv, (int)(qx[k] & 0xFFFFu), (int)(qy[k] & 0xFFFFu), c)) best = 1u;
                }
                if (need >> 16) {
                    uint32_t c[W];
                    if (cob(pv, (int)(qx[k] >> 16), (int)(qy[k] >> 16), c)) best = 1u;
                }
            }
        } else {
#pragma unroll
            for (int h = 0; h < APT; ++h) {
                if (ta[h] < (uint32_t)pr && tb[h] < (uint32_t)pr) {
                    if (qa[h] | qb[h]) {
                        uint32_t c[W];
                        if (cob(pv, qa[h], qb[h], c)) best = 1u;
                    } else if (pany) best = 1u;
                }
            }
        }
        const uint32_t top = block_max_u32<NTH>(best, red);
        if (top == 0) return;
        step_b(pr, pr + 1, pr, pe);
    }

    __device__ void run(const Params& p, int c, bool rezero_q) {
        const int b = p.c0 + c;
        n = cloud_n(p, b);
        m = n >= 2 ? p.m[c] : 0;
        ldT = p.ldT;
        ld = p.ld;
        P = p.P + (size_t)c * p.Emax;
        T = reinterpret_cast<const TT*>(p.T) + (size_t)c * p.N * p.ldT;
        Q = p.Q + (size_t)c * p.N * p.ldT;
        defv = p.defv + (size_t)c * p.Emax;
        Db = p.D + (size_t)b * p.strideB;
        if (n < 2) {
            if (tid == 0) {
                if (n == 1 && p.cap0 > 0) {
                    const size_t o = (size_t)b * p.cap0 * 2;
                    p.bd0[o] = 0.0f;
                    p.bd0[o + 1] = __int_as_float(0x7F800000);
                    if (p.pr0) { p.pr0[o] = 0; p.pr0[o + 1] = -1; }
                }
                p.counts[2 * b] = n;
                p.counts[2 * b + 1] = 0;
                p.status[b] = 0;
            }
            return;
        }
        if (rezero_q) {
            for (size_t q = tid; q < (size_t)n * ldT; q += NTH) Q[q] = 0;
        }
        for (int q = tid; q < n * W; q += NTH) S[q] = 0;
        for (int q = tid; q < kMaxN / 32; q += NTH) hot[q] = 0;
        if (tid < W) { live[tid] = 0; used[tid] = 0; }
        if (tid < 8) ctl[tid] = 0;
        bar();
        int rbase = 0;
        uint32_t pe = kMst, pe_nx = kMst;
        int dv = 0, dv_nx = 0;
        if (tid < m) { pe = __ldg(P + tid); dv = defv[tid]; }
        while (rbase < m && !ctl[3]) {
            // ---- this thread's edge of the chunk stays in registers while the chunk is scanned; the words of
            //      the next chunk are on their way meanwhile (defv of a merging edge or a birth is never used)
            const int rr = rbase + tid;
            {
                const int rn = rr + NTH;
                pe_nx = kMst; dv_nx = 0;
                if (rn < m) { pe_nx = __ldcs(P + rn); dv_nx = __ldcs(defv + rn); }
                if constexpr (NTH == 32) {
                    // a chunk without a visited edge is over in a few instructions: the edge words further ahead
                    // are requested from DRAM eight chunks early (one 128-byte line of P, half a line of defv)
                    const int ra = rbase + 8 * NTH;
                    if (tid == 0 && ra < m) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P + ra));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(defv + ra));
                    }
                }
            }
            const int rend = min(rbase + NTH, m);
            int done = rbase;
            while (true) {
                // first edge of the chunk a live cocycle can see (or a birth)
                const bool flag = rr >= done && !(pe & kMst) &&
                                  ((pe & kBirth) || hotbit(p_i(pe)) || hotbit(p_j(pe)));
                const int first = block_min<NTH>(flag ? rr : 0x7FFFFFFF, (int*)red);
                if (first == 0x7FFFFFFF) break;
                uint32_t fpe, ppe = 0;
                int fdv;
                if constexpr (NTH == 32) {   // the chunk is the warp's registers: shuffles, no shared memory, no barrier
                    fpe = __shfl_sync(kFull, pe, first - rbase);
                    fdv = __shfl_sync(kFull, dv, first - rbase);
                    if constexpr (sizeof(TT) == 2) {
                        // the rank and index rows of the edge that is likely to be visited next (the next flagged rank
                        // of the chunk) start their way from DRAM now: the tables of the resident clouds are ~1 GB,
                        // every visited edge otherwise waits a full DRAM round trip for its four rows
                        const int second = __reduce_min_sync(kFull, (flag && rr > first) ? rr : 0x7FFFFFFF);
                        if (second != 0x7FFFFFFF) {
                            const uint32_t pe2 = __shfl_sync(kFull, pe, second - rbase);
                            if (lane < 16) {
                                const int v2 = (lane & 4) ? p_j(pe2) : p_i(pe2);
                                const char* row = (lane & 8) ? reinterpret_cast<const char*>(Q + (size_t)v2 * ldT)
                                                             : reinterpret_cast<const char*>(T + (size_t)v2 * ldT);
                                const int off = (lane & 3) * 128;
                                if (off < ldT * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
                            }
                        }
                    }
                    const uint32_t pv = __shfl_sync(kFull, pe, (first - rbase + 31) & 31);
                    if (first > rbase) ppe = pv;
                    else if (first > 0) ppe = __ldg(P + first - 1);
                } else {
                    if (rr == first) { pub[0] = pe; pub[2] = (uint32_t)dv; }
                    if (rr == first - 1) pub[1] = pe;
                    bar();
                    fpe = pub[0];
                    fdv = (int)pub[2];
                    if (first > rbase) ppe = pub[1];
                    else if (first > 0) ppe = __ldg(P + first - 1);
                }
                int r1 = first + 1;
                if (!((fpe | ppe) & kTieNext)) {
                    fast_single(first, fpe, fdv);
                } else {
                    int r0 = first;
                    while (r0 > 0 && (__ldg(P + r0 - 1) & kTieNext)) --r0;
                    while (__ldg(P + r1 - 1) & kTieNext) ++r1;
                    generic_run(r0, r1);
                }
                bar();
                done = r1;
                if (done >= rend || ctl[3]) break;
            }
            if (done <= rbase + NTH) {
                rbase += NTH; pe = pe_nx; dv = dv_nx;
            } else {   // a tie run reached beyond the chunk
                rbase = done;
                pe = kMst; dv = 0;
                if (rbase + tid < m) { pe = __ldg(P + rbase + tid); dv = defv[rbase + tid]; }
            }
        }
        bar();
        const bool overflow = ctl[3] != 0;
        if (overflow) {
            if (p.overflow_list) {
                if (tid == 0) p.overflow_list[atomicAdd(p.n_overflow, 1)] = c;
            } else if (tid == 0) {
                p.status[b] = TDA_ST_INTERNAL | (p.nanflag[c] ? TDA_ST_NAN_INPUT : 0);
                p.counts[2 * b + 1] = 0;
            }
            bar();
            return;
        }
        // ---- classes still alive are essential
        if (tid == 0) {
            int k = ctl[2];
            for (int w = 0; w < W; ++w) {
                uint32_t bits = live[w];
                while (bits) {
                    const int s = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (k < capR) {
                        rec[k] = brank[32 * w + s];
                        rec[capR + k] = kEssential;
                        rec[2 * capR + k] = kEssential;
                    }
                    ++k;
                }
            }
            ctl[2] = k;
        }
        bar();
        const int n1 = ctl[2];
        if (n1 > capR) {
            if (tid == 0) {
                p.status[b] = TDA_ST_INTERNAL | (p.nanflag[c] ? TDA_ST_NAN_INPUT : 0);
                p.counts[2 * b + 1] = 0;
            }
            bar();
            return;
        }
        // ---- H1 rows in ripser's order: descending birth rank
        for (int k = tid; k < n1; k += NTH) {
            const uint32_t br = rec[k];
            int pos = 0;
            for (int t = 0; t < n1; ++t) pos += rec[t] > br;
            if (pos < p.cap1) {
                const size_t o = ((size_t)b * p.cap1 + pos) * 2;
                const uint32_t dr = rec[capR + k], tr = rec[2 * capR + k];
                const uint32_t pb = __ldg(P + br);
                p.bd1[o] = dist(p_i(pb), p_j(pb));
                float dth = __int_as_float(0x7F800000);
                if (tr != kEssential) { const uint32_t pd = __ldg(P + dr); dth = dist(p_i(pd), p_j(pd)); }
                p.bd1[o + 1] = dth;
                if (p.pr1) {
                    p.pr1[o] = c2(p_i(pb)) + p_j(pb);
                    p.pr1[o + 1] = (tr == kEssential) ? -1ll : (long long)tr;
                }
            }
        }
        if (tid == 0) {
            p.counts[2 * b + 1] = n1;
            p.status[b] = (p.nanflag[c] ? TDA_ST_NAN_INPUT : 0) | (n1 > p.cap1 ? TDA_ST_H1_TRUNCATED : 0);
        }
        bar();
    }
};

template <int NTH, int APT, int W, bool SG, typename TT>
__global__ void __launch_bounds__(NTH) sweep_kernel(Params p, int rezero_q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sweep<NTH, APT, W, SG, TT> s;
    uint32_t* base = (uint32_t*)smem_raw;
    s.hot = base;            base += kMaxN / 32;
    s.live = base;           base += W;
    s.used = base;           base += W;
    s.brank = base;          base += 32 * W;
    s.bstart = base;         base += 32 * W;
    s.cv = base;             base += W;
    s.peval = base;          base += W;
    s.pub = base;            base += 4;
    s.red = base;            base += 34;
    s.ctl = (int*)base;      base += 8;
    s.S = SG ? p.sglob + (size_t)blockIdx.x * p.N * W : base;
    s.capP = p.capP;
    s.capR = p.capR;
    s.phic = p.phic + (size_t)blockIdx.x * p.capP * W;
    s.pcr = p.pcr + (size_t)blockIdx.x * p.capP;
    s.act = p.act + (size_t)blockIdx.x * p.Emax;
    s.rec = p.rec + (size_t)blockIdx.x * 3 * p.capR;
    s.tid = threadIdx.x;
    s.lane = threadIdx.x & 31;
    s.warp = threadIdx.x >> 5;
    const int total = p.worklist ? *p.n_work : p.C;
    __shared__ int next_pos;
    while (true) {
        if (threadIdx.x == 0) next_pos = atomicAdd(p.queue, 1);
        __syncthreads();
        const int t = next_pos;
        __syncthreads();
        if (t >= total) break;
        const int c = p.worklist ? p.worklist[t] : (p.order ? p.order[t] : t);
        s.run(p, c, rezero_q != 0);
        __syncthreads();
    }
}

template <int NTH, int EPT, int MINB, bool NINE> static cudaError_t launch_rank_small(const Params& p, cudaStream_t st) {
    const size_t smem = RankSmall<NTH, EPT, NINE>::kSmem;
    cudaError_t e = cudaFuncSetAttribute(rank_small_kernel<NTH, EPT, MINB, NINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = p.C < kSmCount * MINB ? p.C : kSmCount * MINB;
    rank_small_kernel<NTH, EPT, MINB, NINE><<<grid, NTH, smem, st>>>(p);
    return cudaSuccess;
}

template <int W> static size_t sweep_smem(int N, bool sg) {
    size_t words = kMaxN / 32 + W + W + 32 * W + 32 * W + W + W + 4 + 34 + 8;
    if (!sg) words += (size_t)N * W;
    return words * 4;
}

// ------------------------------------------------------------------------------------ host side
struct Plan {
    int N, ldT, ib, C, grid1, grid2, nth, apt, capP, capR;
    long long Emax;
    int tbytes;   // bytes per rank of T: 2 up to 256 points, else 4
    size_t sortbuf, skey, P, T, Q, defv, m, nanflag, counters, nbirth, order, list, list0, phic1, phic2, pcr, act, rec, sglob, sorthist, lev, levR, total;
    int sortG, ntiles;   // big clouds: clouds per sort group (their sort arrays fit L2), tiles per cloud
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static int bits_for(long long v) { int b = 1; while ((1ll << b) < v) ++b; return b; }
constexpr int kGrid2 = 148;
constexpr int kSms = 148;  // B200; grids and the workspace layout are sized for it

// resident sweep CTAs per SM on the first tier (threads and shared memory)
static int tier1_ctas_per_sm(int N, int nth) {
    int by_threads = 2048 / nth;
    if (by_threads > 32) by_threads = 32;  // resident CTAs per SM
    int by_smem = (int)((227 * 1024) / ((N > 1024 ? sweep_smem<16>(N, false) : sweep_smem<8>(N, false)) + 1024));
    int r = by_threads < by_smem ? by_threads : by_smem;
    if (N > 256 && r > 2) r = 2;  // ~128 registers per thread with several apexes per thread
    if (N > 1024) r = 1;          // the tables of two resident clouds per SM (2 x 148 x 76 MB) thrash L2
    if (nth == 64 && r > 12) r = 12;  // measured: 129-256 points run best with 12 two-warp CTAs per SM
    return r < 1 ? 1 : r;
}

static bool make_plan(int B, int N, size_t ws_bytes, Plan& pl) {
    pl.N = N;
    pl.ldT = (N + 31) & ~31;
    pl.Emax = c2(N);
    pl.ib = bits_for(pl.Emax);
    pl.nth = N <= 256 ? 32 : (N <= 512 ? 256 : 512);
    pl.apt = N <= 128 ? 4 : (N <= 256 ? 8 : (N <= 1024 ? 2 : 4));
    pl.tbytes = N <= 256 ? 2 : 4;
    long long cp = 64ll * N;
    pl.capP = (int)(cp < kCapPMax ? cp : kCapPMax);
    pl.capR = (int)(pl.Emax < kCapRMax ? (pl.Emax < 64 ? 64 : pl.Emax) : kCapRMax);
    long long cmax = 1ll << 15;
    const long long cap_items = (1ll << 31) - 1;   // classify_kernel indexes (cloud, edge) pairs of a chunk
    if (cmax * pl.Emax > cap_items) cmax = cap_items / pl.Emax;
    if (cmax < 1) cmax = 1;
    const int per_sm = tier1_ctas_per_sm(N, pl.nth);
    auto layout = [&](int C) {
        size_t o = 0;
        pl.C = C;
        pl.grid1 = C < kSms * per_sm ? C : kSms * per_sm;
        pl.grid2 = C < kGrid2 ? C : kGrid2;
        pl.counters = o; o += 256;
        pl.nbirth = o; o += al((size_t)C * 4);   // (zeroed together with the counters)
        pl.order = o; o += al((size_t)C * 4);
        pl.m = o; o += al((size_t)C * 4);
        pl.nanflag = o; o += al((size_t)C * 4);
        pl.list = o; o += al((size_t)C * 4);
        pl.list0 = o; o += al((size_t)C * 4);
        // sort ping-pong of the grid-wide sort (16 bytes per edge); up to 256 points the sort lives in shared memory
        // and the region only holds the list of tie-run members left to the rank-row walk (8 bytes per edge at most)
        pl.sortbuf = o; o += al((size_t)C * pl.Emax * (N > 256 ? 16 : 8));
        pl.skey = o; o += al((size_t)C * pl.Emax * 4);
        pl.P = o; o += al((size_t)C * pl.Emax * 4);
        pl.defv = o; o += al((size_t)C * pl.Emax * 2);
        pl.T = o; o += al((size_t)C * N * pl.ldT * pl.tbytes);
        pl.Q = o; o += al((size_t)C * N * pl.ldT * 2);
        pl.phic1 = o; o += al((size_t)pl.grid1 * pl.capP * (N > 1024 ? 16 : 8) * 4);
        pl.phic2 = o; o += al((size_t)pl.grid2 * pl.capP * 32 * 4);
        pl.pcr = o; o += al((size_t)pl.grid1 * pl.capP * 4);
        pl.act = o; o += al((size_t)pl.grid1 * pl.Emax * 4);
        pl.rec = o; o += al((size_t)pl.grid1 * 3 * pl.capR * 4);
        pl.sglob = o; o += al((size_t)pl.grid2 * N * 32 * 4);
        pl.sorthist = o;
        pl.ntiles = (int)((pl.Emax + kRankTile - 1) / kRankTile);
        pl.sortG = 1;
        if (N > 256) {
            long long g = ((long long)96 << 20) / (16 * pl.Emax);   // 16 bytes of sort arrays per edge, 96 of the 126 MB of L2
            pl.sortG = (int)(g < 1 ? 1 : (g > 64 ? 64 : g));
            if (pl.sortG > C) pl.sortG = C;
            o += al((size_t)pl.sortG * pl.ntiles * 256 * 4);
            pl.lev = o; o += al((size_t)C * kLevels * N * kLevWords * 4);
            pl.levR = o; o += al((size_t)C * 2 * kLevels * 4);
        }
        pl.total = o;
    };
    long long C = B < cmax ? B : cmax;
    if (C < 1) C = 1;
    // sizing query (ws_bytes == 0): aim at <= 64 GB of the 180 GB (big clouds are latency-bound per CTA: the
    // more of them are resident at once, two per SM, the better), at least one cloud
    const size_t budget = ws_bytes ? ws_bytes : ((size_t)64 << 30);
    layout((int)C);
    while (pl.total > budget && C > 1) { C = (C + 1) / 2; layout((int)C); }
    return ws_bytes == 0 || pl.total <= ws_bytes;
}

// W0 > 0: a narrow first tier (small clouds hold few classes at once: two mask words instead of eight
// quarter the gathers and the shared memory of a visited edge); W1: the regular first tier; then W = 32
template <int NTH, int APT, int W0, int W1, typename TT>
static cudaError_t launch_sweeps(Params p, const Plan& pl, char* w8, cudaStream_t st) {
    int* counters = (int*)(w8 + pl.counters);
    cudaError_t e;
    p.phic = (uint32_t*)(w8 + pl.phic1);
    if (W0 > 0) {
        constexpr int WN = W0 > 0 ? W0 : 2;
        ProfScope prof("rips_large_sweep_t0", st);
        p.worklist = nullptr; p.n_work = nullptr;
        p.overflow_list = (int*)(w8 + pl.list0); p.n_overflow = counters + 1;
        p.queue = counters + 8;
        const size_t smem = sweep_smem<WN>(p.N, false);
        e = cudaFuncSetAttribute(sweep_kernel<NTH, APT, WN, false, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int grid = p.C < pl.grid1 ? p.C : pl.grid1;
        sweep_kernel<NTH, APT, WN, false, TT><<<grid, NTH, smem, st>>>(p, 0);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    {
        ProfScope prof("rips_large_sweep_t1", st);
        p.worklist = W0 > 0 ? (const int*)(w8 + pl.list0) : nullptr; p.n_work = W0 > 0 ? counters + 1 : nullptr;
        p.overflow_list = (int*)(w8 + pl.list); p.n_overflow = counters;
        p.queue = counters + 9;
        const size_t smem = sweep_smem<W1>(p.N, false);
        e = cudaFuncSetAttribute(sweep_kernel<NTH, APT, W1, false, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int grid = p.C < pl.grid1 ? p.C : pl.grid1;
        sweep_kernel<NTH, APT, W1, false, TT><<<grid, NTH, smem, st>>>(p, W0 > 0 ? 1 : 0);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    {
        ProfScope prof("rips_large_sweep_t2", st);
        p.worklist = (const int*)(w8 + pl.list); p.n_work = counters;
        p.overflow_list = nullptr; p.n_overflow = nullptr;
        p.phic = (uint32_t*)(w8 + pl.phic2);
        p.queue = counters + 10;
        constexpr bool SG = (NTH * APT > 1024);   // above 1,024 points S[v][32] does not fit shared memory
        const size_t smem = sweep_smem<32>(p.N, SG);
        e = cudaFuncSetAttribute(sweep_kernel<NTH, APT, 32, SG, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int grid = p.C < pl.grid2 ? p.C : pl.grid2;
        sweep_kernel<NTH, APT, 32, SG, TT><<<grid, NTH, smem, st>>>(p, 1);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace rips_large
}  // namespace tda

using namespace tda;
using namespace tda::rips_large;

extern "C" size_t tda_rips_h01_large_workspace_bytes(int B, int N) {
    if (B < 0 || N < 2 || N > kMaxN) return 0;
    Plan pl;
    make_plan(B < 1 ? 1 : B, N, 0, pl);
    return pl.total;
}

extern "C" int tda_rips_h01_large(const float* D, const int* npts, int B, int N, int ld, long long strideB,
                                  float thresh, float* bd0, long long* pr0, int cap0, float* bd1, long long* pr1,
                                  int cap1, int* counts, int* status, void* ws, size_t ws_bytes, void* stream) {
    if (!D || !bd0 || !bd1 || !counts || !status || !ws || B < 0 || cap0 < 0 || cap1 < 0 || ld < N) return TDA_E_ARG;
    if (N < 2 || N > kMaxN) return TDA_E_SIZE;
    if (B == 0) return 0;
    Plan pl;
    if (ws_bytes < 1024 || !make_plan(B, N, ws_bytes, pl)) return TDA_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* w8 = (char*)ws;
    Params p;
    p.D = D; p.npts = npts; p.strideB = strideB ? strideB : (long long)ld * ld; p.ld = ld; p.N = N; p.B = B;
    p.thresh = thresh;
    p.bd0 = bd0; p.pr0 = pr0; p.bd1 = bd1; p.pr1 = pr1; p.counts = counts; p.status = status;
    p.cap0 = cap0; p.cap1 = cap1;
    p.ldT = pl.ldT; p.ib = pl.ib; p.Emax = pl.Emax;
    p.sortbuf = (uint32_t*)(w8 + pl.sortbuf); p.skey = (uint32_t*)(w8 + pl.skey);
    p.P = (uint32_t*)(w8 + pl.P); p.T = (void*)(w8 + pl.T); p.Q = (uint16_t*)(w8 + pl.Q);
    p.defv = (uint16_t*)(w8 + pl.defv); p.m = (int*)(w8 + pl.m); p.nanflag = (int*)(w8 + pl.nanflag);
    p.capP = pl.capP; p.capR = pl.capR;
    p.phic = nullptr; p.pcr = (uint32_t*)(w8 + pl.pcr); p.act = (uint32_t*)(w8 + pl.act);
    p.rec = (uint32_t*)(w8 + pl.rec); p.sglob = (uint32_t*)(w8 + pl.sglob);
    p.worklist = nullptr; p.n_work = nullptr; p.overflow_list = nullptr; p.n_overflow = nullptr;
    p.queue = nullptr; p.order = nullptr; p.nbirth = (int*)(w8 + pl.nbirth);
    p.left = (uint2*)(w8 + pl.sortbuf);   // (the sort arrays are free once the rank kernel is done)
    p.n_left = (int*)(w8 + pl.counters) + 2;
    p.lev = N > 256 ? (const uint32_t*)(w8 + pl.lev) : nullptr;
    p.levR = N > 256 ? (const int*)(w8 + pl.levR) : nullptr;
    cudaError_t e;
    for (int c0 = 0; c0 < B; c0 += pl.C) {
        const int C = (B - c0) < pl.C ? (B - c0) : pl.C;
        p.c0 = c0; p.C = C;
        if ((e = cudaMemsetAsync(w8 + pl.counters, 0, pl.order - pl.counters, st)) != cudaSuccess) return (int)e;
        {
            // keys + per-CTA radix sort + rank matrix in one kernel (fills T, Q, m, nanflag of the chunk)
            ProfScope prof("rips_large_rank", st);
            if (N <= 128) e = launch_rank_small<512, 16, 3, true>(p, st);    // E <= 8,128: three clouds per SM
            else if (N <= 170) e = launch_rank_small<512, 32, 2, false>(p, st);   // E <= 14,365, T <= 64 KB: two (8-bit digits: the 9-bit counters would cost the second CTA)
            else if (N <= 256) e = launch_rank_small<1024, 32, 1, true>(p, st);   // E <= 32,640: 227 KB, one
            else {
                // the grid-wide sort, a group of clouds at a time (their arrays stay in L2 from the keys to the rank matrix)
                if ((e = cudaMemsetAsync(w8 + pl.m, 0, pl.list - pl.m, st)) != cudaSuccess) return (int)e;   // m, nanflag
                if ((e = cudaMemsetAsync(w8 + pl.Q, 0, (size_t)C * N * pl.ldT * 2, st)) != cudaSuccess) return (int)e;
                cudaFuncSetAttribute(scatter_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRankSmem);
                SortGroup sg;
                sg.ntiles = pl.ntiles;
                sg.hist = (uint32_t*)(w8 + pl.sorthist);
                for (int g0 = 0; g0 < C; g0 += pl.sortG) {
                    sg.g0 = g0;
                    sg.G = (C - g0) < pl.sortG ? (C - g0) : pl.sortG;
                    int gx = (2 * kSms + sg.G - 1) / sg.G;   // CTAs per cloud for the block- and rank-strided kernels
                    { const int nb = (N + 31) / 32; if (gx > nb * (nb + 1) / 2) gx = nb * (nb + 1) / 2; }
                    keys_big_kernel<<<dim3(gx, sg.G), kRankThreads, 0, st>>>(p, sg);
                    for (int pass = 0; pass < 4; ++pass) {
                        hist_big_kernel<<<dim3(pl.ntiles, sg.G), kRankThreads, 0, st>>>(p, sg, pass);
                        scatter_big_kernel<<<dim3(pl.ntiles, sg.G), kRankThreads, kRankSmem, st>>>(p, sg, pass);
                    }
                    int gf = (2 * kSms + sg.G - 1) / sg.G;
                    if (gf > pl.ntiles * 8) gf = pl.ntiles * 8;
                    finish_big_kernel<<<dim3(gf, sg.G), kRankThreads, 0, st>>>(p, sg);
                    levels_big_kernel<<<dim3(gf, sg.G), 256, 0, st>>>(p, sg, (uint32_t*)(w8 + pl.lev), (int*)(w8 + pl.levR));
                    count_launch(11);
                }
                e = cudaGetLastError();
            }
            if (e != cudaSuccess) return (int)e;
            count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        }
        {
            ProfScope prof("rips_large_kruskal", st);
            if (N > 256) {   // (up to 256 points Kruskal rides on the classification's walk of the edge list)
                kruskal_kernel<1024><<<C, 1024, 0, st>>>(p);
                count_launch();
            }
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        }
        {
            ProfScope prof("rips_large_classify", st);
            long long blocks = (pl.Emax * C + 255) / 256;
            if (blocks > 148 * 64) blocks = 148 * 64;
            if (N > 256) {
                classify_kernel<uint32_t, false><<<(unsigned)blocks, 256, 0, st>>>(p);
                count_launch();
            } else {
                // a warp per cloud; 9.7 KB of shared memory per warp at 129-256 points: two-warp CTAs, 11 per SM
                const int nwords = N <= 128 ? 4 : 8, wpb = N <= 128 ? 8 : 2;
                const size_t smem = (size_t)wpb * (32 * nwords * nwords + 32 * nwords + 16 * nwords) * 4;
                int g = (C + wpb - 1) / wpb;
                const int gmax = kSms * (N <= 128 ? 8 : 11);
                if (g > gmax) g = gmax;
                if (N <= 128) classify_bits_kernel<4><<<g, 32 * wpb, smem, st>>>(p);
                else classify_bits_kernel<8><<<g, 32 * wpb, smem, st>>>(p);
                count_launch();
                if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
                ProfScope prof2("rips_large_classify_tied", st);
                classify_kernel<uint16_t, true><<<kSms * 4, 256, 0, st>>>(p);   // tie runs across a step boundary
                count_launch();
            }
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        }
        p.order = nullptr;
        if (C > 1 && C <= 8192) {   // heaviest clouds first (quadratic in the clouds of the chunk: 0.05 ms at 8,192)
            order_kernel<<<(C + 255) / 256, 256, 0, st>>>(p, (int*)(w8 + pl.order));
            count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
            p.order = (const int*)(w8 + pl.order);
        }
        // small clouds: one or two warps per cloud (the visited edges are issue-bound there); above 256 points 256
        // threads, above 512 points 512 threads (measured against 256 and 1,024: a visited edge is a chain of
        // global round trips and block barriers, more threads shorten the row scans, more warps lengthen the
        // barriers), 2-4 apexes per thread
        if (N <= 128) e = launch_sweeps<32, 4, 2, 8, uint16_t>(p, pl, w8, st);
        else if (N <= 256) e = launch_sweeps<32, 8, 2, 8, uint16_t>(p, pl, w8, st);
        else if (N <= 512) e = launch_sweeps<256, 2, 0, 8, uint32_t>(p, pl, w8, st);
        else if (N <= 1024) e = launch_sweeps<512, 2, 0, 8, uint32_t>(p, pl, w8, st);
        else e = launch_sweeps<512, 4, 0, 16, uint32_t>(p, pl, w8, st);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
