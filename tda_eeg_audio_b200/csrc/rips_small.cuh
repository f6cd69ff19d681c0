// rips_small.cuh — declarations of the N <= 64 Rips engine (rips_small.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tda {
namespace rips_small {

constexpr int kMaxN = 64;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kEssential = 0xFFFFFFFFu;
// P[rank] = j | i << 6 | flags
constexpr uint32_t kMst = 1u << 12;      // edge merges two components (H0 death)
constexpr uint32_t kTie = 1u << 13;      // the next edge in the order has the same length
constexpr uint32_t kTiePrev = 1u << 14;  // the previous edge in the order has the same length
__device__ __forceinline__ int p_i(uint32_t p) { return (p >> 6) & 63; }
__device__ __forceinline__ int p_j(uint32_t p) { return p & 63; }

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
    uint32_t* rec_global;  // per-warp record scratch (PHI_GLOBAL tiers), 4*R words per warp
    uint8_t* defv_global;  // per-warp tie-run scratch, Epad bytes per warp (all tiers)
};

__host__ __device__ inline int c2(int i) { return i * (i - 1) / 2; }
__host__ __device__ inline int c3(int i) { return i * (i - 1) * (i - 2) / 6; }

// Offset of row `row` of a window's distance matrix, to which the column i > row is added.  ld >= N:
// dense row-major matrix.  ld == 0: the condensed upper triangle in row-major order (0,1),(0,2),...,
// (1,2),... -- the float32 vector ripser.py hands its C++ core (DParam).
__device__ __forceinline__ long long d_rowoff(int row, int ld, int N) {
    return ld ? (long long)row * ld : (long long)(row * (2 * N - row - 1) / 2 - row - 1);
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
__device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t u = (k >> 31) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ int tri_index(int x, int y, int z) {
    const int a = max(x, max(y, z)), c = min(x, min(y, z)), b = x + y + z - a - c;
    return c3(a) + c2(b) + c;
}

}  // namespace rips_small
}  // namespace tda
