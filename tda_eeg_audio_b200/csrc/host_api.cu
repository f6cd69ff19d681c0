// host_api.cu — host-pointer front ends (the "e2e" path): chunked H2D -> kernels -> D2H
// pipelines over three streams so the PCIe copies of chunk k+1 / k-1 hide behind the kernels of
// chunk k.  Staging buffers are cached per process (the only state the library keeps).
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "common.cuh"
#include "tda_b200.h"

namespace {

struct DevBuf {
    void* dev = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (dev) cudaFree(dev);
        dev = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&dev, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
};
struct Stage {
    DevBuf buf;
    cudaStream_t stream = nullptr;
    cudaError_t ensure(size_t need) {
        if (!stream) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        return buf.ensure(need);
    }
};
constexpr int kStages = 3;
Stage g_stage[kStages];
DevBuf g_feats, g_table;
std::mutex g_mu;

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

// Shared body.  feats_dev != nullptr => per-window features are computed into it (B,2,11).
// condensed: D holds N(N-1)/2 floats per window (upper triangle, row-major) instead of N x N.
int run_host(const float* D, bool condensed, int B, int N, float thresh, float* bd0, long long* pr0, float* bd1,
             long long* pr1, int* counts, int cap1, int* status, double* feats_dev, double* feats_host) {
    int chunk = 32768;
    if (chunk > B) chunk = B;
    const size_t inE = condensed ? (size_t)N * (N - 1) / 2 : (size_t)N * N;   // floats per window
    const size_t inB = inE * 4;
    const size_t o_bd0 = (size_t)N * 2 * 4, o_pr0 = (size_t)N * 2 * 8;
    const size_t o_bd1 = (size_t)cap1 * 2 * 4, o_pr1 = (size_t)cap1 * 2 * 8;
    const size_t wsB = tda_rips_h01_workspace_bytes(chunk, N);
    size_t off = 0;
    const size_t f_in = off;   off += al(inB * chunk);
    const size_t f_bd0 = off;  off += al(o_bd0 * chunk);
    const size_t f_pr0 = off;  off += al(pr0 ? o_pr0 * chunk : 0);
    const size_t f_bd1 = off;  off += al(o_bd1 * chunk + 16);
    const size_t f_pr1 = off;  off += al(pr1 ? o_pr1 * chunk : 0);
    const size_t f_cnt = off;  off += al((size_t)chunk * 8);
    const size_t f_st = off;   off += al((size_t)chunk * 4);
    const size_t f_ws = off;   off += al(wsB);
    cudaError_t e;
    for (int s = 0; s < kStages; ++s) {
        e = g_stage[s].ensure(off);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = 0;
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++k) {
        const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
        Stage& S = g_stage[k % kStages];
        char* d = (char*)S.buf.dev;
        cudaStream_t st = S.stream;
        // a stage is reused only after everything queued on its stream is done (stream order)
        e = cudaMemcpyAsync(d + f_in, D + (size_t)b0 * inE, inB * nb, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = (int)e; break; }
        rc = tda_rips_h01_batched((const float*)(d + f_in), nb, N, condensed ? 0 : N, 0, thresh, (float*)(d + f_bd0),
                                  pr0 ? (long long*)(d + f_pr0) : nullptr, (float*)(d + f_bd1),
                                  pr1 ? (long long*)(d + f_pr1) : nullptr, (int*)(d + f_cnt), cap1,
                                  (int*)(d + f_st), d + f_ws, wsB, st);
        if (rc != 0) break;
        if (feats_dev) {
            double* fo = feats_dev + (size_t)b0 * 22;
            rc = tda_pers_features((const float*)(d + f_bd0), N, (const int*)(d + f_cnt), 2, nb, fo, 22, st);
            if (rc != 0) break;
            rc = tda_pers_features((const float*)(d + f_bd1), cap1, (const int*)(d + f_cnt) + 1, 2, nb, fo + 11, 22, st);
            if (rc != 0) break;
            if (feats_host)
                cudaMemcpyAsync(feats_host + (size_t)b0 * 22, fo, (size_t)nb * 22 * 8, cudaMemcpyDeviceToHost, st);
        }
        if (bd0) cudaMemcpyAsync(bd0 + (size_t)b0 * N * 2, d + f_bd0, o_bd0 * nb, cudaMemcpyDeviceToHost, st);
        if (pr0) cudaMemcpyAsync(pr0 + (size_t)b0 * N * 2, d + f_pr0, o_pr0 * nb, cudaMemcpyDeviceToHost, st);
        if (cap1 > 0 && bd1) {
            cudaMemcpyAsync(bd1 + (size_t)b0 * cap1 * 2, d + f_bd1, o_bd1 * nb, cudaMemcpyDeviceToHost, st);
            if (pr1) cudaMemcpyAsync(pr1 + (size_t)b0 * cap1 * 2, d + f_pr1, o_pr1 * nb, cudaMemcpyDeviceToHost, st);
        }
        if (counts) cudaMemcpyAsync(counts + (size_t)b0 * 2, d + f_cnt, (size_t)nb * 8, cudaMemcpyDeviceToHost, st);
        if (status) cudaMemcpyAsync(status + b0, d + f_st, (size_t)nb * 4, cudaMemcpyDeviceToHost, st);
        e = cudaGetLastError();
        if (e != cudaSuccess) { rc = (int)e; break; }
    }
    for (int s = 0; s < kStages; ++s) {
        if (g_stage[s].stream) {
            e = cudaStreamSynchronize(g_stage[s].stream);
            if (e != cudaSuccess && rc == 0) rc = (int)e;
        }
    }
    return rc;
}

}  // namespace

namespace {
int rips_host(const float* D, bool condensed, int B, int N, float thresh, float* bd0, long long* pr0, float* bd1,
              long long* pr1, int* counts, int cap1, int* status, int device) {
    if (!D || !bd0 || !bd1 || !counts || !status || B < 0 || cap1 < 0) return TDA_E_ARG;
    if (N < 2 || N > 64) return TDA_E_SIZE;
    if (B == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mu);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    return run_host(D, condensed, B, N, thresh, bd0, pr0, bd1, pr1, counts, cap1, status, nullptr, nullptr);
}

int features_host(const float* D, bool condensed, int R, int Bd, int Wn, int N, float thresh, int cap1, float* bd0,
                  float* bd1, int* counts, int* status, double* feats, double* table, int device) {
    if (!D || !table || R < 0 || Bd < 0 || Wn < 0 || cap1 < 1) return TDA_E_ARG;
    if (N < 2 || N > 64) return TDA_E_SIZE;
    const long long Bll = (long long)R * Bd * Wn;
    if (Bll > 0x7FFFFFFF) return TDA_E_SIZE;
    const int B = (int)Bll;
    if (B == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mu);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    e = g_feats.ensure((size_t)B * 22 * 8);
    if (e != cudaSuccess) return (int)e;
    e = g_table.ensure((size_t)R * Bd * 44 * 8);
    if (e != cudaSuccess) return (int)e;
    int rc = run_host(D, condensed, B, N, thresh, bd0, nullptr, bd1, nullptr, counts, cap1, status,
                      (double*)g_feats.dev, feats);
    if (rc != 0) return rc;
    cudaStream_t st = g_stage[0].stream;
    rc = tda_aggregate_windows((const double*)g_feats.dev, R, Bd, Wn, (double*)g_table.dev, st);
    if (rc != 0) return rc;
    e = cudaMemcpyAsync(table, g_table.dev, (size_t)R * Bd * 44 * 8, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize(st);
}
}  // namespace

extern "C" int tda_rips_h01_host(const float* D, int B, int N, float thresh, float* bd0, long long* pr0,
                                 float* bd1, long long* pr1, int* counts, int cap1, int* status, int device) {
    return rips_host(D, false, B, N, thresh, bd0, pr0, bd1, pr1, counts, cap1, status, device);
}

extern "C" int tda_rips_h01_condensed_host(const float* Dc, int B, int N, float thresh, float* bd0, long long* pr0,
                                           float* bd1, long long* pr1, int* counts, int cap1, int* status,
                                           int device) {
    return rips_host(Dc, true, B, N, thresh, bd0, pr0, bd1, pr1, counts, cap1, status, device);
}

extern "C" int tda_eeg_features_host(const float* D, int R, int Bd, int Wn, int N, float thresh, int cap1,
                                     float* bd0, float* bd1, int* counts, int* status, double* feats,
                                     double* table, int device) {
    return features_host(D, false, R, Bd, Wn, N, thresh, cap1, bd0, bd1, counts, status, feats, table, device);
}

extern "C" int tda_eeg_features_condensed_host(const float* Dc, int R, int Bd, int Wn, int N, float thresh,
                                               int cap1, float* bd0, float* bd1, int* counts, int* status,
                                               double* feats, double* table, int device) {
    return features_host(Dc, true, R, Bd, Wn, N, thresh, cap1, bd0, bd1, counts, status, feats, table, device);
}
