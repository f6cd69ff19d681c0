// host_api.cu — host-pointer front ends (the "e2e" path): chunked H2D -> kernels -> D2H
// pipelines over three streams so the PCIe copies of chunk k+1 / k-1 hide behind the kernels of
// chunk k.  Staging buffers and streams are cached per device ordinal (the only state the library
// keeps); calls are serialised by one mutex.
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
constexpr int kMaxDevices = 64;
struct DeviceState {
    Stage stage[kStages];
    DevBuf feats, table;
};
DeviceState* g_dev[kMaxDevices];   // created on first use of an ordinal; device memory belongs to that device
std::mutex g_mu;

// the state of CUDA ordinal `device`, which becomes the current device
int device_state(int device, DeviceState** out) {
    if (device < 0 || device >= kMaxDevices) return TDA_E_ARG;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    if (!g_dev[device]) g_dev[device] = new DeviceState();
    *out = g_dev[device];
    return 0;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

// Shared body.  feats_dev != nullptr => per-window features are computed into it (B,2,11).
// condensed: D holds N(N-1)/2 floats per window (upper triangle, row-major) instead of N x N.
// in_f64: D holds N x N float64 per window (the reference's own argument type); the chunk is
// symmetrised, clamped and cast on the device (tda_symmetrize_f64_to_f32) before the Rips kernels.
int run_host(DeviceState& G, const void* D, bool condensed, bool in_f64, int B, int N, float thresh, float* bd0, long long* pr0, float* bd1,
             long long* pr1, int* counts, int cap1, int* status, double* feats_dev, double* feats_host) {
    int chunk = 32768;
    if (chunk > B) chunk = B;
    const size_t inE = condensed ? (size_t)N * (N - 1) / 2 : (size_t)N * N;   // values per window
    const size_t inB = inE * (in_f64 ? 8 : 4);
    const size_t o_bd0 = (size_t)N * 2 * 4, o_pr0 = (size_t)N * 2 * 8;
    const size_t o_bd1 = (size_t)cap1 * 2 * 4, o_pr1 = (size_t)cap1 * 2 * 8;
    const size_t wsB = tda_rips_h01_workspace_bytes(chunk, N);
    size_t off = 0;
    const size_t f_in = off;   off += al(inB * chunk);
    const size_t f_in32 = off; off += al(in_f64 ? (size_t)N * N * 4 * chunk : 0);
    const size_t f_bd0 = off;  off += al(o_bd0 * chunk);
    const size_t f_pr0 = off;  off += al(pr0 ? o_pr0 * chunk : 0);
    const size_t f_bd1 = off;  off += al(o_bd1 * chunk + 16);
    const size_t f_pr1 = off;  off += al(pr1 ? o_pr1 * chunk : 0);
    const size_t f_cnt = off;  off += al((size_t)chunk * 8);
    const size_t f_st = off;   off += al((size_t)chunk * 4);
    const size_t f_ws = off;   off += al(wsB);
    cudaError_t e;
    for (int s = 0; s < kStages; ++s) {
        e = G.stage[s].ensure(off);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = 0;
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++k) {
        const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
        Stage& S = G.stage[k % kStages];
        char* d = (char*)S.buf.dev;
        cudaStream_t st = S.stream;
        // a stage is reused only after everything queued on its stream is done (stream order)
        e = cudaMemcpyAsync(d + f_in, (const char*)D + (size_t)b0 * inB, inB * nb, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = (int)e; break; }
        const float* din = (const float*)(d + f_in);
        if (in_f64) {
            rc = tda_symmetrize_f64_to_f32((const double*)(d + f_in), nb, N, (float*)(d + f_in32), st);
            if (rc != 0) break;
            din = (const float*)(d + f_in32);
        }
        rc = tda_rips_h01_batched(din, nb, N, condensed ? 0 : N, 0, thresh, (float*)(d + f_bd0),
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
            if (feats_host) {
                e = cudaMemcpyAsync(feats_host + (size_t)b0 * 22, fo, (size_t)nb * 22 * 8, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) { rc = (int)e; break; }
            }
        }
        e = cudaSuccess;
        auto d2h = [&](void* dst, const void* src, size_t bytes) {
            if (e == cudaSuccess && dst && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        };
        d2h(bd0 ? bd0 + (size_t)b0 * N * 2 : nullptr, d + f_bd0, o_bd0 * nb);
        d2h(pr0 ? pr0 + (size_t)b0 * N * 2 : nullptr, d + f_pr0, o_pr0 * nb);
        d2h(bd1 ? bd1 + (size_t)b0 * cap1 * 2 : nullptr, d + f_bd1, o_bd1 * nb);
        d2h(pr1 ? pr1 + (size_t)b0 * cap1 * 2 : nullptr, d + f_pr1, o_pr1 * nb);
        d2h(counts ? counts + (size_t)b0 * 2 : nullptr, d + f_cnt, (size_t)nb * 8);
        d2h(status ? status + b0 : nullptr, d + f_st, (size_t)nb * 4);
        if (e != cudaSuccess) { rc = (int)e; break; }
        e = cudaGetLastError();
        if (e != cudaSuccess) { rc = (int)e; break; }
    }
    for (int s = 0; s < kStages; ++s) {
        if (G.stage[s].stream) {
            e = cudaStreamSynchronize(G.stage[s].stream);
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
    DeviceState* G = nullptr;
    if (int rc = device_state(device, &G)) return rc;
    return run_host(*G, D, condensed, false, B, N, thresh, bd0, pr0, bd1, pr1, counts, cap1, status, nullptr, nullptr);
}

int features_host(const void* D, bool condensed, bool in_f64, int R, int Bd, int Wn, int N, float thresh, int cap1, float* bd0,
                  float* bd1, int* counts, int* status, double* feats, double* table, int device) {
    if (!D || !table || R < 0 || Bd < 0 || Wn < 0 || cap1 < 1) return TDA_E_ARG;
    if (N < 2 || N > 64) return TDA_E_SIZE;
    const long long Bll = (long long)R * Bd * Wn;
    if (Bll > 0x7FFFFFFF) return TDA_E_SIZE;
    const int B = (int)Bll;
    if (B == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceState* G = nullptr;
    if (int rc0 = device_state(device, &G)) return rc0;
    cudaError_t e = G->feats.ensure((size_t)B * 22 * 8);
    if (e != cudaSuccess) return (int)e;
    e = G->table.ensure((size_t)R * Bd * 44 * 8);
    if (e != cudaSuccess) return (int)e;
    int rc = run_host(*G, D, condensed, in_f64, B, N, thresh, bd0, nullptr, bd1, nullptr, counts, cap1, status,
                      (double*)G->feats.dev, feats);
    if (rc != 0) return rc;
    cudaStream_t st = G->stage[0].stream;
    rc = tda_aggregate_windows((const double*)G->feats.dev, R, Bd, Wn, (double*)G->table.dev, st);
    if (rc != 0) return rc;
    e = cudaMemcpyAsync(table, G->table.dev, (size_t)R * Bd * 44 * 8, cudaMemcpyDeviceToHost, st);
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
    return features_host(D, false, false, R, Bd, Wn, N, thresh, cap1, bd0, bd1, counts, status, feats, table, device);
}

extern "C" int tda_eeg_features_condensed_host(const float* Dc, int R, int Bd, int Wn, int N, float thresh,
                                               int cap1, float* bd0, float* bd1, int* counts, int* status,
                                               double* feats, double* table, int device) {
    return features_host(Dc, true, false, R, Bd, Wn, N, thresh, cap1, bd0, bd1, counts, status, feats, table, device);
}

extern "C" int tda_eeg_features_f64_host(const double* D64, int R, int Bd, int Wn, int N, float thresh, int cap1,
                                         float* bd0, float* bd1, int* counts, int* status, double* feats,
                                         double* table, int device) {
    return features_host(D64, false, true, R, Bd, Wn, N, thresh, cap1, bd0, bd1, counts, status, feats, table, device);
}
