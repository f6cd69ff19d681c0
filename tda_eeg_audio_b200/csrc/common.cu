// common.cu — version / launch accounting / per-kernel event timing for libtda_b200.so
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

struct Sample {
    std::string name;
    cudaEvent_t e0, e1;
};
static std::atomic<int> g_prof{0};
static std::mutex g_prof_mu;
static std::vector<Sample> g_samples;

ProfScope::ProfScope(const char* name, cudaStream_t st) : name_(name), st_(st) {
    if (!g_prof.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&e0_) != cudaSuccess) { e0_ = nullptr; return; }
    cudaEventRecord(e0_, st_);
}
ProfScope::~ProfScope() {
    if (!e0_) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0_); return; }
    cudaEventRecord(e1, st_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_samples.push_back(Sample{name_, e0_, e1});
}
}  // namespace tda

extern "C" int tda_version(void) { return 100; }
extern "C" unsigned long long tda_launch_count(void) {
    return tda::g_launches.load(std::memory_order_relaxed);
}

extern "C" int tda_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(tda::g_prof_mu);
    for (auto& s : tda::g_samples) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
    tda::g_samples.clear();
    tda::g_prof.store(on ? 1 : 0);
    return 0;
}

extern "C" int tda_profile_query(const char* kernel, double* total_ms, int* launches) {
    if (!kernel || !total_ms || !launches) return TDA_E_ARG;
    std::lock_guard<std::mutex> lk(tda::g_prof_mu);
    double tot = 0;
    int n = 0;
    for (auto& s : tda::g_samples) {
        if (s.name != kernel) continue;
        cudaError_t e = cudaEventSynchronize(s.e1);
        if (e != cudaSuccess) return (int)e;
        float ms = 0;
        e = cudaEventElapsedTime(&ms, s.e0, s.e1);
        if (e != cudaSuccess) return (int)e;
        tot += ms;
        ++n;
    }
    *total_ms = tot;
    *launches = n;
    return 0;
}
