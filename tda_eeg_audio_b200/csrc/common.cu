// common.cu — version / launch accounting for libtda_b200.so
#include <atomic>

#include "common.cuh"
#include "tda_b200.h"

namespace tda {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace tda

extern "C" int tda_version(void) { return 100; }
extern "C" unsigned long long tda_launch_count(void) {
    return tda::g_launches.load(std::memory_order_relaxed);
}
