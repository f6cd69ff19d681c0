// common.cuh — shared helpers of libtda_b200.so (host side bookkeeping only).
#pragma once
#include <cuda_runtime.h>

namespace tda {
// every kernel launch made by the library goes through here so that
// tda_launch_count() (bench.py's "gpu_launches") is a count, not a guess
void count_launch(int n = 1);

// Optional per-kernel timing with CUDA events on the launching stream
// (tda_profile_enable / tda_profile_query).  Costs nothing when disabled.
struct ProfScope {
    ProfScope(const char* name, cudaStream_t st);
    ~ProfScope();
    const char* name_;
    cudaStream_t st_;
    cudaEvent_t e0_ = nullptr;
};
}  // namespace tda
