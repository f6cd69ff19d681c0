// Microbenchmark: what the FP64 pipe of this GPU sustains for non-fused DMUL / DADD streams, as a
// function of independent chains per thread (ILP) and resident warps per SM.  The roofline denominator
// of the zero-phase IIR kernel (which may not use FMA).   nvcc -arch=sm_100a -O3 fp64_rate.cu -o fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool FMA>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (FMA) x[i] = fma(x[i], a, b);
            else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
template <int ILP, bool FMA> void run(int warps_per_sm, double* d) {
    int sms = 148, threads = 128, blocks = sms * warps_per_sm / 4, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, FMA><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP, FMA><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)blocks * threads * iters * ILP * (FMA ? 1 : 2);
    printf("{\"fma\": %d, \"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"T_fp64_instr_per_s\": %.3f}\n", (int)FMA, ILP, warps_per_sm, ms, instr / ms / 1e9);
}
int main() {
    double* d; cudaMalloc(&d, 8);
    for (int w : {4, 8, 16, 32}) { run<1, false>(w, d); run<2, false>(w, d); run<4, false>(w, d); run<8, false>(w, d); }
    for (int w : {16, 32}) { run<4, true>(w, d); run<8, true>(w, d); }
    return 0;
}
