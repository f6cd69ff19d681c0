#!/bin/bash
# One GPU call for everything that is staged behind an environment switch:
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/run_staged_ab.sh'
# 1. times the generations of the IIR pass kernel and compares their outputs bit for bit,
# 2. times the FP64-tensor-pipe Gram against the FMA one and reports their largest differences,
# 3. runs the signal-side parity suites (scipy / numpy / reference fixtures) WITH the staged kernels on.
# Adopt a staged kernel (flip its default in csrc/iir.cu / csrc/corrdist.cu) only if 1-2 show a gain
# and 3 is green.
set -u
mkdir -p gpurun_out
timeout 90 python tools/ab_iir.py 0,1,3 > gpurun_out/staged_ab_iir.jsonl 2> gpurun_out/staged_ab_iir.err
cat gpurun_out/staged_ab_iir.jsonl; tail -2 gpurun_out/staged_ab_iir.err
timeout 90 python tools/ab_corrdist.py 128 > gpurun_out/staged_ab_corrdist.jsonl 2> gpurun_out/staged_ab_corrdist.err
cat gpurun_out/staged_ab_corrdist.jsonl; tail -2 gpurun_out/staged_ab_corrdist.err
TDA_IIR=3 TDA_CORRDIST=mma timeout 120 python -m pytest tests/test_dsp_gpu.py tests/test_audio_gpu.py \
    tests/test_drivers_gpu.py tests/test_coupling_gpu.py -m gpu -q 2>&1 | tail -5
