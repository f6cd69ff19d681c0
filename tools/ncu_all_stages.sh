#!/bin/bash
# One `ncu --set full` capture per kernel family north_star names, each on its stage workload
# (tools/profile_stage.py), reports into gpurun_out/ (bring them back, summarise with
# tools/ncu_summary.py into profiles/).  Each plain run must exit 0 before its ncu run.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/ncu_all_stages.sh'
set -u
mkdir -p gpurun_out
run() {  # tag stage kernel-regex skip
    local tag=$1 stage=$2 regex=$3 skip=$4
    timeout 300 python tools/profile_stage.py $stage > gpurun_out/stage_$tag.log 2>&1 || { echo "$tag: plain run failed"; tail -3 gpurun_out/stage_$tag.log; return; }
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f \
        -o gpurun_out/r02_$tag python tools/profile_stage.py $stage > gpurun_out/ncu_$tag.log 2>&1
    echo "$tag: rc=$? $(ls -la gpurun_out/r02_$tag.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
run iir_forward   iir         'iir_pass_kernel'        2
run iir_backward  iir         'iir_pass_kernel'        3
run corrdist      corrdist    'corrdist_mma_kernel'    1
run features      features    'pers_features_kernel'   4
run wasserstein_h0 wasserstein 'wasserstein_kernel'    2
run wasserstein_h1 wasserstein 'wasserstein_kernel'    3
run takens_dist   takens      'pairwise_kernel'        1
run takens_cloud  takens      'takens_kernel'          1
run resample      resample    'resample_kernel'        1
