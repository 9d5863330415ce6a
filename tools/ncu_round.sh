#!/bin/bash
# Profiling round on the GPU box (one GPU):
#   1. launch list of one bench step (ncu --metrics gpu__time_duration.sum, cold-cache serialised launches)
#   2. ncu --set full captures of the hot kernels inside one denoiser forward
#   bash tools/ncu_round.sh <tag>  ->  gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches.csv python tools/profile_step.py > $out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
run() {  # name, kernel regex (demangled), skip, count
    timeout 400 $NCU -k "regex:$2" --launch-skip $3 -c $4 -o $out/${tag}_$1 python tools/profile_step.py --sample-steps 1 \
        > $out/${tag}_$1.log 2>&1
    echo "$1 rc=$?"
}
run attention "attention_kernel" 14 4
run dwconv "dwconv" 4 4
run grn "grn_fused" 4 3
run ln "ln_rows_kernel<.bool.0>" 8 2
run gemm_resid "cg2_kernel<.int.256, .int.3>" 4 4
run gemm_gelu "cg2_kernel<.int.256, .int.2>" 4 4
run gemm_f16 "cg2_kernel<.int.256, .int.0>" 100 2
run sampler "fused_sampler" 0 1
gzip -f $out/${tag}_launches.csv
ls -la $out | tail -30
