#!/bin/bash
# ncu --set full capture of main-path GEMM launches of one denoiser forward, selected by epilogue mode
# (3 = RESID+FiLM, 2 = GELU+GRN statistic, 0 = F16) through the demangled template arguments
tag=${1:-r01}
mkdir -p gpurun_out
for spec in "resid:cg2_kernel<.int.256, .int.3>:2:8" "gelu:cg2_kernel<.int.256, .int.2>:4:4" "f16:cg2_kernel<.int.256, .int.0>:180:2"; do
    IFS=: read name rx skip cnt <<< "$spec"
    timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none -f \
        --kernel-name-base demangled -k "regex:$rx" --launch-skip $skip -c $cnt -o gpurun_out/${tag}_gemm_$name \
        python tools/profile_step.py --sample-steps 1 > gpurun_out/${tag}_gemm_$name.log 2>&1
    echo "gemm_$name rc=$?"
done
