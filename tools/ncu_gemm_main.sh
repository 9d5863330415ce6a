#!/bin/bash
# ncu --set full capture of main-path GEMM launches (level-1 ResBlock / attention projections) of one denoiser forward
tag=${1:-r01}
mkdir -p gpurun_out
timeout 500 ncu --profile-from-start off --set full --import-source on --clock-control none -f \
    -k regex:gemm_f16_cg2 --launch-skip 110 -c 16 -o gpurun_out/${tag}_gemm_main python tools/profile_step.py --sample-steps 1 \
    > gpurun_out/${tag}_gemm_main.log 2>&1
echo "gemm_main rc=$?"
