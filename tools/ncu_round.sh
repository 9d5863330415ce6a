#!/bin/bash
# ncu captures of the hot kernels inside one sample() step of the bench workload (run on the GPU box):
#   bash tools/ncu_round.sh <tag>      -> gpurun_out/<tag>_*.ncu-rep  (read locally with `ncu -i ... --page raw --csv`)
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f"
run() {  # name, kernel regex, skip, count, [env]
    env $5 timeout 400 $NCU -k regex:$2 --launch-skip $3 -c $4 -o $out/${tag}_$1 python tools/profile_step.py --sample-steps 1 > $out/${tag}_$1.log 2>&1
    echo "$1 rc=$?"
}
run attention attention_kernel 14 4
run dwconv_patch dwconv 4 4
run dwconv_warp dwconv 4 4 PB200_DWCONV_WARP=1
run grn grn_fused 4 4
run gemm_cg2 gemm_f16_cg2 24 12
run sampler fused_sampler 0 1
