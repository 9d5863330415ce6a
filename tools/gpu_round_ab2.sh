#!/bin/bash
# A/B of the opt-in kernels after the elect.sync change (their earlier losses were measured with ~100-cycle issue loops per tcgen05.mma)
O=gpurun_out; mkdir -p $O; TAG=${1:-r2r}
show() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d["roofline"]["families"].items()}
    print(sys.argv[2] + ":", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms/step", "gemm-frac", round(d["roofline"]["frac"], 3), fam)
except Exception as e:
    print(sys.argv[2] + ": FAILED", e)
PY
}
V="python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $V > $O/${TAG}_vqgan64.json 2>/dev/null; show $O/${TAG}_vqgan64.json "vqgan bs=64 default"
PB200_VQ_MLP_FUSED=1 timeout 300 $V > $O/${TAG}_vqgan64_mlpfused.json 2>/dev/null; show $O/${TAG}_vqgan64_mlpfused.json "vqgan bs=64 fused MLP"
S="python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $S > $O/${TAG}_sample64.json 2>/dev/null; show $O/${TAG}_sample64.json "sample64 default"
PB200_ATTN_TC_WIDE=1 timeout 300 $S > $O/${TAG}_sample64_tcwide.json 2>/dev/null; show $O/${TAG}_sample64_tcwide.json "sample64 tcgen05 attention at 392 keys"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $B > $O/${TAG}_bench.json 2>/dev/null; show $O/${TAG}_bench.json "sample default"
PB200_GRN_FOLD=1 timeout 300 $B > $O/${TAG}_bench_grnfold.json 2>/dev/null; show $O/${TAG}_bench_grnfold.json "sample GRN fold"
PB200_ATTN_NO_TT=1 timeout 300 $B > $O/${TAG}_bench_attn_tc.json 2>/dev/null; show $O/${TAG}_bench_attn_tc.json "sample row-major tcgen05 attention"
