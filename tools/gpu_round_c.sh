#!/bin/bash
# Round C: attention kernel iteration (tests, bench A/B against the mma.sync kernel, in-kernel timeline), 64x64-latent workload A/B
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
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
timeout 400 python -m pytest tests/test_gpu_attention.py -q --no-header -p no:cacheprovider 2>&1 | tail -25
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; show $O/${TAG}_bench_default.json "default (transposed tcgen05 attention)"
PB200_ATTN_NO_TT=1 timeout 300 $B > $O/${TAG}_bench_tc.json 2> /dev/null; show $O/${TAG}_bench_tc.json "row-major tcgen05 attention"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_legacy.json 2> /dev/null; show $O/${TAG}_bench_legacy.json "attention legacy (mma.sync)"
PB200_TRACE=attention_tt:$O/${TAG}_trace_attn.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
S="python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 400 $S > $O/${TAG}_sample64.json 2> /dev/null; show $O/${TAG}_sample64.json "sample64 default"
PB200_ATTN_LEGACY=1 timeout 400 $S > $O/${TAG}_sample64_legacy.json 2> /dev/null; show $O/${TAG}_sample64_legacy.json "sample64 attention legacy"
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
timeout 400 $NCU -k "regex:attention_tt_kernel" --launch-skip 14 -c 3 -o /tmp/${TAG}_attn python tools/profile_step.py --sample-steps 1 > $O/${TAG}_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_attn.ncu-rep > $O/${TAG}_ncu_attention.md 2>&1; cat $O/${TAG}_ncu_attention.md | cut -c1-330
du -sh $O
