#!/bin/bash
# Round E: attention_tt iteration (tests, bench A/B against the mma.sync kernel, timeline, ncu).  bash tools/gpu_round_e.sh <tag>
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
show() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d["roofline"]["families"].items()}
    print(sys.argv[2] + ":", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms/step", "e2e", round(d["e2e"]["value"], 1),
          "gemm-frac", round(d["roofline"]["frac"], 3), fam)
except Exception as e:
    print(sys.argv[2] + ": FAILED", e)
PY
}
timeout 400 python -m pytest tests/test_gpu_attention.py tests/test_gpu_blocks.py -q --no-header -rf -p no:cacheprovider 2>&1 | tail -8

timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_r2.py -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -5
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; show $O/${TAG}_bench.json "default"

PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_attn_legacy.json 2> /dev/null; show $O/${TAG}_bench_attn_legacy.json "mma.sync attention"
PB200_TRACE=attention_tt:$O/${TAG}_trace_attention_tt.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
run() {  # name, kernel regex (demangled), skip, count, script...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 400 $NCU -k "regex:$rx" --launch-skip $skip -c $cnt -o /tmp/${TAG}_$name "$@" > $O/${TAG}_$name.log 2>&1
    echo "$name rc=$?"
    python tools/ncu_summary.py /tmp/${TAG}_$name.ncu-rep >> $O/${TAG}_ncu_summary.md 2>> $O/${TAG}_ncu_summary.err
    python tools/ncu_hot.py /tmp/${TAG}_$name.ncu-rep "::regex:$rx:1" 30 > $O/${TAG}_${name}_hot.txt 2>&1
}
P1="python tools/profile_step.py --sample-steps 1"
run attention_tt "attention_tt_kernel" 14 3 $P1

cut -c1-400 $O/${TAG}_ncu_summary.md | tail -14
head -20 $O/${TAG}_attention_tt_hot.txt | cut -c1-170
