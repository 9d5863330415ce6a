#!/bin/bash
# Round D: the evidence round of the current build on one B200 -- whole -m gpu suite, the driver's own bench invocation, attention
# kernel A/B (transposed tcgen05 / row-major tcgen05 / mma.sync), GRN fold A/B, ncu launch list + --set full summaries of the hot
# kernels (summarised on the box; .ncu-rep files stay in /tmp), in-kernel timeline of the attention kernel.
#   bash tools/gpu_round_d.sh <tag> [skip-tests]
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_smi.txt
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
if [ -z "$2" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1
  echo "pytest -m gpu rc=$?"; tail -12 $O/${TAG}_pytest.log
fi
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; show $O/${TAG}_bench_n1.json "default (driver invocation)"
tail -c 1500 $O/${TAG}_bench_n1.json; echo
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
PB200_ATTN_NO_TT=1 timeout 300 $B > $O/${TAG}_bench_attn_tc.json 2> /dev/null; show $O/${TAG}_bench_attn_tc.json "row-major tcgen05 attention"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_attn_legacy.json 2> /dev/null; show $O/${TAG}_bench_attn_legacy.json "mma.sync attention"
PB200_GRN_FOLD=1 timeout 300 $B > $O/${TAG}_bench_grnfold.json 2> /dev/null; show $O/${TAG}_bench_grnfold.json "GRN fold on"
PB200_TRACE=attention_tt:$O/${TAG}_trace_attention_tt.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
# ---- ncu: launch list of one 2-step sample() and full captures, summarised here
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file /tmp/${TAG}_launches.csv python tools/profile_step.py --sample-steps 2 > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py /tmp/${TAG}_launches.csv $O/${TAG}_launches_summary.md "Round 2 (${TAG}): sample() bs=64, 32x32 latents, 2 steps CFG" > /dev/null 2>&1
head -30 $O/${TAG}_launches_summary.md; gzip -c /tmp/${TAG}_launches.csv > $O/${TAG}_launches.csv.gz
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
run dwconv "dwconv" 4 3 $P1
run gemm_resid "cg2_kernel<.int.256, .int.3, .bool.0>" 4 3 $P1
run gemm_gelu "cg2_kernel<.int.256, .int.2, .bool.0>" 4 2 $P1
run grn "grn_(scale|apply)_kernel" 4 4 $P1
run sampler "fused_sampler" 0 1 $P1
cat > /tmp/vq_prof.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
vq = bench.build_vqgan(torch.device("cuda", 0)); vq.pack_weights()
img = torch.rand(16, 3, 256, 256, device="cuda")
idx = vq.encode(img)[2]; vq.decode_indices(idx); torch.cuda.synchronize()
torch.cuda.profiler.start()
idx = vq.encode(img)[2]; out = vq.decode_indices_u8(idx); torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
if [ -z "$3" ]; then
run vq_nearest "vq_nearest" 0 1 python /tmp/vq_prof.py
run vq_dw "vq_dw_residual" 2 2 python /tmp/vq_prof.py
run vq_conv "gemm_f16_kernel<.int.[0-9]+, .int.1, .int.[12]>" 0 2 python /tmp/vq_prof.py
run vq_inout "vq_(in|out)_block" 0 2 python /tmp/vq_prof.py
fi
cut -c1-400 $O/${TAG}_ncu_summary.md | tail -60
head -14 $O/${TAG}_attention_tt_hot.txt
rm -f $O/*.log.tmp; du -sh $O
