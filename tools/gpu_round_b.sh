#!/bin/bash
# Round B on the GPU box: A/B knobs, secondary workloads with their reference-CUDA baselines, ncu launch list + full captures
# SUMMARISED ON THE BOX (the .ncu-rep files stay there: gpurun copies back at most 64 MiB).
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
for f in attention grnfold vqmlp; do
  if ls $O/*_${f}_fallback.txt > /dev/null 2>&1; then
    case $f in attention) export PB200_ATTN_LEGACY=1;; grnfold) export PB200_GRN_FOLD_BROKEN=1;; vqmlp) export PB200_VQ_MLP_UNFUSED=1;; esac
    echo "round B runs with the $f fallback"
  fi
done
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
show() {  # file label
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
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; show $O/${TAG}_bench_default.json "default"
PB200_ATTN_TAILS_TMA=1 timeout 300 $B > $O/${TAG}_bench_tails_tma.json 2> /dev/null; show $O/${TAG}_bench_tails_tma.json "attention tails by TMA"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_attn_legacy.json 2> /dev/null; show $O/${TAG}_bench_attn_legacy.json "attention legacy (mma.sync)"
PB200_NO_GRN_FOLD=1 timeout 300 $B > $O/${TAG}_bench_nogrnfold.json 2> /dev/null; show $O/${TAG}_bench_nogrnfold.json "no GRN fold"
VB="python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $VB > $O/${TAG}_vqgan64.json 2> $O/${TAG}_vqgan64.err; show $O/${TAG}_vqgan64.json "vqgan bs=64"
PB200_VQ_MLP_UNFUSED=1 timeout 300 $VB > $O/${TAG}_vqgan64_unfused.json 2> /dev/null; show $O/${TAG}_vqgan64_unfused.json "vqgan bs=64 unfused MLP"
timeout 600 python bench.py --workload vqgan --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_vqgan_full.json 2> $O/${TAG}_vqgan_full.err; show $O/${TAG}_vqgan_full.json "vqgan bs=256 (+reference-cuda)"
timeout 600 python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_sample64.json 2> $O/${TAG}_sample64.err; show $O/${TAG}_sample64.json "sample64 (+reference-cuda)"
# ---- ncu: launch list of one sample() and full captures, summarised here
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file /tmp/${TAG}_launches.csv python tools/profile_step.py --sample-steps 2 > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py /tmp/${TAG}_launches.csv $O/${TAG}_launches_summary.md "Round 2 (${TAG}): sample() bs=64, 32x32 latents, 2 steps CFG" > /dev/null 2>&1; head -30 $O/${TAG}_launches_summary.md; gzip -c /tmp/${TAG}_launches.csv > $O/${TAG}_launches.csv.gz
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
run() {  # name, kernel regex (demangled), skip, count, script...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 500 $NCU -k "regex:$rx" --launch-skip $skip -c $cnt -o /tmp/${TAG}_$name "$@" > $O/${TAG}_$name.log 2>&1
    echo "$name rc=$?"
    python tools/ncu_summary.py /tmp/${TAG}_$name.ncu-rep >> $O/${TAG}_ncu_summary.md 2>> $O/${TAG}_ncu_summary.err
    python tools/ncu_hot.py /tmp/${TAG}_$name.ncu-rep "::regex:$rx:1" 30 > $O/${TAG}_${name}_hot.txt 2>&1
}
run attention_tc "attention_tc_kernel" 14 3 python tools/profile_step.py --sample-steps 1
run grn_scale "grn_scale_kernel" 4 2 python tools/profile_step.py --sample-steps 1
run gemm_resid_ascale "cg2_kernel<.int.256, .int.3, .bool.1>" 4 3 python tools/profile_step.py --sample-steps 1
run gemm_gelu "cg2_kernel<.int.256, .int.2, .bool.0>" 4 2 python tools/profile_step.py --sample-steps 1
run dwconv "dwconv" 4 3 python tools/profile_step.py --sample-steps 1
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
run vq_mlp "vq_mlp_fused" 2 2 python /tmp/vq_prof.py
run vq_nearest "vq_nearest" 0 1 python /tmp/vq_prof.py
run vq_dw "vq_dw_residual" 2 2 python /tmp/vq_prof.py
run vq_conv "gemm_f16_kernel<.int.[0-9]+, .int.1, .int.[12]>" 0 2 python /tmp/vq_prof.py
run vq_inout "vq_(in|out)_block" 0 2 python /tmp/vq_prof.py
tail -40 $O/${TAG}_ncu_summary.md
head -12 $O/${TAG}_attention_tc_hot.txt
bash tools/sanitize.sh ${TAG}
du -sh $O
