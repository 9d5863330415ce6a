#!/bin/bash
# Final evidence round of round 2 (one B200): whole -m gpu suite, the driver's bench invocation, the secondary workloads with their
# reference-CUDA arms, the reference arms, attention A/B, ncu launch list + --set full summaries of every hot kernel of both models
# (summarised on the box), attention timeline, compute-sanitizer.      bash tools/gpu_round_final.sh <tag>
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_smi.txt
show() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d["roofline"]["families"].items()}
    tc = d.get("torch_cuda_baseline", {})
    print(sys.argv[2] + ":", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms/step", "e2e", round(d["e2e"]["value"], 1),
          "gemm-frac", round(d["roofline"]["frac"], 3), fam, "| with_decode", d.get("with_decode", {}).get("value"),
          "| ref-cuda", tc.get("value"), tc.get("ours_over_reference_cuda"))
except Exception as e:
    print(sys.argv[2] + ": FAILED", e)
PY
}
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1
echo "pytest -m gpu rc=$?"; tail -6 $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; show $O/${TAG}_bench_n1.json "sample (driver invocation)"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/${TAG}_bench_reference_cpu.json 2> /dev/null; tail -c 600 $O/${TAG}_bench_reference_cpu.json; echo
timeout 900 python bench.py --workload vqgan --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_vqgan256.json 2> $O/${TAG}_vqgan256.err; show $O/${TAG}_vqgan256.json "vqgan bs=256"
timeout 900 python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_sample64.json 2> $O/${TAG}_sample64.err; show $O/${TAG}_sample64.json "sample64 bs=16"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
PB200_ATTN_NO_TT=1 timeout 300 $B > $O/${TAG}_bench_attn_tc.json 2> /dev/null; show $O/${TAG}_bench_attn_tc.json "row-major tcgen05 attention"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_attn_legacy.json 2> /dev/null; show $O/${TAG}_bench_attn_legacy.json "mma.sync attention"
PB200_ATTN_LEGACY=1 timeout 400 python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_sample64_legacy.json 2> /dev/null; show $O/${TAG}_sample64_legacy.json "sample64 mma.sync attention"
PB200_TRACE=attention_tt:$O/${TAG}_trace_attention_tt.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
# ---- ncu
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file /tmp/${TAG}_launches.csv python tools/profile_step.py --sample-steps 2 > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py /tmp/${TAG}_launches.csv $O/${TAG}_launches_summary.md "Round 2 final (${TAG}): sample() bs=64, 32x32 latents, 2 steps CFG" > /dev/null 2>&1
head -16 $O/${TAG}_launches_summary.md; gzip -c /tmp/${TAG}_launches.csv > $O/${TAG}_launches.csv.gz
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
run gemm_resid "cg2_kernel<.int.256, .int.[36], .bool.0>" 4 4 $P1
run gemm_gelu "cg2_kernel<.int.256, .int.2, .bool.0>" 4 2 $P1
run gemm_qkv "cg2_kernel<.int.256, .int.7, .bool.0>" 2 2 $P1
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
run vq_front "vq_front_patch_kernel|row_stats_kernel" 4 4 python /tmp/vq_prof.py
run vq_mlp_gemms "cg2_kernel<.int.[0-9]+, .int.[23], .bool.0>" 6 4 python /tmp/vq_prof.py
run vq_nearest "vq_nearest" 0 1 python /tmp/vq_prof.py
run vq_conv "gemm_f16_kernel<.int.[0-9]+, .int.1, .int.[12]>" 0 2 python /tmp/vq_prof.py
run vq_inout "vq_(in|out)_block|vq_dec_head" 0 3 python /tmp/vq_prof.py
cut -c1-400 $O/${TAG}_ncu_summary.md | grep -v "^|---" | grep "^|" | grep -v "kernel | dur" | cut -c1-200
bash tools/sanitize.sh ${TAG}
du -sh $O
