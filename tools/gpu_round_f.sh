#!/bin/bash
# Round F: codec ResBlock front fused into one patch kernel (+ row statistics pre-pass): tests, codec bench A/B, ncu of the new kernels;
# attention producer back-off A/B on the headline bench.   bash tools/gpu_round_f.sh <tag>
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
show() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d["roofline"]["families"].items()}
    print(sys.argv[2] + ":", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms/step", "e2e", round(d["e2e"]["value"], 1),
          "gemm-frac", round(d["roofline"]["frac"], 3), fam, d.get("with_decode", {}).get("value"))
except Exception as e:
    print(sys.argv[2] + ": FAILED", e)
PY
}
timeout 600 python -m pytest tests/test_gpu_vqgan.py tests/test_gpu_attention.py tests/test_gpu_parity_r2.py -q --no-header -rf -p no:cacheprovider -x 2>&1 | tail -8
PB200_VQ_FRONT_REGS=1 timeout 300 python -m pytest tests/test_gpu_vqgan.py -q --no-header -rf -p no:cacheprovider -x -k "resblock or roundtrip or decode" 2>&1 | tail -3
V="python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $V > $O/${TAG}_vqgan64.json 2> $O/${TAG}_vqgan64.err; show $O/${TAG}_vqgan64.json "vqgan bs=64 fused front (128 regs)"
PB200_VQ_FRONT_REGS=1 timeout 300 $V > $O/${TAG}_vqgan64_regs.json 2> /dev/null; show $O/${TAG}_vqgan64_regs.json "vqgan bs=64 fused front (168 regs)"
PB200_VQ_FRONT_UNFUSED=1 timeout 300 $V > $O/${TAG}_vqgan64_unfused.json 2> /dev/null; show $O/${TAG}_vqgan64_unfused.json "vqgan bs=64 three-launch front"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; show $O/${TAG}_bench.json "sample default"
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
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
run() {  # name, kernel regex (demangled), skip, count, script...
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 400 $NCU -k "regex:$rx" --launch-skip $skip -c $cnt -o /tmp/${TAG}_$name "$@" > $O/${TAG}_$name.log 2>&1
    echo "$name rc=$?"
    python tools/ncu_summary.py /tmp/${TAG}_$name.ncu-rep >> $O/${TAG}_ncu_summary.md 2>> $O/${TAG}_ncu_summary.err
    python tools/ncu_hot.py /tmp/${TAG}_$name.ncu-rep "::regex:$rx:1" 30 > $O/${TAG}_${name}_hot.txt 2>&1
}
run vq_front "vq_front_patch_kernel|row_stats_kernel" 4 4 python /tmp/vq_prof.py
run attention_tt "attention_tt_kernel" 14 2 python tools/profile_step.py --sample-steps 1
cut -c1-400 $O/${TAG}_ncu_summary.md | tail -16
head -16 $O/${TAG}_vq_front_hot.txt | cut -c1-170
