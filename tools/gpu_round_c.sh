#!/bin/bash
# Round C: attention load-path variants (tests + bench A/B + in-kernel timelines), GRN-fold knobs, codec MLP timeline
TAG=${1:-r2e}
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
for mode in "" "PB200_ATTN_ALL_CP=1" "PB200_ATTN_TAILS_TMA=1"; do
  env $mode timeout 300 python -m pytest tests/test_gpu_attention.py -k tcgen05 -q --no-header -p no:cacheprovider 2>&1 | tail -2 | sed "s/^/[attention tests ${mode:-default}] /"
done
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; show $O/${TAG}_bench_default.json "default (TMA 64-wide + cp.async tails, non-blocking)"
PB200_ATTN_ALL_CP=1 timeout 300 $B > $O/${TAG}_bench_allcp.json 2> /dev/null; show $O/${TAG}_bench_allcp.json "attention all cp.async"
PB200_ATTN_TAILS_TMA=1 timeout 300 $B > $O/${TAG}_bench_tailstma.json 2> /dev/null; show $O/${TAG}_bench_tailstma.json "attention all TMA"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_legacy.json 2> /dev/null; show $O/${TAG}_bench_legacy.json "attention legacy"
PB200_GRN_FOLD=1 timeout 300 $B > $O/${TAG}_bench_grnfold.json 2> /dev/null; show $O/${TAG}_bench_grnfold.json "GRN fold on (256-wide)"
PB200_GRN_FOLD=1 PB200_GRN_FOLD_128=1 timeout 300 $B > $O/${TAG}_bench_grnfold128.json 2> /dev/null; show $O/${TAG}_bench_grnfold128.json "GRN fold on (256+128)"
PB200_TRACE=attention_tc:$O/${TAG}_trace_attn_default.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
PB200_ATTN_ALL_CP=1 PB200_TRACE=attention_tc:$O/${TAG}_trace_attn_allcp.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
PB200_ATTN_TAILS_TMA=1 PB200_TRACE=attention_tc:$O/${TAG}_trace_attn_tma.txt timeout 200 python tools/profile_step.py --sample-steps 1 > /dev/null 2>&1
cat > /tmp/vq_prof.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
vq = bench.build_vqgan(torch.device("cuda", 0)); vq.pack_weights()
img = torch.rand(16, 3, 256, 256, device="cuda")
for _ in range(6):
    idx = vq.encode(img)[2]; out = vq.decode_indices_u8(idx)
torch.cuda.synchronize()
PY
PB200_TRACE=vq_mlp:$O/${TAG}_trace_vqmlp.txt timeout 200 python /tmp/vq_prof.py > /dev/null 2>&1
ls -la $O | grep trace
VB="python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
timeout 300 $VB > $O/${TAG}_vqgan64.json 2> /dev/null; show $O/${TAG}_vqgan64.json "vqgan bs=64 fused MLP"
PB200_VQ_MLP_UNFUSED=1 timeout 300 $VB > $O/${TAG}_vqgan64_unfused.json 2> /dev/null; show $O/${TAG}_vqgan64_unfused.json "vqgan bs=64 unfused MLP"
du -sh $O
