#!/bin/bash
# Round B on the GPU box: A/B knobs, secondary workloads, reference-CUDA baselines for them, ncu launch list + full captures.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
if ls $O/*_attention_fallback.txt > /dev/null 2>&1; then export PB200_ATTN_LEGACY=1; echo "round B runs with PB200_ATTN_LEGACY=1"; fi
if ls $O/*_dwslab_fallback.txt > /dev/null 2>&1; then export PB200_DWCONV_NOSLAB=1; echo "round B runs with PB200_DWCONV_NOSLAB=1"; fi
if ls $O/*_vqmlp_fallback.txt > /dev/null 2>&1; then export PB200_VQ_MLP_UNFUSED=1; echo "round B runs with PB200_VQ_MLP_UNFUSED=1"; fi
if ls $O/*_grnfold_fallback.txt > /dev/null 2>&1; then export PB200_NO_GRN_FOLD=1; echo "round B runs with PB200_NO_GRN_FOLD=1"; fi
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cuda-baseline"
for sb in 0 32 16; do
  PB200_SUBBATCH=$sb timeout 300 $B > $O/${TAG}_bench_subbatch${sb}.json 2> $O/${TAG}_bench_subbatch${sb}.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_subbatch${sb}.json").read().strip().splitlines()[-1])
    print("subbatch $sb:", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms", {k: v["ms"] for k, v in d["roofline"]["families"].items()})
except Exception as e:
    print("subbatch $sb: FAILED", e)
PY
done
for cfg in "32 2" "16 2" "16 4"; do
  set -- $cfg
  PB200_SUBBATCH=$1 PB200_STREAMS=$2 timeout 300 $B > $O/${TAG}_bench_sub$1_str$2.json 2> $O/${TAG}_bench_sub$1_str$2.err
  python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_bench_sub$1_str$2.json').read().strip().splitlines()[-1]); print('subbatch $1 streams $2:', round(d['value'],1), 'img/s', round(d['ms_per_step'],1), 'ms')
except Exception as e: print('subbatch $1 streams $2 FAILED', e)"
done
PB200_DWCONV_NOSLAB=1 timeout 300 $B > $O/${TAG}_bench_noslab.json 2> $O/${TAG}_bench_noslab.err
python -c "
import json; d=json.loads(open('$O/${TAG}_bench_noslab.json').read().strip().splitlines()[-1]); print('no dwconv slab:', round(d['value'],1), {k: v['ms'] for k, v in d['roofline']['families'].items() if 'dw' in k})"
PB200_NO_GRN_FOLD=1 timeout 300 $B > $O/${TAG}_bench_nogrnfold.json 2> $O/${TAG}_bench_nogrnfold.err
python -c "
import json; d=json.loads(open('$O/${TAG}_bench_nogrnfold.json').read().strip().splitlines()[-1]); print('no GRN fold:', round(d['value'],1), {k: v['ms'] for k, v in d['roofline']['families'].items()})"
PB200_ATTN_LEGACY=1 timeout 300 $B > $O/${TAG}_bench_attn_legacy.json 2> $O/${TAG}_bench_attn_legacy.err
python -c "
import json; d=json.loads(open('$O/${TAG}_bench_attn_legacy.json').read().strip().splitlines()[-1]); print('attn legacy:', round(d['value'],1), {k: v['ms'] for k, v in d['roofline']['families'].items() if 'att' in k})"
for vs in 0 2 4 8 16; do
  PB200_VQ_SUBBATCH=$vs timeout 300 python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_vqgan_sub${vs}.json 2> $O/${TAG}_vqgan_sub${vs}.err
  python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_vqgan_sub${vs}.json').read().strip().splitlines()[-1]); print('vqgan sub $vs:', round(d['value'],1), 'img/s', {k: round(v['ms'],2) for k, v in d['roofline']['families'].items()})
except Exception as e: print('vqgan sub $vs FAILED', e)"
done
PB200_VQ_MLP_UNFUSED=1 timeout 300 python bench.py --workload vqgan --batch 64 --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_vqgan_unfused.json 2> $O/${TAG}_vqgan_unfused.err
python -c "
import json
try:
    d=json.loads(open('$O/${TAG}_vqgan_unfused.json').read().strip().splitlines()[-1]); print('vqgan unfused MLP:', round(d['value'],1), 'img/s', {k: round(v['ms'],2) for k, v in d['roofline']['families'].items()})
except Exception as e: print('vqgan unfused FAILED', e)"
timeout 600 python bench.py --workload vqgan --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_vqgan_full.json 2> $O/${TAG}_vqgan_full.err
tail -c 1200 $O/${TAG}_vqgan_full.json
timeout 600 python bench.py --workload sample64 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_sample64.json 2> $O/${TAG}_sample64.err
tail -c 1200 $O/${TAG}_sample64.json
# ---- ncu: launch list of one sample() and full captures of the kernels this round changed
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${TAG}_launches.csv python tools/profile_step.py --sample-steps 2 > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none -f --kernel-name-base demangled"
run() {  # name, kernel regex (demangled), skip, count
    timeout 400 $NCU -k "regex:$2" --launch-skip $3 -c $4 -o $O/${TAG}_$1 python tools/profile_step.py --sample-steps 1 > $O/${TAG}_$1.log 2>&1
    echo "$1 rc=$?"
}
run attention_tc "attention_tc_kernel" 14 3
run grn_apply "grn_apply_kernel" 4 3
run gemm_resid "cg2_kernel<.int.256, .int.3>" 4 3
gzip -f $O/${TAG}_launches.csv
cat > /tmp/vq_prof.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
vq = bench.build_vqgan(torch.device("cuda", 0)); vq.pack_weights()
img = torch.rand(8, 3, 256, 256, device="cuda")
idx = vq.encode(img)[2]; vq.decode_indices(idx); torch.cuda.synchronize()
torch.cuda.profiler.start()
idx = vq.encode(img)[2]; out = vq.decode_indices_u8(idx); torch.cuda.synchronize()
torch.cuda.profiler.stop()
PY
timeout 500 $NCU -c 120 -o $O/${TAG}_vqgan python /tmp/vq_prof.py > $O/${TAG}_vqgan.log 2>&1
echo "vqgan ncu rc=$?"
ls -la $O | grep ${TAG} | tail -40
