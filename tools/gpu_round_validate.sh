#!/bin/bash
# Validation of the last build of round 2: whole -m gpu suite, benches, ncu of the rewritten codec edge kernels, sanitizer
TAG=${1:-r2y}; O=gpurun_out; mkdir -p $O
show() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    fam = {k: round(v["ms"], 2) for k, v in d["roofline"]["families"].items()}
    print(sys.argv[2] + ":", round(d["value"], 1), "img/s", round(d["ms_per_step"], 1), "ms/step", "e2e", round(d["e2e"]["value"], 1),
          "gemm-frac", round(d["roofline"]["frac"], 3), fam, "| with_decode", d.get("with_decode", {}).get("value"))
except Exception as e:
    print(sys.argv[2] + ": FAILED", e)
PY
}
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench.json 2> /dev/null; show $O/${TAG}_bench.json "sample"
timeout 400 python bench.py --workload vqgan --steps 3 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_vqgan256.json 2> /dev/null; show $O/${TAG}_vqgan256.json "vqgan bs=256"
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
timeout 400 $NCU -k "regex:vq_(in|out)_block|vq_dec_head" -c 3 -o /tmp/${TAG}_vq_inout python /tmp/vq_prof.py > $O/${TAG}_vq_inout.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_vq_inout.ncu-rep > $O/${TAG}_ncu_summary.md 2>&1; cut -c1-330 $O/${TAG}_ncu_summary.md | tail -4
bash tools/sanitize.sh ${TAG}
