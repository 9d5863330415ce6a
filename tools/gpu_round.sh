#!/bin/bash
# One GPU round on the box: the attention kernel's own tests first (a trap there would poison the CUDA context of everything
# after it in the same process), then the whole -m gpu suite, then the bench.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 600 python -m pytest tests/test_gpu_attention.py -q --no-header -rf -p no:cacheprovider > gpurun_out/${TAG}_pytest_attention.log 2>&1
if [ $? -ne 0 ]; then
  echo "attention tests FAILED: rest of the round runs with PB200_ATTN_LEGACY=1" | tee gpurun_out/${TAG}_attention_fallback.txt
  export PB200_ATTN_LEGACY=1
fi
tail -15 gpurun_out/${TAG}_pytest_attention.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "a_scale" -q --no-header -rf -p no:cacheprovider > gpurun_out/${TAG}_pytest_ascale.log 2>&1
if [ $? -ne 0 ]; then
  echo "a_scale GEMM tests FAILED: rest of the round runs with PB200_NO_GRN_FOLD=1" | tee gpurun_out/${TAG}_grnfold_fallback.txt
  export PB200_GRN_FOLD_BROKEN=1
fi
tail -8 gpurun_out/${TAG}_pytest_ascale.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --deselect tests/test_gpu_attention.py > gpurun_out/${TAG}_pytest.log 2>&1
tail -40 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
