#!/usr/bin/env bash
# round 2, GPU call K (1 GPU): the reference's train.py through the drop-in vs the reference itself; attention-backward
# item-walking CTAs at 2/4/8 per SM inside the step; N=1024 GEMM tile probe; 124M document-masked bench line
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_dropin.py -m gpu -q -x > gpurun_out/r2k_pytest_dropin.log 2>&1
tail -30 gpurun_out/r2k_pytest_dropin.log | cut -c1-400
for v in 0 2 4 8; do
  PLM_ATTN_BWD_VARIANT=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2k_bench_bwd$v.json 2> gpurun_out/r2k_bench_bwd$v.err
  echo "bwd variant $v: $(grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_bwd$v.json | head -1) $(grep -o '"attn_bwd": [0-9.]*' gpurun_out/r2k_bench_bwd$v.json | head -1)"
done
timeout 200 python tools/gpu_kernel_check.py --case gemm_n1024_probe > gpurun_out/r2k_gemm_n1024_probe.log 2>&1
tail -1 gpurun_out/r2k_gemm_n1024_probe.log | cut -c1-1500
timeout 300 python bench.py --config 124m_doc --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2k_bench_124m_doc.json 2> gpurun_out/r2k_bench_124m_doc.err
grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_124m_doc.json | head -1; grep -o '"mfu": {[^}]*}' gpurun_out/r2k_bench_124m_doc.json
