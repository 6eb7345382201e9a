#!/usr/bin/env bash
# round 2, GPU call Q: fc2 input-gradient GEMM fused with the GLU backward
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "glu_backward or swiglu" > gpurun_out/r2q_pytest_glu.log 2>&1
tail -6 gpurun_out/r2q_pytest_glu.log | cut -c1-400
timeout 200 python tools/gpu_kernel_check.py --case glu_bwd_perf > gpurun_out/r2q_glu_bwd_perf.log 2>&1
tail -1 gpurun_out/r2q_glu_bwd_perf.log | cut -c1-900
for f in 1 0; do
  PLM_FUSE_GLU_BWD=$f timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2q_bench_fuse$f.json 2> gpurun_out/r2q_bench_fuse$f.err
  echo "fuse=$f $(grep -o '"value": [0-9.]*' gpurun_out/r2q_bench_fuse$f.json | head -1) $(grep -o '"by_kernel_ms.*' gpurun_out/r2q_bench_fuse$f.json | cut -c1-330)"; tail -2 gpurun_out/r2q_bench_fuse$f.err | cut -c1-300
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1
tail -4 gpurun_out/r2q_pytest.log | cut -c1-300
