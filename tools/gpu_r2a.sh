#!/usr/bin/env bash
# round 2, GPU call A: first contact of the new attention forward + reference-on-the-box + sanitizers
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python tools/gpu_kernel_check.py --only attn_fwd --out gpurun_out/r2a_attn_fwd_check.json --timeout 120 > gpurun_out/r2a_attn_fwd_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2a_attn_perf.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_420m.json 2> gpurun_out/r2a_bench_420m.err
SAN_TIMEOUT=200 timeout 1500 tools/sanitize.sh > gpurun_out/r2a_sanitize.log 2>&1
tail -3 gpurun_out/r2a_attn_perf.log | cut -c1-1500
tail -5 gpurun_out/r2a_pytest.log
tail -c 600 gpurun_out/r2a_bench_420m.json
cat gpurun_out/sanitizer_summary.txt
