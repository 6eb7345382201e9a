#!/usr/bin/env bash
# round 2, GPU call H: per-epilogue GEMM kernels, LM head fused with the cross-entropy forward, persistent attention
# backward as a default candidate
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case lmhead_perf > gpurun_out/r2h_lmhead_perf.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case gemm_epi_perf > gpurun_out/r2h_gemm_epi_perf.log 2>&1
PLM_BENCH_DETAIL=gpurun_out/r2h_bench_detail.txt timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
PLM_ATTN_BWD_VARIANT=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_full.py -m gpu -q -x -k "attention or flash or loss_curve or document" > gpurun_out/r2h_pytest_bwd1.log 2>&1
PLM_ATTN_BWD_VARIANT=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2h_bench_bwd1.json 2> gpurun_out/r2h_bench_bwd1.err
tail -5 gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest_bwd1.log
tail -1 gpurun_out/r2h_lmhead_perf.log | cut -c1-1500
tail -1 gpurun_out/r2h_gemm_epi_perf.log | cut -c1-2500
for f in gpurun_out/r2h_bench.json gpurun_out/r2h_bench_bwd1.json; do grep -o '"value": [0-9.]*' $f | head -1; grep -o '"by_kernel_ms.*' $f | cut -c1-700; done
tail -3 gpurun_out/r2h_bench.err
