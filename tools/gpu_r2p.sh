#!/usr/bin/env bash
# round 2, GPU call P: the final code of the round — full GPU test suite, smoke(), the default bench line
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1
tail -4 gpurun_out/r2p_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -1 gpurun_out/r2p_smoke.log
PLM_BENCH_DETAIL=gpurun_out/r2p_bench_detail.txt timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
cut -c1-2200 gpurun_out/r2p_bench.json; tail -2 gpurun_out/r2p_bench.err
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2p_attn_perf.log 2>&1; tail -1 gpurun_out/r2p_attn_perf.log | cut -c1-1500
