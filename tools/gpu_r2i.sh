#!/usr/bin/env bash
# round 2, GPU call I: fp32-residual GEMM epilogue with the residual tile streamed in through TMA
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
timeout 600 python tools/gpu_kernel_check.py --only gemm --out gpurun_out/r2i_gemm_check.json --timeout 120 > gpurun_out/r2i_gemm_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case gemm_epi_perf > gpurun_out/r2i_gemm_epi_perf.log 2>&1
PLM_BENCH_DETAIL=gpurun_out/r2i_bench_detail.txt timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -5 gpurun_out/r2i_pytest.log
tail -1 gpurun_out/r2i_gemm_epi_perf.log | cut -c1-600
grep -o '"value": [0-9.]*' gpurun_out/r2i_bench.json | head -1; grep -o '"by_kernel_ms.*' gpurun_out/r2i_bench.json | cut -c1-400
grep "^gemm .*1, 1, 3)" gpurun_out/r2i_bench_detail.txt
tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_gemm_check.json'))
bad=[(k,v) for k,v in d.items() if not k.endswith('__secs') and isinstance(v,dict) and (v.get('error') or v.get('nan') or v.get('rel_to_max',0)>2e-2)]
print('gemm check cases', sum(1 for k in d if not k.endswith('__secs')), 'bad', bad)
PY
