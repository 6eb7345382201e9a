#!/usr/bin/env bash
# round 2, GPU call N: RoPE (cos,sin) rows cached in shared memory through TMA; cross-entropy epilogue with independent
# reduction chains; ce_grad with two loads in flight
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1
tail -4 gpurun_out/r2n_pytest.log | cut -c1-300
timeout 400 python tools/gpu_kernel_check.py --only gemm --out gpurun_out/r2n_gemm_check.json --timeout 120 > gpurun_out/r2n_gemm_check.log 2>&1
timeout 200 python tools/gpu_kernel_check.py --case gemm_epi_perf > gpurun_out/r2n_gemm_epi_perf.log 2>&1
tail -1 gpurun_out/r2n_gemm_epi_perf.log | cut -c1-420
timeout 200 python tools/gpu_kernel_check.py --case lmhead_perf > gpurun_out/r2n_lmhead_perf.log 2>&1
tail -1 gpurun_out/r2n_lmhead_perf.log | cut -c1-700
PLM_BENCH_DETAIL=gpurun_out/r2n_bench_detail.txt timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
grep -o '"value": [0-9.]*' gpurun_out/r2n_bench.json | head -1; grep -o '"by_kernel_ms.*' gpurun_out/r2n_bench.json | cut -c1-420
grep "3072, 1024, 1, 1, 1)\|lmhead\|ce_grad  " gpurun_out/r2n_bench_detail.txt | cut -c1-150
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_gemm_check.json'))
bad=[(k,v) for k,v in d.items() if not k.endswith('__secs') and isinstance(v,dict) and (v.get('error') or v.get('nan') or v.get('rel_to_max',0)>2e-2)]
print('gemm check cases', sum(1 for k in d if not k.endswith('__secs')), 'bad', bad)
PY
