#!/usr/bin/env bash
# round 2, GPU call C: three-stream forward with MUFU ticket lock + tile rotation + deep ring; new parity tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python tools/gpu_kernel_check.py --only attn_fwd3 --out gpurun_out/r2c_attn_fwd3_check.json --timeout 120 > gpurun_out/r2c_attn_fwd3_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2c_attn_perf.log 2>&1
PLM_ATTN_FWD_VARIANT=11 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_v11.log 2>&1
for v in 10 11; do
  PLM_ATTN_FWD_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c_bench_v$v.json 2> gpurun_out/r2c_bench_v$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd3_kernel' -c 4 -f -o gpurun_out/r2c_attn python tools/profile_one.py attn 10,11 > gpurun_out/r2c_ncu_attn.log 2>&1
tail -2 gpurun_out/r2c_attn_perf.log | cut -c1-2500
tail -15 gpurun_out/r2c_pytest_v11.log
for v in 10 11; do grep -o '"by_kernel_ms.*' gpurun_out/r2c_bench_v$v.json | cut -c1-300; grep -o '"value": [0-9.]*' gpurun_out/r2c_bench_v$v.json | head -1; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_attn_fwd3_check.json'))
bad=[(k,v) for k,v in d.items() if not k.endswith('__secs') and (v.get('error') or v.get('nan') or v.get('lse_nan') or v.get('rel_to_max',0)>2e-2 or v.get('lse_max_abs',0)>1e-3)]
print('fwd3 cases', sum(1 for k in d if not k.endswith('__secs')), 'bad', bad)
PY
