#!/usr/bin/env bash
# round 2, GPU call F: persistent backward with a dedicated dQ issuer; forward default = three streams + 1/4 FMA exp2
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python tools/gpu_kernel_check.py --only attn_bwd --out gpurun_out/r2f_attn_bwd_check.json --timeout 120 > gpurun_out/r2f_attn_bwd_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2f_attn_perf.log 2>&1
PLM_ATTN_BWD_VARIANT=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1
PLM_ATTN_BWD_VARIANT=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_persistent_kernel' -c 2 -f -o gpurun_out/r2f_attn_bwd env PLM_ATTN_BWD_VARIANT=1 python tools/profile_one.py attn 11 > gpurun_out/r2f_ncu_attn.log 2>&1
tail -2 gpurun_out/r2f_attn_perf.log | cut -c1-3000
tail -8 gpurun_out/r2f_pytest.log
grep -o '"by_kernel_ms.*' gpurun_out/r2f_bench.json | cut -c1-300; grep -o '"value": [0-9.]*' gpurun_out/r2f_bench.json | head -1
python - <<'PY'
import json
for f in ('gpurun_out/r2f_attn_bwd_check.json',):
    d=json.load(open(f))
    bad=[(k,v) for k,v in d.items() if not k.endswith('__secs') and (v.get('error') or v.get('nan') or v.get('rel_to_max',0)>2e-2)]
    print(f, 'cases', sum(1 for k in d if not k.endswith('__secs')), 'bad', bad)
PY
