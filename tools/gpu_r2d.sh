#!/usr/bin/env bash
# round 2, GPU call D: effective MUFU lock (release ordered behind the exp2 burst) in forward and backward
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python tools/gpu_kernel_check.py --only attn_fwd3 --out gpurun_out/r2d_attn_fwd3_check.json --timeout 120 > gpurun_out/r2d_attn_fwd3_check.log 2>&1
timeout 600 python tools/gpu_kernel_check.py --only attn_bwd --out gpurun_out/r2d_attn_bwd_check.json --timeout 120 > gpurun_out/r2d_attn_bwd_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2d_attn_perf.log 2>&1
PLM_ATTN_FWD_VARIANT=10 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest_v10.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_(fwd3|bwd)_kernel' -c 6 -f -o gpurun_out/r2d_attn python tools/profile_one.py attn 10,11 > gpurun_out/r2d_ncu_attn.log 2>&1
tail -2 gpurun_out/r2d_attn_perf.log | cut -c1-3000
tail -12 gpurun_out/r2d_pytest_v10.log
python - <<'PY'
import json
for f in ('gpurun_out/r2d_attn_fwd3_check.json','gpurun_out/r2d_attn_bwd_check.json'):
    d=json.load(open(f))
    bad=[(k,v) for k,v in d.items() if not k.endswith('__secs') and (v.get('error') or v.get('nan') or v.get('lse_nan') or v.get('rel_to_max',0)>2e-2 or v.get('lse_max_abs',0)>1e-3)]
    print(f, 'cases', sum(1 for k in d if not k.endswith('__secs')), 'bad', bad)
PY
