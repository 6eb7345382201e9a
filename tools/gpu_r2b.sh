#!/usr/bin/env bash
# round 2, GPU call B: three-stream attention forward (variants 10..12), sleeping mbarrier waits, leaner backward
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python tools/gpu_kernel_check.py --only attn_fwd3 --out gpurun_out/r2b_attn_fwd3_check.json --timeout 120 > gpurun_out/r2b_attn_fwd3_check.log 2>&1
timeout 600 python tools/gpu_kernel_check.py --only attn_bwd --out gpurun_out/r2b_attn_bwd_check.json --timeout 120 > gpurun_out/r2b_attn_bwd_check.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2b_attn_perf.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
for v in 0 10 11 12; do
  PLM_ATTN_FWD_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2b_bench_v$v.json 2> gpurun_out/r2b_bench_v$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_(fwd|fwd3|bwd)_kernel' -c 8 -f -o gpurun_out/r2b_attn python tools/profile_one.py attn 0,11 > gpurun_out/r2b_ncu_attn.log 2>&1
tail -2 gpurun_out/r2b_attn_perf.log | cut -c1-2500
tail -4 gpurun_out/r2b_pytest.log
for v in 0 10 11 12; do tail -c 1500 gpurun_out/r2b_bench_v$v.json | grep -o '"by_kernel_ms.*' | cut -c1-300; grep -o '"value": [0-9.]*' gpurun_out/r2b_bench_v$v.json | head -1; done
python - <<'PY'
import json
for f in ('gpurun_out/r2b_attn_fwd3_check.json','gpurun_out/r2b_attn_bwd_check.json'):
    d=json.load(open(f))
    for k,v in d.items():
        if k.endswith('__secs'): continue
        print(k, {kk:(round(vv,5) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('rel_to_max','lse_max_abs','nan','lse_nan','error','dq_rel','dk_rel','dv_rel')})
PY
