#!/usr/bin/env bash
# round 2, GPU call G: state of HEAD after the container was re-created — tests, default bench line, attention timings,
# ncu launch list and full captures of the attention kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1
timeout 300 python tools/gpu_kernel_check.py --case attn_perf > gpurun_out/r2g_attn_perf.log 2>&1
PLM_BENCH_DETAIL=gpurun_out/r2g_bench_detail.txt timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4300 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_(fwd|bwd)' -c 4 -f -o gpurun_out/r2g_attn python tools/profile_one.py attn > gpurun_out/r2g_ncu_attn.log 2>&1
tail -2 gpurun_out/r2g_attn_perf.log | cut -c1-3000
tail -8 gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_bench.json | cut -c1-4000
tail -3 gpurun_out/r2g_bench.err
cat gpurun_out/r2g_bench_reference.json | cut -c1-1500
gzip -f gpurun_out/r2g_launches.csv
ls -la gpurun_out | tail -12
