#!/usr/bin/env bash
# round 2, GPU call M: evidence for the final code of the round — smoke(), the default bench line and the reference arm,
# the ncu launch list of the same command, and full captures of the GEMM kinds (incl. the LM head + cross-entropy one)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
PLM_BENCH_DETAIL=gpurun_out/r2m_bench_detail.txt timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4300 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2m_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|ce_grad|ce_finalize' -c 7 -f -o gpurun_out/r2m_lmhead python tools/profile_one.py lmhead > gpurun_out/r2m_ncu_lmhead.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -c 8 -f -o gpurun_out/r2m_gemm python tools/profile_one.py gemm > gpurun_out/r2m_ncu_gemm.log 2>&1
cut -c1-1500 gpurun_out/r2m_bench.json; tail -2 gpurun_out/r2m_bench.err
cut -c1-300 gpurun_out/r2m_bench_reference.json
gzip -f gpurun_out/r2m_launches.csv
ls -la gpurun_out | grep r2m
