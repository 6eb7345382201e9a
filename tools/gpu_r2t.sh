#!/usr/bin/env bash
# round 2, GPU call T: the final code of the round — GPU tests, smoke(), default bench line, ncu launch list of the
# same command, full capture of the GLU-backward GEMM kind next to the unfused pair
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1
tail -3 gpurun_out/r2t_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; tail -1 gpurun_out/r2t_smoke.log
PLM_BENCH_DETAIL=gpurun_out/r2t_bench_detail.txt timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
cut -c1-1900 gpurun_out/r2t_bench.json; tail -2 gpurun_out/r2t_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4300 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2t_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|swiglu_bwd' -c 6 -f -o gpurun_out/r2t_glu_bwd python tools/profile_one.py glu_bwd > gpurun_out/r2t_ncu_glu_bwd.log 2>&1
gzip -f gpurun_out/r2t_launches.csv
ls -la gpurun_out | grep r2t
