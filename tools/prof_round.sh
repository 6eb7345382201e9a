set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4300 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1f_bench_under_ncu.log 2>&1
for k in gemm attn norm; do
  case $k in gemm) rx='gemm_kernel';; attn) rx='attn_(fwd|bwd)_kernel';; norm) rx='rmsnorm';; esac
  ncu --set full --clock-control none --import-source on -k regex:$rx -c 8 -f -o gpurun_out/r1f_$k python tools/profile_one.py $k > gpurun_out/r1f_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -8
