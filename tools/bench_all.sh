cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r1f_bench_420m.json 2> gpurun_out/r1f_bench_420m.err
python bench.py --config 124m_doc --no-cpu-baseline > gpurun_out/r1f_bench_124m_doc.json 2>> gpurun_out/r1f_bench_420m.err
python bench.py --config 1p5b --no-cpu-baseline > gpurun_out/r1f_bench_1p5b.json 2>> gpurun_out/r1f_bench_420m.err
python bench.py --config reduced --no-cpu-baseline > gpurun_out/r1f_bench_reduced.json 2>> gpurun_out/r1f_bench_420m.err
tail -c 300 gpurun_out/r1f_bench_420m.err
for f in gpurun_out/r1f_bench_*.json; do tail -1 $f | cut -c1-160; done
