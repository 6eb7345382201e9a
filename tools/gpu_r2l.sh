#!/usr/bin/env bash
# round 2, GPU call L (8 GPUs): data-parallel tests, the default 420M line at 8 ranks (graph-captured DP micro-step), and
# BASELINE configs (4) document-masked 124M and (5) 1.5B / T4096 with signSGD at 8xB200.  Every step has its own timeout.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x > gpurun_out/r2l_pytest_ddp.log 2>&1
tail -5 gpurun_out/r2l_pytest_ddp.log | cut -c1-300
run8() { port=$1; shift; tmo=$1; shift; timeout $tmo python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@"; }
run8 29621 300 --steps 5 --warmup 3 > gpurun_out/r2l_bench_420m_n8.json 2> gpurun_out/r2l_bench_420m_n8.err
if ! grep -q '"value"' gpurun_out/r2l_bench_420m_n8.json; then
  echo "graphed DP run failed: eager DP micro-step"; tail -5 gpurun_out/r2l_bench_420m_n8.err | cut -c1-300
  export PLM_DP_GRAPH=0
  run8 29622 300 --steps 5 --warmup 3 > gpurun_out/r2l_bench_420m_n8_eager.json 2> gpurun_out/r2l_bench_420m_n8_eager.err
fi
run8 29623 240 --config 124m_doc --steps 8 --warmup 3 > gpurun_out/r2l_bench_124m_doc_n8.json 2> gpurun_out/r2l_bench_124m_doc_n8.err
run8 29624 420 --config 1p5b --optim signSGD --steps 3 --warmup 3 > gpurun_out/r2l_bench_1p5b_signsgd_n8.json 2> gpurun_out/r2l_bench_1p5b_signsgd_n8.err
for f in gpurun_out/r2l_bench_*.json; do echo $f; grep -o '"value": [0-9.]*' $f | head -1; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"dp_equiv.*' $f | cut -c1-200; done
tail -3 gpurun_out/r2l_bench_1p5b_signsgd_n8.err | cut -c1-300
