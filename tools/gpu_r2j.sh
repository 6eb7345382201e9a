#!/usr/bin/env bash
# round 2, GPU call J (2 GPUs): data-parallel equivalence tests, the reference's train.py through the drop-in, and the
# 2-rank bench with the last micro-step captured (NCCL included) vs launched eagerly
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddp.py tests/test_gpu_train_dropin.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1
tail -25 gpurun_out/r2j_pytest.log
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 8 --warmup 3; }
PLM_DP_GRAPH=1 run2 29611 > gpurun_out/r2j_bench_n2_graph.json 2> gpurun_out/r2j_bench_n2_graph.err
PLM_DP_GRAPH=0 run2 29612 > gpurun_out/r2j_bench_n2_eager.json 2> gpurun_out/r2j_bench_n2_eager.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
for f in n2_graph n2_eager n1; do echo $f; grep -o '"value": [0-9.]*' gpurun_out/r2j_bench_$f.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_bench_$f.json | head -1; grep -o '"dp_equiv.*' gpurun_out/r2j_bench_$f.json | cut -c1-400; tail -4 gpurun_out/r2j_bench_$f.err | cut -c1-300; done
