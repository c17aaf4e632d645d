#!/usr/bin/env bash
# N=1 and N=2 bench back to back on one 2-GPU box (gpurun --gpus 2); results in gpurun_out/
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_on2.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err
python tools/show_bench.py gpurun_out/bench_n1_on2.json gpurun_out/bench_n2.json | grep -v "^    "
