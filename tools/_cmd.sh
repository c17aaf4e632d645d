python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py 2>&1 | grep "dp_check\|Error" | head -5
for mode in nccl multimem; do
echo "=== DALLE_B200_DP=$mode"
DALLE_B200_DP=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --extra "" 2> gpurun_out/dp_$mode.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
grep -i "error" gpurun_out/dp_$mode.err | head -3
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --extra "" --no-graph 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'])"
