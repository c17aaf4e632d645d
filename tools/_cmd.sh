run() { echo "=== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --extra "" 2>gpurun_out/dp8.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
run DALLE_B200_DP_OP=sum DALLE_B200_DP_OVERLAP=1
run DALLE_B200_DP_OP=sum DALLE_B200_DP_OVERLAP=0 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING
grep -i "nvls\|algo\|AllReduce" gpurun_out/dp8.err | head -8 | cut -c1-220
