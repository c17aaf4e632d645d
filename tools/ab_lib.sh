# A/B of two library builds on one box: libdalle_b200.so (new) vs libdalle_b200_old.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or epilogue" 2>&1 | tail -2
for v in new old new old; do
  if [ $v = old ]; then export DALLE_B200_LIB=$PWD/dalle_pytorch_b200/libdalle_b200_old.so; else unset DALLE_B200_LIB; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2>/dev/null; echo $v; python tools/show_bench.py gpurun_out/bench_ab_$v.json 2>/dev/null | grep -v "x10256\|16384"
done
