timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -2
for pat in full axial_row axial_col; do timeout 120 python tools/attn_probe.py --backend tc --pattern $pat 2>&1 | grep "^\["; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_auto.json 2>/dev/null; python tools/show_bench.py gpurun_out/bench_c2_auto.json 2>/dev/null | head -1
