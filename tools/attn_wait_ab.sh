timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for pat in full axial_row axial_col; do timeout 120 python tools/attn_probe.py --backend tc --pattern $pat 2>&1 | grep "^\["; done
DALLE_B200_ATTN_WAIT=4 timeout 100 python tools/attn_timeline.py 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_auto.json 2>/dev/null; python tools/show_bench.py gpurun_out/bench_c2_auto.json 2>/dev/null | head -1
