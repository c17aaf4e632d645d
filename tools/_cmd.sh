python tools/gemm_gap.py 2>&1 | tail -8
echo "=== TMA store off"; DALLE_B200_GEMM_TMA_STORE=0 python tools/gemm_gap.py 2>&1 | head -4
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider --tb=line -k "gemm" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-250
