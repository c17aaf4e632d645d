for m in cols rows; do echo "=== DALLE_B200_EPI=$m"; DALLE_B200_EPI=$m python tools/gemm_gap.py 2>&1 | tail -3; done
DALLE_B200_EPI=rows timeout 600 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider --tb=line -k "gemm" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-250
