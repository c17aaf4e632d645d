timeout 900 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider --tb=line -k "attention or axial_gather or inplace" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-200
for pat in full axial_col; do python tools/attn_probe.py --pattern $pat | grep "^\["; done
