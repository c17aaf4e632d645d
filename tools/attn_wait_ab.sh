for m in 0 16; do echo "mode=$m"; DALLE_B200_ATTN_WAIT=$m timeout 120 python tools/attn_probe.py --backend tc --pattern full 2>&1 | grep "^\["; done
