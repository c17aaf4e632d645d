"""Times generate_images(use_cache=True) at the C2 weights on one GPU, host-indexed loop vs graph replay (bench.measure_decode).
    python tools/decode_probe.py [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench

if __name__ == '__main__':
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.cuda.set_device(0)

    class A:
        pass
    out = bench.measure_decode(A(), {'dev': torch.device('cuda:0')}, batch=batch, host_indexed='--graph-only' not in sys.argv)
    print(json.dumps(out))
