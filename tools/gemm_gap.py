"""Times the C2-step GEMM shapes of the tcgen05 kernel in one process (CUDA events, 20 launches each).  With
DALLE_B200_GEMM_DBG=nostore|noepi (diagnosis switches in epilogue.cuh) it separates the mainloop from the STORE epilogue."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dalle_pytorch_b200 import ops

SHAPES = [('QKV fwd', 20480, 3072, 1024, False, False, False), ('dgrad FF2', 20480, 4096, 1024, False, True, False),
          ('dgrad to_out', 20480, 1024, 1024, False, True, False), ('dgrad QKV', 20480, 1024, 3072, False, True, False),
          ('dgrad FF1', 20480, 1024, 8192, False, True, False), ('wgrad FF1', 8192, 1024, 20480, True, True, True),
          ('wgrad to_out', 1024, 1024, 20480, True, True, True)]
torch.manual_seed(0)
for name, M, N, K, a_mn, b_mn, f32 in SHAPES:
    A = torch.randn((K, M) if a_mn else (M, K), device='cuda').bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device='cuda').bfloat16()
    fn = lambda: ops.gemm_store(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32 if f32 else None)
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f'{name:14s} {M}x{N}x{K}: {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s')
