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

# fused epilogues (out-proj / FF2 + LayerScale + residual, FF1 + GEGLU): DALLE_B200_EPI=rows|cols selects the epilogue layout
M = 20480
for name, N, K in (('resid out-proj', 1024, 1024), ('resid FF2', 1024, 4096)):
    A = torch.randn(M, K, device='cuda').bfloat16(); W = torch.randn(N, K, device='cuda').bfloat16()
    bias = torch.randn(N, device='cuda'); resid = torch.randn(M, N, device='cuda'); scale = torch.rand(N, device='cuda')
    fn = lambda: ops.gemm_resid(A, W, bias, resid, scale, 1.0, keep_y=True)
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f'{name:14s} {M}x{N}x{K}: {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s')
A = torch.randn(M, 1024, device='cuda').bfloat16(); W1 = torch.randn(8192, 1024, device='cuda').bfloat16(); b1 = torch.randn(8192, device='cuda')
fn = lambda: ops.gemm_geglu(A, W1, b1, keep_u=True)
for _ in range(3):
    fn()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    fn()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
print(f'{"FF1 + GEGLU":14s} {M}x8192x1024: {ms * 1e3:7.1f} us  {2.0 * M * 8192 * 1024 / ms / 1e9:7.0f} TFLOP/s')
