"""Attention-only probe at the C2 shape (b=16,h=16,n=1280): times each backend's fwd/bwd with CUDA events.
   python tools/attn_probe.py [--backend tc|mma|simt] [--pattern full|axial_row|axial_col] [--iters N]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dalle_pytorch_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--backend', default='tc')
ap.add_argument('--pattern', default='full')
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--batch', type=int, default=16)
ap.add_argument('--gather', type=int, default=1, help='axial patterns: 1 = gathered kernels (default), 0 = dense-tile predicate kernels')
args = ap.parse_args()
os.environ['DALLE_B200_ATTN'] = args.backend
code = {'full': 0, 'axial_row': 1, 'axial_col': 2, 'conv_like': 3}[args.pattern]
b, h, n, dh = args.batch, 16, 1280, 64
torch.manual_seed(0)
q = (torch.randn(b, h, n, dh, device='cuda') * dh ** -0.5).to(torch.bfloat16)
k = torch.randn(b, h, n, dh, device='cuda').to(torch.bfloat16)
v = torch.randn(b, h, n, dh, device='cuda').to(torch.bfloat16)
g = torch.randn(b, n, h * dh, device='cuda').to(torch.bfloat16)
spec = ops.AttnSpec(code, causal=True, text_len=257, fmap=32, kernel_size=5, dilation=1)
pairs = {'full': 819840, 'axial_row': 312928, 'axial_col': 312928}.get(args.pattern, 819840)


def timeit(fn):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / args.iters


lay = ops.gather_layout(spec, torch.bfloat16, n) if args.gather else None
if lay is not None:
    pad = lambda t: torch.cat([t, torch.zeros(b, h, lay.n_alloc - n, dh, device='cuda', dtype=t.dtype)], 2).contiguous()
    q, k, v = pad(q), pad(k), pad(v)
    gb = ops.attn_dout_buffer(lay, b * n, h * dh, q.device, q.dtype)
    gb.copy_(g.view(b * n, h * dh))
    g = gb.view(b, n, h * dh)
out, lse = ops.attn_fwd(spec, q, k, v, lay=lay)
t_f = timeit(lambda: ops.attn_fwd(spec, q, k, v, lay=lay))
t_b = timeit(lambda: ops.attn_bwd(spec, q, k, v, out, lse, g, None, None, 1.0, lay=lay))
fl = 4.0 * dh * pairs * b * h
print(f'[{args.backend} {args.pattern}{" gather" if lay is not None else ""}] fwd {t_f:.3f} ms ({fl / t_f / 1e9:.0f} TFLOP/s algorithmic)  bwd {t_b:.3f} ms ({2.5 * fl / t_b / 1e9:.0f} TFLOP/s)')
