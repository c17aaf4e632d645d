"""Debug: clock64 timeline of one dK/dV CTA (block (1,0)) at the C2 attention shape.  DALLE_B200_ATTN_WAIT=4 enables the probe."""
import ctypes, os, sys
os.environ['DALLE_B200_ATTN_WAIT'] = os.environ.get('DALLE_B200_ATTN_WAIT', '4')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dalle_pytorch_b200 import ops, _lib
b, h, n, dh = 16, 16, 1280, 64
torch.manual_seed(0)
q = (torch.randn(b, h, n, dh, device='cuda') * dh ** -0.5).to(torch.bfloat16)
k = torch.randn(b, h, n, dh, device='cuda').to(torch.bfloat16)
v = torch.randn(b, h, n, dh, device='cuda').to(torch.bfloat16)
g = torch.randn(b, n, h * dh, device='cuda').to(torch.bfloat16)
spec = ops.AttnSpec(0, causal=True, text_len=257, fmap=32, kernel_size=5, dilation=1)
o, lse = ops.attn_fwd(spec, q, k, v, None)
for _ in range(3):
    ops.attn_bwd(spec, q, k, v, o, lse, g, None, None, 1.0, None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 96)()
lib = _lib.lib() if callable(getattr(_lib, 'lib', None)) else _lib.LIB
lib.dalle_b200_debug_attn_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.dalle_b200_debug_attn_timeline(buf, 96) == 0
names = ['mma:q_full', 'mma:S,dP issued', 'mma:ps_ready', 'mma:dV,dK issued', 'sm0:wait st_full', 'sm0:st_full', 'sm0:computed', 'sm0:arrived',
         'sm4:wait st_full', 'sm4:st_full', 'sm4:computed', 'sm4:arrived']
t0 = min(x for x in buf if x > 0)
for it in range(6):
    print(f'iteration {it}: ' + '  '.join(f'{names[s]}={buf[s * 6 + it] - t0}' for s in range(12)))
