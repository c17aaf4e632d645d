"""How busy is the GPU during one C2 step?  Kineto (torch.profiler, CUDA activities only) timeline: sum of kernel durations vs the
span from the first kernel start to the last kernel end, plus the distribution of the gaps between consecutive kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dalle_pytorch_b200 as D
from torch.profiler import profile, ProfilerActivity

D.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
vae = D.TokenVAE(image_size=256, num_layers=3, num_tokens=8192)
m = D.DALLE(dim=1024, vae=vae, num_text_tokens=10000, text_seq_len=256, depth=12, heads=16, dim_head=64).cuda().train()
text = torch.randint(1, 10000, (16, 256)).cuda()
image = torch.randint(0, 8192, (16, 1024)).cuda()


def step():
    for p in m.parameters():
        p.grad = None
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = m(text, image, return_loss=True)
    loss.backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA],
            key=lambda t: t[0])
span = ev[-1][1] - ev[0][0]
busy = sum(e[1] - e[0] for e in ev)
gaps = sorted(max(0, ev[i + 1][0] - ev[i][1]) for i in range(len(ev) - 1))
print(f'{len(ev)} kernels over 3 steps, span {span / 3e3:.2f} ms/step, busy {busy / 3e3:.2f} ms/step ({100 * busy / span:.1f} %), '
      f'idle {(span - busy) / 3e3:.2f} ms/step')
n = len(gaps)
print('gap percentiles (us): p50 %.1f p90 %.1f p99 %.1f max %.1f; sum of gaps > 20 us: %.2f ms/step' %
      (gaps[n // 2], gaps[int(n * .9)], gaps[int(n * .99)], gaps[-1], sum(g for g in gaps if g > 20) / 3e3))
big = sorted(((ev[i + 1][0] - ev[i][1], ev[i][2][:50], ev[i + 1][2][:50]) for i in range(len(ev) - 1)), reverse=True)[:8]
for g, a, b in big:
    print(f'  {g:8.1f} us between {a} -> {b}')
