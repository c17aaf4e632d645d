"""Host-side cost of one C2 step: wall time to ENQUEUE a step (GPU queue never full: a sync every step) vs GPU time, and the
top Python functions by cumulative time (cProfile on the main thread; the backward runs on the autograd thread and shows up as the
time inside loss.backward())."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dalle_pytorch_b200 as D

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
    t1 = time.perf_counter()
    loss.backward()
    return t1


for _ in range(4):
    step()
torch.cuda.synchronize()
fw, bw = [], []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t1 = step()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    fw.append(t1 - t0); bw.append(t2 - t1)
    last = t3 - t0
print(f'enqueue forward {1e3 * min(fw):.2f} ms, enqueue backward {1e3 * min(bw):.2f} ms, step incl. GPU {1e3 * last:.2f} ms '
      f'(host is ahead of the GPU when enqueue < GPU time)')
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(14)
