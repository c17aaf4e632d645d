"""torchrun --nproc-per-node N tools/dp_check.py : the data-parallel gradient mean through the multimem path (weight-gradient
GEMMs and dalle_b200_mc_add reducing into every GPU's replica through the NVLink multicast address) against the NCCL all-reduce
path and against a single-process evaluation of all ranks' micro-batches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import torch.distributed as dist
import dalle_pytorch_b200 as D
from dalle_pytorch_b200.distributed import GradAllReducer
from dalle_oracle import OracleConfig, make_state_dict, make_inputs

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
cfg = OracleConfig(dim=256, depth=2, heads=4, text_seq_len=64, fmap=16, num_text_tokens=1000, num_image_tokens=512, attn_types=('full', 'axial_col'))
sd = make_state_dict(cfg, seed=2)
D.set_compute_dtype(torch.bfloat16)


def build():
    vae = D.TokenVAE(image_size=8 * cfg.fmap, num_layers=3, num_tokens=cfg.num_image_tokens)
    m = D.DALLE(dim=cfg.dim, vae=vae, num_text_tokens=cfg.num_text_tokens, text_seq_len=cfg.text_seq_len, depth=cfg.depth, heads=cfg.heads,
                attn_types=cfg.attn_types)
    m.load_state_dict(sd)
    return m.cuda().train()


def step(m, r):
    text, image = make_inputs(cfg, 4, seed=100 + r)
    loss = m(text.cuda(), image.cuda(), return_loss=True)
    loss.backward()
    return loss


res = {}
for mode in ('nccl', 'multimem'):
    m = build()
    red = GradAllReducer(list(m.parameters()), mode=mode)
    assert (red.mc is not None) == (mode == 'multimem'), (mode, red.mc)
    for it in range(3):                       # several steps: the clear / barrier protocol must hold from step to step
        red.zero_grad()
        step(m, rank)
        red.finish()
    torch.cuda.synchronize()
    res[mode] = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    red.remove()
# single-process mean over the ranks' micro-batches
m = build()
tot = None
for r in range(world):
    for p in m.parameters():
        p.grad = None
    step(m, r)
    g = {k: p.grad.detach().clone() / world for k, p in m.named_parameters()}
    tot = g if tot is None else {k: tot[k] + g[k] for k in g}
worst = 0.0
for k in tot:
    scale = float(tot[k].abs().max()) + 1e-12
    e1 = float((res['multimem'][k] - tot[k]).abs().max()) / scale
    e2 = float((res['nccl'][k] - tot[k]).abs().max()) / scale
    e3 = float((res['multimem'][k] - res['nccl'][k]).abs().max()) / scale
    worst = max(worst, e1, e3)
    assert e1 < 2e-2 and e3 < 2e-2, (k, e1, e2, e3)
# every rank must hold the same reduced gradient
chk = torch.stack([res['multimem'][k].double().sum() for k in sorted(tot)])
ref = chk.clone()
dist.broadcast(ref, src=0)
assert torch.allclose(chk, ref, rtol=1e-6, atol=1e-9), 'ranks disagree on the multimem-reduced gradients'
if rank == 0:
    print(f'dp_check ok: world {world}, {len(tot)} tensors, worst relative deviation {worst:.2e} (bf16 step, atomics)')
dist.barrier()
dist.destroy_process_group()
