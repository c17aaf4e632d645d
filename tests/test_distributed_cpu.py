"""World-size-2 `gloo` test of the data-parallel plumbing on CPU (the NCCL path uses the same code with
backend='nccl'): parameters are broadcast from rank 0, gradients come out as the mean over ranks through the flat
bucketed all-reduce, gradients written outside autograd (reversible executor style) are reduced by finish(), and
`average_all` averages a scalar (reference deepspeed_backend.py:165-171)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _SideGrads(torch.autograd.Function):
    """Identity whose backward hands two gradients over the way the fused sub-layers / the reversible executor do: one written
    straight into the flat-buffer slot (functional._slot/_commit), one assigned to .grad without telling the reducer."""

    @staticmethod
    def forward(ctx, x, direct, extra, rank):
        ctx.direct, ctx.extra, ctx.rank = direct, extra, rank
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        red = ctx.direct._b200_reducer
        ctx.direct._b200_uses = 1                                      # what functional._note_use records in forward
        slot = red.direct_slot(ctx.direct)
        assert slot is not None and slot.data_ptr() == red.views[ctx.direct].data_ptr()
        slot.copy_(torch.full((3, 2), 10.0 * (ctx.rank + 1)))          # stands in for the weight-gradient GEMM
        red.direct_done(ctx.direct)
        assert red.direct_slot(ctx.direct) is None                     # a second use in the same step must go through autograd
        assert ctx.extra.grad is None                                  # gradients start each step unset (no memset of the flat buffer)
        ctx.extra.grad = torch.full((5,), float(ctx.rank + 1))         # finish() adopts it into the flat buffer
        return g, None, None, None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from dalle_pytorch_b200.distributed import NCCLBackend
    be = NCCLBackend(comm_backend='gloo', bucket_bytes=256)      # tiny buckets -> several of them
    be.initialize()
    assert be.get_world_size() == world and be.get_rank() == rank
    be.check_batch_size(4)
    torch.manual_seed(100 + rank)                                  # different init per rank: broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    extra = torch.nn.Parameter(torch.zeros(5))                     # receives its gradient outside autograd
    model.register_parameter('extra', extra)
    direct = torch.nn.Parameter(torch.zeros(3, 2))                 # gradient written straight into the flat buffer (fused wgrad path)
    model.register_parameter('direct', direct)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model, opt, _, _ = be.distribute(model=model, optimizer=opt)
    w0 = [p.detach().clone() for p in model.parameters()]
    red = model.grad_reducer
    torch.manual_seed(7 + rank)
    x = torch.randn(3, 8)
    loss = _SideGrads.apply(model(x), direct, extra, rank).square().mean()
    loss.backward()
    # the reduction completed inside backward() (end-of-backward callback): .grad is already the mean over ranks
    assert red._finished and not red._works
    grads = [p.grad.detach().clone() for p in model.parameters()]
    red.finish()                                                   # explicit call: idempotent (no second 1/world)
    assert all(torch.equal(g, p.grad) for g, p in zip(grads, model.parameters()))
    avg_loss = be.average_all(loss.detach())
    assert all(p.grad.data_ptr() == red.views[p].data_ptr() for p in model.parameters())
    # the reference loop (train_dalle.py:612-622): backward -> clip_grad_norm_ -> step; the clip sees reduced gradients
    norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    clipped = [p.grad.detach().clone() for p in model.parameters()]
    opt.step()                                                     # pre-hook finish() is a no-op, post-hook resets the step
    assert all(p.grad is None for p in model.parameters())
    w1 = [p.detach().clone() for p in model.parameters()]
    for a, b, g in zip(w0, w1, clipped):
        assert torch.allclose(b, a - 0.1 * g, atol=1e-7)
    # gradient accumulation: two micro-batches, the first under no_sync()
    torch.manual_seed(50 + rank)
    xa, xb = torch.randn(3, 8), torch.randn(3, 8)
    with red.no_sync():
        model(xa).square().mean().backward()
    assert not red._finished and not any(red._launched)
    model(xb).square().mean().backward()
    assert red._finished
    acc = [p.grad.detach().clone() for p in list(model.parameters())[2:6]]
    opt.step()
    w2 = [p.detach().clone() for p in model.parameters()]
    q.put((rank, [w.numpy() for w in w0], [g.numpy() for g in grads], float(avg_loss), float(loss.detach()), float(norm),
           [w.numpy() for w in w1], [g.numpy() for g in acc], [w.numpy() for w in w2], [xa.numpy(), xb.numpy()]))
    # a second backward in the same step must not silently mix local and averaged gradients (checked last: the state after
    # the error is undefined, autograd has already accumulated into the first view when the hook raises)
    model(xa).square().mean().backward()
    try:
        model(xb).square().mean().backward()
        raised = False
    except RuntimeError as e:
        raised = 'no_sync' in str(e)
    assert raised
    be.local_barrier()
    import torch.distributed as dist
    dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, g_a, al_a, l_a, n_a, w1_a, acc_a, w2_a, xs_a), (_, w_b, g_b, al_b, l_b, n_b, w1_b, acc_b, w2_b, xs_b) = res
    import numpy as np
    for a, b in zip(w_a, w_b):
        assert np.array_equal(a, b), 'parameters were not broadcast from rank 0'
    for a, b in zip(g_a, g_b):
        assert np.allclose(a, b), 'ranks disagree on the reduced gradient'
    assert np.allclose(g_a[0], np.full(5, 1.5))                    # `extra` is registered on the container -> first; mean of 1 and 2
    assert np.allclose(g_a[1], np.full((3, 2), 15.0))              # `direct`: mean of 10 and 20, written through direct_slot/direct_done
    assert abs(al_a - (l_a + l_b) / 2) < 1e-6 and abs(al_a - al_b) < 1e-7
    # cross-check against a single-process evaluation of the same two micro-batches
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    tot = [torch.zeros_like(p) for p in model.parameters()]
    for r in range(2):
        torch.manual_seed(7 + r)
        x = torch.randn(3, 8)
        model.zero_grad()
        model(x).square().mean().backward()
        for t, p in zip(tot, model.parameters()):
            t += p.grad / 2
    for t, g in zip(tot, g_a[2:6]):
        assert np.allclose(t.numpy(), g, atol=1e-6)
    # reference-style loop: identical clip coefficient and identical weights on both ranks after the step
    assert abs(n_a - n_b) < 1e-7
    for a, b in zip(w1_a, w1_b):
        assert np.array_equal(a, b), 'replicas diverged after clip_grad_norm_ + step'
    # accumulation under no_sync(): mean over ranks of (grad(xa) + grad(xb)), identical weights afterwards
    for a, b in zip(acc_a, acc_b):
        assert np.allclose(a, b)
    for a, b in zip(w2_a, w2_b):
        assert np.array_equal(a, b)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    with torch.no_grad():
        for p, w in zip(ref.parameters(), w1_a[2:6]):
            p.copy_(torch.from_numpy(w))
    tot = [torch.zeros_like(p) for p in ref.parameters()]
    for xs in (xs_a, xs_b):
        ref.zero_grad()
        for xx in xs:
            ref(torch.from_numpy(xx)).square().mean().backward()
        for t, p in zip(tot, ref.parameters()):
            t += p.grad / 2
    for t, g in zip(tot, acc_a):
        assert np.allclose(t.numpy(), g, atol=1e-6)
