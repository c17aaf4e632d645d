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
    torch.manual_seed(7 + rank)
    x = torch.randn(3, 8)
    loss = model(x).square().mean()
    loss.backward()
    red = model.grad_reducer
    direct._b200_uses = 1                                          # what functional._note_use records in forward
    slot = red.direct_slot(direct)
    assert slot is not None and slot.data_ptr() == red.views[direct].data_ptr()
    slot.copy_(torch.full((3, 2), 10.0 * (rank + 1)))              # stands in for the weight-gradient GEMM
    red.direct_done(direct)
    assert red.direct_slot(direct) is None                         # a second use in the same step must go through autograd
    assert extra.grad is None                                      # gradients start each step unset (no memset of the flat buffer)
    extra.grad = torch.full((5,), float(rank + 1))                 # written outside autograd; finish() adopts it into the flat buffer
    model.grad_reducer.finish()
    grads = [p.grad.detach().clone() for p in model.parameters()]
    avg_loss = be.average_all(loss.detach())
    assert all(p.grad.data_ptr() == model.grad_reducer.views[p].data_ptr() for p in model.parameters())
    opt.step()                                                     # pre-hook finish() is idempotent, post-hook resets the step
    assert all(p.grad is None for p in model.parameters())
    q.put((rank, [w.numpy() for w in w0], [g.numpy() for g in grads], float(avg_loss), float(loss.detach())))
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
    (_, w_a, g_a, al_a, l_a), (_, w_b, g_b, al_b, l_b) = res
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
