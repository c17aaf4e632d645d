"""Optimizer step of the reference trainer on the library kernels.

train_dalle.py:617-619 does `clip_grad_norm_(dalle.parameters(), 0.5)` and `Adam.step()` (train_dalle.py:441: `Adam(lr=3e-4)`).
Here every parameter is a view into ONE flat fp32 buffer, the gradients live in the flat buffer of
`distributed.GradAllReducer` (the data-parallel all-reduce buffer; with a single process it is just a flat gradient buffer),
and the whole step is two launches: `dalle_b200_sumsq` (gradient norm) and `dalle_b200_adam` (clip coefficient read on the
device + Adam update of p, m, v) -- no host synchronisation, no per-parameter kernels.
"""
import torch

from . import ops
from .distributed import GradAllReducer


class FusedAdam:
    """Adam with optional global-norm gradient clipping over flat buffers.

    usage:  opt = FusedAdam(model.parameters(), lr=3e-4, max_grad_norm=0.5)            # single process
            opt = FusedAdam(model.parameters(), reducer=model.grad_reducer, ...)       # after NCCLBackend.distribute(...)
            loss.backward(); opt.step()          # step() finishes the gradient reduction, updates, and resets the gradients
    """

    def __init__(self, params, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None, reducer=None):
        params = [p for p in params if p.requires_grad]
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), 'FusedAdam needs fp32 CUDA parameters'
        self.reducer = reducer if reducer is not None else GradAllReducer(params)
        assert set(self.reducer.params) == set(params), 'the reducer must cover exactly the optimised parameters'
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        # parameters re-homed into one flat buffer laid out exactly like the reducer's gradient buffer
        flat_g = self.reducer.flat
        self.flat_p = torch.empty_like(flat_g)
        base = flat_g.data_ptr()
        with torch.no_grad():
            for p in self.reducer.params:
                gv = self.reducer.views[p]
                off = (gv.data_ptr() - base) // 4
                pv = self.flat_p[off:off + p.numel()].view_as(p)
                pv.copy_(p.data)
                p.data = pv
        self.m = torch.zeros_like(flat_g)
        self.v = torch.zeros_like(flat_g)
        self.gnorm_sq = torch.zeros(1, device=flat_g.device, dtype=torch.float32)
        self.steps = 0

    def zero_grad(self, set_to_none=True):
        self.reducer.zero_grad()

    @torch.no_grad()
    def step(self):
        self.reducer.finish()                 # adopt / all-reduce whatever the hooks have not handled yet
        self.steps += 1
        g = self.reducer.flat
        if self.max_grad_norm > 0:
            self.gnorm_sq.zero_()
            ops.sumsq_(g, self.gnorm_sq)
        ops.adam_(self.flat_p, g, self.m, self.v, self.steps, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                  self.max_grad_norm, self.gnorm_sq if self.max_grad_norm > 0 else None)
        self.reducer.zero_grad()

    def grad_norm(self):
        """Global gradient norm of the last step (device tensor; only tracked when clipping is on)."""
        return self.gnorm_sq.sqrt()

    def state_dict(self):
        return {'steps': self.steps, 'm': self.m, 'v': self.v, 'lr': self.lr, 'betas': self.betas, 'eps': self.eps,
                'weight_decay': self.weight_decay, 'max_grad_norm': self.max_grad_norm}

    def load_state_dict(self, sd):
        self.steps = sd['steps']
        self.m.copy_(sd['m'])
        self.v.copy_(sd['v'])
