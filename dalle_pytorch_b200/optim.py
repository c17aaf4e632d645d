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
from .functional import invalidate_weight_cache


class FusedAdam(torch.optim.Optimizer):
    """Adam with optional global-norm gradient clipping over flat buffers.

    usage:  opt = FusedAdam(model.parameters(), lr=3e-4, max_grad_norm=0.5)            # single process
            opt = FusedAdam(model.parameters(), reducer=model.grad_reducer, ...)       # after NCCLBackend.distribute(...)
            loss.backward(); opt.step()          # step() finishes the gradient reduction, updates, and resets the gradients

    A `torch.optim.Optimizer`: one entry in `param_groups` whose `lr / betas / eps / weight_decay / max_grad_norm` are read at
    every step (so `ReduceLROnPlateau` and the other schedulers of the reference trainer, train_dalle.py:449-459, can drive
    it) and the step pre/post hooks `NCCLBackend.distribute(optimizer=...)` registers.  Checkpoints: `state_dict()` stores the
    two flat moment buffers, not torch.optim.Adam's per-parameter `exp_avg / exp_avg_sq` -- an `opt_state` saved by the
    reference trainer does not load into this class (and vice versa)."""

    def __init__(self, params, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None, reducer=None):
        params = [p for p in params if p.requires_grad]
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), 'FusedAdam needs fp32 CUDA parameters'
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                      max_grad_norm=float(max_grad_norm) if max_grad_norm else 0.0))
        self.reducer = reducer if reducer is not None else GradAllReducer(params)
        assert set(self.reducer.params) == set(params), 'the reducer must cover exactly the optimised parameters'
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

    # hyper-parameters live in param_groups[0] (schedulers write there); these properties keep the plain attribute access
    lr = property(lambda self: self.param_groups[0]['lr'])
    betas = property(lambda self: self.param_groups[0]['betas'])
    eps = property(lambda self: self.param_groups[0]['eps'])
    weight_decay = property(lambda self: self.param_groups[0]['weight_decay'])
    max_grad_norm = property(lambda self: self.param_groups[0]['max_grad_norm'])

    def add_param_group(self, param_group):
        if getattr(self, 'param_groups', None):
            raise NotImplementedError('FusedAdam keeps every parameter in one flat buffer: a single param group')
        super().add_param_group(param_group)

    def zero_grad(self, set_to_none=True):
        self.reducer.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.reducer.finish()                 # adopt / all-reduce whatever the hooks have not handled yet (no-op if already done)
        self.steps += 1
        g = self.reducer.flat
        hp = self.param_groups[0]
        clip = float(hp['max_grad_norm'] or 0.0)
        if clip > 0:
            self.gnorm_sq.zero_()
            ops.sumsq_(g, self.gnorm_sq)
        ops.adam_(self.flat_p, g, self.m, self.v, self.steps, float(hp['lr']), hp['betas'][0], hp['betas'][1], hp['eps'],
                  hp['weight_decay'], clip, self.gnorm_sq if clip > 0 else None)
        invalidate_weight_cache()             # the kernel wrote the parameters through raw pointers: cached bf16 copies are stale
        self.reducer.zero_grad()
        return loss

    def grad_norm(self):
        """Global gradient norm of the last step (device tensor; only tracked when clipping is on)."""
        return self.gnorm_sq.sqrt()

    _HYPER = ('lr', 'betas', 'eps', 'weight_decay', 'max_grad_norm')

    def state_dict(self):
        sd = {'steps': self.steps, 'm': self.m, 'v': self.v}
        sd.update({k: self.param_groups[0][k] for k in self._HYPER})
        return sd

    def load_state_dict(self, sd):
        if 'state' in sd and 'param_groups' in sd:
            raise ValueError('this is a torch.optim state dict (per-parameter exp_avg / exp_avg_sq); FusedAdam stores two flat '
                             'moment buffers and cannot load it')
        for k in ('m', 'v'):
            if tuple(sd[k].shape) != tuple(self.m.shape):
                raise ValueError(f"FusedAdam.load_state_dict: '{k}' has {tuple(sd[k].shape)} elements, this model's flat buffer "
                                 f'has {tuple(self.m.shape)} (different parameter set or order)')
        self.steps = int(sd['steps'])
        self.m.copy_(sd['m'])
        self.v.copy_(sd['v'])
        for k in self._HYPER:
            if k in sd:
                self.param_groups[0][k] = tuple(sd[k]) if k == 'betas' else sd[k]
