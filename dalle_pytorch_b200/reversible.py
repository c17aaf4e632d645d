"""Block executors (reference dalle_pytorch/reversible.py).

SequentialSequence : x = x + f(x); x = x + g(x)                         reversible.py:126-141
ReversibleSequence : y1 = x1 + f(x2); y2 = x2 + g(y1), O(1) activation memory; backward reconstructs
                     x2 = y2 - g(y1), x1 = y1 - f(x2) and re-runs the kernels of each sub-layer   reversible.py:54-157

Both drive the fused sub-layers of functional.py when the wrapped modules allow it (LayerScale.plan()), i.e. in
training / full-sequence evaluation; with an inference cache they run the module-by-module path.
"""
import torch
from torch import nn

from . import ops
from .functional import (attn_sublayer_forward, attn_sublayer_backward, ff_sublayer_forward, ff_sublayer_backward, _note_use,
                         chain_reset, draw_dropout)


def route_args(router, args, depth):
    """reversible.py:8-17"""
    routed_args = [(dict(), dict()) for _ in range(depth)]
    matched_keys = [key for key in args.keys() if key in router]
    for key in matched_keys:
        val = args[key]
        for d, ((f_args, g_args), routes) in enumerate(zip(routed_args, router[key])):
            new_f_args, new_g_args = map(lambda route: ({key: val} if route else {}), routes)
            routed_args[d] = ({**f_args, **new_f_args}, {**g_args, **new_g_args})
    return routed_args


class SequentialSequence(nn.Module):
    def __init__(self, layers, args_route={}, layer_dropout=0.):
        super().__init__()
        assert all(len(route) == len(layers) for route in args_route.values()), \
            'each argument route map must have the same depth as the number of sequential layers'
        self.layers = layers
        self.args_route = args_route
        self.layer_dropout = layer_dropout

    def forward(self, x, **kwargs):
        args = route_args(self.args_route, kwargs, len(self.layers))
        chain_reset()
        for (f, g), (f_args, g_args) in zip(self.layers, args):
            x = f.residual(x, **f_args)      # x + f(x): one fused sub-layer when possible
            x = g.residual(x, **g_args)
        chain_reset()
        return x


class _Sub:
    """One sub-layer of a reversible block resolved to tensors + geometry (see LayerScale.plan)."""
    __slots__ = ('kind', 'geom', 'params', 'cos_t', 'sin_t', 'key_mask')


def _draw(p, x_in):
    """The dropout (seed, offset) pair of one sub-layer call (None when its dropout is off)."""
    return draw_dropout(p.geom, p.kind, x_in, None if p.kind == 'attn' else p.params['w2'].shape[1])


def _run_fwd(p, x_in, resid, sign, save, drop=None):
    P = p.params
    if p.kind == 'attn':
        return attn_sublayer_forward(p.geom, x_in, resid, P['ln_w'], P['ln_b'], P['w_qkv'], P['w_out'], P['b_out'], P['scale'], sign,
                                     p.cos_t, p.sin_t, p.key_mask, save=save, drop=drop)
    return ff_sublayer_forward(p.geom, x_in, resid, P['ln_w'], P['ln_b'], P['w1'], P['b1'], P['w2'], P['b2'], P['scale'], sign,
                               save=save, drop=drop)


def _accumulate(param, grad):
    if grad is None or not param.requires_grad:
        return
    grad = grad.view_as(param)
    if param.grad is None:
        param.grad = grad.contiguous()
    else:
        param.grad.add_(grad)
    # data-parallel overlap: once every use of the parameter in this step has contributed, hand it to the reducer
    # (it copies into the flat buffer and launches the bucket's all-reduce when the bucket is complete)
    r = getattr(param, '_b200_reducer', None)
    if r is not None:
        param._b200_acc = getattr(param, '_b200_acc', 0) + 1
        if param._b200_acc == getattr(param, '_b200_uses', 0):
            r._on_grad(param)


def _run_bwd(p, ctx, d_out, sign, drop=None):
    """Backward of one sub-layer; parameter gradients are accumulated straight into .grad (as the reference's inner
    torch.autograd.backward calls do, reversible.py:80,93).  Returns the gradient w.r.t. the sub-layer input."""
    P = p.params
    if p.kind == 'attn':
        dx, dln_w, dln_b, dw_qkv, dw_out, db_out, dscale = attn_sublayer_backward(p.geom, ctx, d_out, P['ln_w'], P['scale'], sign,
                                                                                   p.cos_t, p.sin_t, p.key_mask, drop=drop)
        for name, gr in (('ln_w', dln_w), ('ln_b', dln_b), ('w_qkv', dw_qkv), ('w_out', dw_out), ('b_out', db_out), ('scale', dscale)):
            _accumulate(P[name], gr)
    else:
        dx, dln_w, dln_b, dw1, db1, dw2, db2, dscale = ff_sublayer_backward(p.geom, ctx, d_out, P['ln_w'], P['scale'], sign, drop=drop)
        for name, gr in (('ln_w', dln_w), ('ln_b', dln_b), ('w1', dw1), ('b1', db1), ('w2', dw2), ('b2', db2), ('scale', dscale)):
            _accumulate(P[name], gr)
    return dx


class _ReversibleFunction(torch.autograd.Function):
    """reversible.py:108-124 with the block body replaced by fused kernels.  Only the final (y1, y2) is kept -- plus, when
    dropout is active, the (seed, offset) pair every sub-layer's mask was drawn with: the recomputation in backward() passes the
    same pairs, which is this executor's form of Deterministic.record_rng / set_rng (reversible.py:20-50, 64-66, 77, 90)."""

    @staticmethod
    def forward(ctx, x, plans):
        x1 = x2 = x                                             # cat([x, x]) then chunk (reversible.py:150, 61)
        drops = []
        for pf, pg in plans:
            for pl in (pf, pg):
                _note_use(*[t for t in pl.params.values() if t is not None])
            df = _draw(pf, x2)
            x1, _ = _run_fwd(pf, x2, x1, 1.0, save=False, drop=df)       # y1 = x1 + f(x2)
            dg = _draw(pg, x1)
            x2, _ = _run_fwd(pg, x1, x2, 1.0, save=False, drop=dg)       # y2 = x2 + g(y1)
            drops.append((df, dg))
        ctx.plans, ctx.drops = plans, drops
        ctx.y1, ctx.y2 = x1, x2
        return ops.axpby(x1.mul(0.5), x2, 0.5)                  # stack(chunk(out)).mean(0) (reversible.py:157)

    @staticmethod
    def backward(ctx, d_out):
        y1, y2 = ctx.y1, ctx.y2
        dy1 = d_out.contiguous() * 0.5
        dy2 = dy1.clone()
        for (pf, pg), (df, dg) in zip(reversed(ctx.plans), reversed(ctx.drops)):
            # g: recompute g(y1) with activations kept (same dropout mask), x2 = y2 - g(y1)          (reversible.py:77-83)
            x2, gctx = _run_fwd(pg, y1, y2, -1.0, save=True, drop=dg)
            dg_in = _run_bwd(pg, gctx, dy2, 1.0, drop=dg)        # gradient of (+g) w.r.t. y1; sign of the recompute is irrelevant
            del gctx
            dx1 = ops.axpby(dy1, dg_in, 1.0)                     # dx1 = dy1 + y1.grad       (reversible.py:86)
            # f: recompute f(x2), x1 = y1 - f(x2)                                (reversible.py:90-96)
            x1, fctx = _run_fwd(pf, x2, y1, -1.0, save=True, drop=df)
            df_in = _run_bwd(pf, fctx, dx1, 1.0, drop=df)
            del fctx
            dx2 = ops.axpby(dy2, df_in, 1.0)                     # dx2 = dy2 + x2.grad       (reversible.py:99)
            y1, y2, dy1, dy2 = x1, x2, dx1, dx2
        ctx.y1 = ctx.y2 = None
        return ops.axpby(dy1, dy2, 1.0), None                   # x was duplicated: both halves flow back to it


class ReversibleBlock(nn.Module):
    """Parameter container keeping the reference's module names (blocks.{i}.f.net / blocks.{i}.g.net)."""

    def __init__(self, f, g):
        super().__init__()
        self.f = Deterministic(f)
        self.g = Deterministic(g)


class Deterministic(nn.Module):
    """reversible.py:20-50: records the RNG state before a call and replays it for the recomputation.  The fused executor
    above carries the dropout (seed, offset) pairs itself; this module gives the module-by-module path the reference's
    interface: `record_rng=True` snapshots torch's CPU / CUDA generators, `set_rng=True` runs the wrapped module under those states;
    the library's dropout kernels draw their Philox offsets from torch's CPU generator (functional.DropoutRNG), so torch dropouts
    and library dropouts both repeat their masks."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.cpu_state = None
        self.cuda_state = None

    def record_rng(self, *args):
        self.cpu_state = torch.get_rng_state()
        self.cuda_state = torch.cuda.get_rng_state() if torch.cuda.is_available() and torch.cuda.is_initialized() else None

    def forward(self, *args, record_rng=False, set_rng=False, **kwargs):
        if record_rng:
            self.record_rng(*args)
        if not set_rng:
            return self.net(*args, **kwargs)
        assert self.cpu_state is not None, 'set_rng=True before any record_rng=True call'
        devices = [torch.cuda.current_device()] if self.cuda_state is not None else []
        with torch.random.fork_rng(devices=devices, enabled=True):
            torch.set_rng_state(self.cpu_state)
            if self.cuda_state is not None:
                torch.cuda.set_rng_state(self.cuda_state)
            return self.net(*args, **kwargs)


class ReversibleSequence(nn.Module):
    def __init__(self, blocks, args_route={}):
        super().__init__()
        self.args_route = args_route
        self.blocks = nn.ModuleList([ReversibleBlock(f=f, g=g) for f, g in blocks])

    def forward(self, x, **kwargs):
        args = route_args(self.args_route, kwargs, len(self.blocks))
        plans = []
        for blk, (f_args, g_args) in zip(self.blocks, args):
            pf = blk.f.net.plan(x, **f_args)
            pg = blk.g.net.plan(x, **g_args)
            if pf is None or pg is None:
                raise NotImplementedError('reversible executor needs fusable sub-layers (no inference cache, no sandwich norm)')
            plans.append((pf, pg))
        x = x.float().contiguous()
        if torch.is_grad_enabled():
            return _ReversibleFunction.apply(x, plans)
        x1 = x2 = x
        for pf, pg in plans:
            x1, _ = _run_fwd(pf, x2, x1, 1.0, save=False, drop=_draw(pf, x2))
            x2, _ = _run_fwd(pg, x1, x2, 1.0, save=False, drop=_draw(pg, x1))
        return ops.axpby(x1.mul(0.5), x2, 0.5)
