"""Fused sub-layers of the DALL-E transformer block, as sequences of libdalle_b200 kernels.

One attention sub-layer  = LayerScale(PreNorm(PreShiftToken(Attention*)))   (reference transformer.py:279-292)
One feed-forward sub-layer = LayerScale(PreNorm(PreShiftToken(FeedForward)))

    out = resid + sign * scale * F( Shift( LN(x_in) ) )

`x_in` and `resid` are separate so that the same code serves the sequential executor (x_in is resid,
reversible.py:139-140) and the reversible one (y1 = x1 + f(x2), reversible.py:64-66).  Forward and backward are
plain functions over tensors (`*_forward` returns a context tuple, `*_backward` consumes it); thin
torch.autograd.Function wrappers expose them to autograd, and the reversible executor calls them directly to
recompute activations block by block.

Kernel sequence, attention sub-layer (bf16 mode: every GEMM is the tcgen05 kernel):
  fwd  ln_shift_fwd -> gemm[QKV: rotary+scale+head split] -> attn_fwd -> gemm[RESID: bias+LayerScale+residual]
  bwd  scale_bwd -> gemm dgrad(to_out) , gemm wgrad(to_out) -> attn_bwd -> gemm dgrad(to_qkv), gemm wgrad(to_qkv)
       -> ln_shift_bwd (adds the residual-branch gradient)
feed-forward sub-layer:
  fwd  ln_shift_fwd -> gemm[GEGLU: bias + a*gelu(g)] -> gemm[RESID]
  bwd  scale_bwd -> gemm dgrad(net.3)[GEGLU_BWD] , gemm wgrad(net.3) -> colsum (b1) -> gemm dgrad(net.0), gemm wgrad(net.0)
       -> ln_shift_bwd
"""
import os
import threading

import torch

from . import ops


FUSE_GEGLU_BWD = False      # True: dgrad(net.3) with the GEGLU adjoint fused in its epilogue (EPI_GEGLU_BWD)
# True: the LayerNorm/shift backward kernel of sub-layer k also forms the LayerScale adjoint of sub-layer k-1 (whose output is
# this sub-layer's input) while dx is in registers, instead of a separate scale_bwd pass that re-reads dx (d = 1024 only).
FUSE_UPSTREAM_SCALE_BWD = os.environ.get('DALLE_B200_FUSE_SCALE_BWD', '1') != '0'


class _SubRec:
    """What sub-layer k must know about sub-layer k-1 to run its LayerScale adjoint, plus the slot the results come back in.

    Identity of "the tensor sub-layer k-1 produced / the gradient sub-layer k wrote" is established with the tensor itself held
    alive (so the caching allocator cannot hand its address to anything else) plus its version counter (so an in-place
    accumulation by the autograd engine's input buffer -- a second consumer of the stream tensor -- is detected): a bare
    data_ptr() comparison would accept a recycled address or `dx + other` sitting at dx's address."""
    __slots__ = ('y', 'scale', 'sign', 'out', 'out_version', 'pre', 'pre_dx', 'pre_version')

    def __init__(self, y, scale, sign, out):
        self.y, self.scale, self.sign = y, scale, sign
        self.out, self.out_version = out, out._version       # released as soon as the next sub-layer has linked (or chain_reset)
        self.pre, self.pre_dx, self.pre_version = None, None, -1

    def produced(self, x):
        """Is `x` the (unmodified) output this record's sub-layer returned?"""
        o = self.out
        return (o is not None and x.data_ptr() == o.data_ptr() and x._version == self.out_version and x.shape == o.shape
                and x.dtype == o.dtype)

    def take_pre(self, d_out):
        """The pre-computed LayerScale adjoint, if `d_out` is exactly the dx it was formed from."""
        pre, dx = self.pre, self.pre_dx
        self.pre = self.pre_dx = None
        if pre is None or dx is None:
            return None
        if d_out.data_ptr() != dx.data_ptr() or d_out._version != self.pre_version or d_out.shape != dx.shape:
            return None
        return pre


_chain = threading.local()


def chain_reset():
    """Called at the start and the end of every executor forward: drops the tail record's reference to its output (a record
    is reachable from its own output through grad_fn -> ctx, so the reference must not outlive the link step)."""
    last = getattr(_chain, 'last', None)
    if last is not None:
        last.out = None
    _chain.last = None


def _chain_link(x_in, resid_is_input, rec):
    """Returns the record of the sub-layer that produced `x_in` (if it is the immediately preceding fused sub-layer and x_in is
    also the residual, i.e. the sequential executor) and makes `rec` the new tail."""
    prev = getattr(_chain, 'last', None)
    _chain.last = rec
    if prev is None:
        return None
    ok = FUSE_UPSTREAM_SCALE_BWD and resid_is_input and x_in.shape[-1] == 1024 and x_in.is_cuda and prev.produced(x_in)
    prev.out = None                  # the stream tensor is owned by the executor from here on
    return prev if ok else None


class DropoutRNG:
    """(seed, offset) pairs for dalle_b200_dropout.  The seed is torch's (torch.manual_seed), the offset -- the first Philox
    counter of the call -- is a 62-bit draw from torch's CPU generator: a run is reproducible from its seed exactly like torch's
    own dropout, every dropout site gets its own counter range, and restoring the CPU generator state (reversible.py:20-50,
    Deterministic.set_rng) repeats the draw.  The pair a forward pass drew is kept by its autograd node / the reversible executor and
    passed again to the backward pass and to the recomputation."""

    @staticmethod
    def draw(numel):
        return (torch.initial_seed(), int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item()))


def draw_dropout(g, kind, x_in, hidden=None):
    """The (seed, offset) pair of one sub-layer call, or None when its dropout is off."""
    if g.p_drop <= 0:
        return None
    b, n, d = x_in.shape
    return DropoutRNG.draw(b * n * (d if kind == 'attn' else hidden))


class SublayerGeom:
    """Static description of a sub-layer (everything that is not a tensor).  p_drop > 0: dropout (training) after the output
    projection of attention (attention.py:53-56) / after GEGLU in the feed-forward (transformer.py:117)."""

    def __init__(self, *, dtype, text_len=0, fmap=0, do_ln=True, do_shift=False, heads=0, dim_head=64, attn_spec=None,
                 q_scale=None, eps=1e-5, p_drop=0.0):
        self.p_drop = float(p_drop)
        self.dtype = dtype
        self.text_len, self.fmap = text_len, fmap
        self.do_ln, self.do_shift = do_ln, do_shift
        self.heads, self.dim_head = heads, dim_head
        self.attn_spec = attn_spec
        self.q_scale = q_scale if q_scale is not None else dim_head ** -0.5
        self.eps = eps

    def shift_active(self, n):
        # transformer.py:160-161: sequences shorter than the text length are not shifted
        return self.do_shift and n >= self.text_len


# bf16 copies of the fp32 master weights, kept ON the parameter object and valid for one value of its version counter: a weight
# is cast once after each in-place update (optimizer step, load_state_dict) instead of once per use -- 49 cast launches and 1.4 GB
# of traffic per C2 step otherwise.  (The copy lives and dies with the parameter: a cache keyed by address would hand a new
# model the copy of a freed one.)  `weight_cache_enabled(False)` (used while a training step that UPDATES the weights is captured
# into a CUDA graph: Python does not run at replay, so a cached copy would go stale) forces the cast at every use.
_wcache_on = True
_wcache_epoch = 0


def weight_cache_enabled(flag):
    global _wcache_on
    old, _wcache_on = _wcache_on, bool(flag)
    return old


def invalidate_weight_cache():
    """For writers that update parameters through raw pointers (FusedAdam's kernel, collectives): torch's version counter does
    not see those writes."""
    global _wcache_epoch
    _wcache_epoch += 1


def _w(weight, dtype):
    """Weights are stored in fp32 (reference checkpoints); bf16 mode uses the bf16 copy cached on the parameter (see above)."""
    if dtype == torch.float32:
        return weight.detach().contiguous()
    if not _wcache_on or not weight.is_leaf:
        return ops.cast_bf16(weight.detach().contiguous())
    hit = getattr(weight, '_b200_bf16', None)
    if hit is not None and hit[0] == (weight._version, _wcache_epoch, weight.data_ptr()):
        return hit[1]
    c = ops.cast_bf16(weight.detach().contiguous())
    weight._b200_bf16 = ((weight._version, _wcache_epoch, weight.data_ptr()), c)
    return c


def _up_args(up, da):
    """(up_y, up_scale, up_sign) for ops.ln_shift_bwd, or None when the upstream record cannot be fused with this call."""
    if up is None:
        return None
    if up.y is not None and up.y.dtype != da.dtype:
        return None
    return (up.y, up.scale, up.sign)


def _up_store(up, res):
    """Files the fused results in the upstream sub-layer's record; returns dx."""
    if not isinstance(res, tuple):
        return res
    dx, up_dy, up_dscale, up_dbias = res
    up.pre, up.pre_dx, up.pre_version = (up_dy, up_dscale, up_dbias), dx, dx._version
    return dx


# =====================================================================================================
# attention sub-layer
# =====================================================================================================
def attn_sublayer_forward(g: SublayerGeom, x_in, resid, ln_w, ln_b, w_qkv, w_out, b_out, scale, sign, cos_t, sin_t,
                          key_mask=None, save=True, drop=None):
    """drop = (seed, offset) from DropoutRNG.draw(b*n*d) when g.p_drop > 0 (None = no dropout in this call)."""
    b, n, d = x_in.shape
    x_in = x_in.contiguous()
    shift = g.shift_active(n)
    a1, mean, rstd = ops.ln_shift_fwd(x_in, ln_w, ln_b, g.dtype, g.text_len, g.fmap, do_ln=g.do_ln, do_shift=shift, eps=g.eps)
    wq, wo = _w(w_qkv, g.dtype), _w(w_out, g.dtype)
    lay = ops.gather_layout(g.attn_spec, g.dtype, n, g.dim_head, key_mask)      # axial row / column on the gathered kernels
    q, k, v = ops.gemm_qkv_auto(a1, wq, b, n, g.heads, g.dim_head, cos_t, sin_t, g.q_scale, n_alloc=None if lay is None else lay.n_alloc)
    o, lse = ops.attn_fwd(g.attn_spec, q, k, v, key_mask, lay=lay)
    keep_y = save and scale is not None
    r2 = None if resid is None else resid.contiguous().view(b * n, d)
    sc = None if scale is None else scale.detach().reshape(-1).contiguous()
    if drop is None:
        out, y = ops.gemm_resid(o.view(b * n, -1), wo, b_out, r2, sc, sign, keep_y=keep_y)
    else:       # to_out -> Dropout -> LayerScale -> residual: the mask sits between the projection and the scale
        y = ops.dropout_(ops.gemm_store(o.view(b * n, -1), wo, bias=b_out), g.p_drop, *drop)
        out = ops.resid_scale(y, r2, sc, sign)
        if not keep_y:
            y = None
    out = out.view(b, n, d)
    ctx = None
    if save:
        ctx = (x_in, mean, rstd, a1, wq, wo, q, k, v, o, lse, y, shift, lay)
    return out, ctx


def attn_sublayer_backward(g: SublayerGeom, ctx, d_out, ln_w, scale, sign, cos_t, sin_t, key_mask=None, dres=None, wslots=None,
                           pre=None, up=None, drop=None):
    """d_out: gradient w.r.t. `out` [b,n,d] fp32.  Returns (dx_in, dln_w, dln_b, dw_qkv, dw_out, db_out, dscale).
    `dres` (optional, fp32) is added to dx_in inside the LayerNorm-backward kernel (sequential executor: the residual
    branch gradient, which equals d_out).  `wslots` (optional) = (dw_qkv_out, dw_out_out): preallocated fp32 destinations
    (views into the data-parallel flat gradient buffer) the weight-gradient GEMMs write straight into."""
    x_in, mean, rstd, a1, wq, wo, q, k, v, o, lse, y, shift, lay = ctx
    s_qkv, s_out = wslots if wslots is not None else (None, None)
    b, n, d = x_in.shape
    M = b * n
    d_out = d_out.contiguous().view(M, d)
    sc = None if scale is None else scale.detach().reshape(-1).contiguous()
    pool = torch.zeros(6, d, device=x_in.device, dtype=torch.float32)     # one fill for dscale, db_out, dln_w, dln_b (+ upstream dscale, dbias)
    if drop is not None:     # adjoint of the dropout between to_out and the LayerScale: same mask on the gradient; the bias sits before it
        dy, dscale, _ = ops.scale_bwd(d_out, y, sc, sign, g.dtype, want_dbias=False, zeroed=(pool[0], pool[1]))
        ops.dropout_(dy, g.p_drop, *drop)
        db_out = ops.colsum(dy)
    elif pre is not None:      # already formed by the downstream sub-layer's LayerNorm backward (FUSE_UPSTREAM_SCALE_BWD)
        dy, dscale, db_out = pre
    else:
        dy, dscale, db_out = ops.scale_bwd(d_out, y, sc, sign, g.dtype, zeroed=(pool[0], pool[1]))
    d_o = ops.gemm_store(dy, wo, a_mn=False, b_mn=True, out=ops.attn_dout_buffer(lay, M, wo.shape[1], dy.device, dy.dtype))   # [M, inner]
    dw_out = ops.gemm_store(dy, o.view(M, -1), a_mn=True, b_mn=True, out_dtype=torch.float32, out=s_out)   # [d, inner]
    dqkv = ops.attn_bwd(g.attn_spec, q, k, v, o, lse, d_o.view(b, n, -1), cos_t, sin_t, g.q_scale, key_mask, lay=lay)
    da1 = ops.gemm_store(dqkv, wq, a_mn=False, b_mn=True)                                 # [M, d]
    dw_qkv = ops.gemm_store(dqkv, a1, a_mn=True, b_mn=True, out_dtype=torch.float32, out=s_qkv)      # [3*inner, d]
    dln_w = dln_b = None
    if g.do_ln:
        dln_w, dln_b = pool[2], pool[3]
    dx = ops.ln_shift_bwd(da1, x_in, mean, rstd, ln_w, None if dres is None else dres.contiguous(), g.text_len, g.fmap,
                          do_ln=g.do_ln, do_shift=shift, dgamma=dln_w, dbeta=dln_b, up=_up_args(up, da1), up_zeroed=pool[4:6])
    if up is not None:
        dx = _up_store(up, dx)
    if dscale is not None:
        dscale = dscale.view_as(scale)
    return dx, dln_w, dln_b, dw_qkv, dw_out, db_out, dscale


# =====================================================================================================
# feed-forward sub-layer
# =====================================================================================================
def ff_sublayer_forward(g: SublayerGeom, x_in, resid, ln_w, ln_b, w1, b1, w2, b2, scale, sign, save=True, drop=None):
    """drop = (seed, offset) from DropoutRNG.draw(b*n*hidden) when g.p_drop > 0."""
    b, n, d = x_in.shape
    x_in = x_in.contiguous()
    shift = g.shift_active(n)
    a2, mean, rstd = ops.ln_shift_fwd(x_in, ln_w, ln_b, g.dtype, g.text_len, g.fmap, do_ln=g.do_ln, do_shift=shift, eps=g.eps)
    w1c, w2c = _w(w1, g.dtype), _w(w2, g.dtype)
    h, u = ops.gemm_geglu(a2, w1c, b1, keep_u=save)
    if drop is not None:
        ops.dropout_(h, g.p_drop, *drop)          # net.2 (transformer.py:117); h is only kept for the net.3 weight gradient
    keep_y = save and scale is not None
    out, y = ops.gemm_resid(h, w2c, b2, None if resid is None else resid.contiguous().view(b * n, d),
                            None if scale is None else scale.detach().reshape(-1).contiguous(), sign, keep_y=keep_y)
    out = out.view(b, n, d)
    ctx = None
    if save:
        ctx = (x_in, mean, rstd, a2, w1c, w2c, u, h, y, shift)
    return out, ctx


def ff_sublayer_backward(g: SublayerGeom, ctx, d_out, ln_w, scale, sign, dres=None, wslots=None, pre=None, up=None, drop=None):
    """Returns (dx_in, dln_w, dln_b, dw1, db1, dw2, db2, dscale).  `wslots` = (dw1_out, dw2_out), see attn_sublayer_backward."""
    x_in, mean, rstd, a2, w1c, w2c, u, h, y, shift = ctx
    s_w1, s_w2 = wslots if wslots is not None else (None, None)
    b, n, d = x_in.shape
    M = b * n
    d_out = d_out.contiguous().view(M, d)
    sc = None if scale is None else scale.detach().reshape(-1).contiguous()
    H2 = u.shape[1]
    pool = torch.zeros(6 * d + H2, device=x_in.device, dtype=torch.float32)   # one fill for dscale, db2, dln_w, dln_b, upstream dscale/dbias, db1
    if pre is not None:
        dy, dscale, db2 = pre
    else:
        dy, dscale, db2 = ops.scale_bwd(d_out, y, sc, sign, g.dtype, zeroed=(pool[:d], pool[d:2 * d]))
    if FUSE_GEGLU_BWD:
        du = ops.gemm_geglu_bwd(dy, w2c, u)                                               # [M, 2H]
        db1 = ops.colsum(du)
    else:   # measured faster on B200 (profiles/): plain dgrad GEMM + one streaming pass that also forms the bias gradient
        dh = ops.gemm_store(dy, w2c, a_mn=False, b_mn=True)                               # [M, H]
        if drop is not None:
            ops.dropout_(dh, g.p_drop, *drop)                                             # same mask as the forward
        du, db1 = ops.geglu_bwd(dh, u, zeroed=pool[6 * d:])
    dw2 = ops.gemm_store(dy, h, a_mn=True, b_mn=True, out_dtype=torch.float32, out=s_w2)            # [d, H]
    da2 = ops.gemm_store(du, w1c, a_mn=False, b_mn=True)                                  # [M, d]
    dw1 = ops.gemm_store(du, a2, a_mn=True, b_mn=True, out_dtype=torch.float32, out=s_w1)           # [2H, d]
    dln_w = dln_b = None
    if g.do_ln:
        dln_w, dln_b = pool[2 * d:3 * d], pool[3 * d:4 * d]
    dx = ops.ln_shift_bwd(da2, x_in, mean, rstd, ln_w, None if dres is None else dres.contiguous(), g.text_len, g.fmap,
                          do_ln=g.do_ln, do_shift=shift, dgamma=dln_w, dbeta=dln_b, up=_up_args(up, da2),
                          up_zeroed=pool[4 * d:6 * d].view(2, d))
    if up is not None:
        dx = _up_store(up, dx)
    if dscale is not None:
        dscale = dscale.view_as(scale)
    return dx, dln_w, dln_b, dw1, db1, dw2, db2, dscale


# =====================================================================================================
# autograd wrappers
# =====================================================================================================
def _note_use(*params):
    """Forward-time use count per parameter (reset by GradAllReducer.zero_grad): a weight used by ONE sub-layer in the
    step may have its gradient GEMM write directly into the flat data-parallel buffer; shared weights
    (shared_attn_ids / shared_ff_ids, transformer.py:261-292) go through autograd's accumulation."""
    for p in params:
        if getattr(p, '_b200_reducer', None) is not None:
            p._b200_uses = getattr(p, '_b200_uses', 0) + 1


def _slot(p):
    r = getattr(p, '_b200_reducer', None)
    if r is None or getattr(p, '_b200_uses', 0) != 1:
        return None
    return r.direct_slot(p)


def _commit(p, slot, grad):
    """grad was written into `slot` (the flat-buffer view): tell the reducer and hand autograd nothing."""
    if slot is None:
        return grad
    p._b200_reducer.direct_done(p)
    return None


class AttnSublayerFn(torch.autograd.Function):
    """out = resid + sign*scale*Attn(Shift(LN(x_in))).  If `resid_is_input`, resid := x_in and the residual-branch
    gradient is fused into the LayerNorm backward kernel."""

    @staticmethod
    def forward(ctx, g, resid_is_input, sign, cos_t, sin_t, key_mask, x_in, resid, ln_w, ln_b, w_qkv, w_out, b_out, scale):
        r = x_in if resid_is_input else resid
        ctx.drop = DropoutRNG.draw(x_in.numel()) if g.p_drop > 0 else None
        out, saved = attn_sublayer_forward(g, x_in, r, ln_w, ln_b, w_qkv, w_out, b_out, scale, sign, cos_t, sin_t, key_mask,
                                           save=True, drop=ctx.drop)
        ctx.g, ctx.resid_is_input, ctx.sign = g, resid_is_input, sign
        ctx.cos_t, ctx.sin_t, ctx.key_mask = cos_t, sin_t, key_mask
        ctx.saved = saved
        ctx.ln_w, ctx.scale = ln_w, scale
        ctx.has_resid = resid is not None
        ctx.wparams = (w_qkv, w_out)
        _note_use(w_qkv, w_out)
        ctx.rec = ctx.up = None
        if ctx.drop is not None:
            chain_reset()            # the LayerScale adjoint is not a plain function of dx when a dropout mask sits in between
        elif any(ctx.needs_input_grad) and saved is not None:
            y = saved[11]
            ctx.rec = _SubRec(y, None if scale is None else scale.detach().reshape(-1).contiguous(), sign, out)
            ctx.up = _chain_link(x_in, resid_is_input, ctx.rec)
        return out

    @staticmethod
    def backward(ctx, d_out):
        g = ctx.g
        d_out = d_out.contiguous()
        dres = d_out if ctx.resid_is_input else None
        slots = tuple(_slot(p) for p in ctx.wparams)
        rec = ctx.rec
        pre = rec.take_pre(d_out) if rec is not None else None
        dx, dln_w, dln_b, dw_qkv, dw_out, db_out, dscale = attn_sublayer_backward(
            g, ctx.saved, d_out, ctx.ln_w, ctx.scale, ctx.sign, ctx.cos_t, ctx.sin_t, ctx.key_mask, dres=dres, wslots=slots,
            pre=pre, up=ctx.up, drop=ctx.drop)
        ctx.rec = ctx.up = None
        dw_qkv = _commit(ctx.wparams[0], slots[0], dw_qkv)
        dw_out = _commit(ctx.wparams[1], slots[1], dw_out)
        ctx.saved = None
        d_resid = d_out if (ctx.has_resid and not ctx.resid_is_input) else None
        return (None, None, None, None, None, None, dx, d_resid, dln_w, dln_b, dw_qkv, dw_out, db_out, dscale)


class FFSublayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, resid_is_input, sign, x_in, resid, ln_w, ln_b, w1, b1, w2, b2, scale):
        r = x_in if resid_is_input else resid
        ctx.drop = DropoutRNG.draw(x_in.shape[0] * x_in.shape[1] * w2.shape[1]) if g.p_drop > 0 else None
        out, saved = ff_sublayer_forward(g, x_in, r, ln_w, ln_b, w1, b1, w2, b2, scale, sign, save=True, drop=ctx.drop)
        ctx.g, ctx.resid_is_input, ctx.sign = g, resid_is_input, sign
        ctx.saved = saved
        ctx.ln_w, ctx.scale = ln_w, scale
        ctx.has_resid = resid is not None
        ctx.wparams = (w1, w2)
        _note_use(w1, w2)
        ctx.rec = ctx.up = None
        if any(ctx.needs_input_grad) and saved is not None:
            y = saved[8]
            ctx.rec = _SubRec(y, None if scale is None else scale.detach().reshape(-1).contiguous(), sign, out)
            ctx.up = _chain_link(x_in, resid_is_input, ctx.rec)
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        dres = d_out if ctx.resid_is_input else None
        slots = tuple(_slot(p) for p in ctx.wparams)
        rec = ctx.rec
        pre = rec.take_pre(d_out) if rec is not None else None
        dx, dln_w, dln_b, dw1, db1, dw2, db2, dscale = ff_sublayer_backward(ctx.g, ctx.saved, d_out, ctx.ln_w, ctx.scale, ctx.sign,
                                                                              dres=dres, wslots=slots, pre=pre, up=ctx.up, drop=ctx.drop)
        ctx.rec = ctx.up = None
        dw1 = _commit(ctx.wparams[0], slots[0], dw1)
        dw2 = _commit(ctx.wparams[1], slots[1], dw2)
        ctx.saved = None
        d_resid = d_out if (ctx.has_resid and not ctx.resid_is_input) else None
        return (None, None, None, dx, d_resid, dln_w, dln_b, dw1, db1, dw2, db2, dscale)


class EmbedTokensFn(torch.autograd.Function):
    """tokens = cat(text_emb(text_ids), image_emb(image_ids)) (dalle_pytorch.py:616-630) as two gathers writing one [b, n, d]
    buffer; the backward scatter-adds the token gradients into the two tables with fp32 atomics."""

    @staticmethod
    def forward(ctx, text_ids, image_ids, w_text, w_image):
        B, Lt = text_ids.shape
        Li = 0 if image_ids is None else image_ids.shape[1]
        d = w_text.shape[1]
        out = torch.empty(B, Lt + Li, d, device=w_text.device, dtype=torch.float32)
        ops.embed_fwd(text_ids, w_text.detach().float(), out, 0)
        if Li:
            ops.embed_fwd(image_ids, w_image.detach().float(), out, Lt)
        ctx.save_for_backward(text_ids, image_ids)
        ctx.shapes = (w_text.shape, None if w_image is None else w_image.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        text_ids, image_ids = ctx.saved_tensors
        d_out = d_out.contiguous().float()
        Lt = text_ids.shape[1]
        dw_text = torch.zeros(ctx.shapes[0], device=d_out.device, dtype=torch.float32)
        ops.embed_bwd(text_ids, d_out, dw_text, 0)
        dw_image = None
        if image_ids is not None and image_ids.shape[1] > 0:
            dw_image = torch.zeros(ctx.shapes[1], device=d_out.device, dtype=torch.float32)
            ops.embed_bwd(image_ids, d_out, dw_image, Lt)
        return None, None, dw_text, dw_image


class LayerNormFn(torch.autograd.Function):
    """Plain LayerNorm on the ln_shift kernels (used for sandwich norm, transformer.py:96,102)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        x3 = x.contiguous().view(1, -1, shp[-1]).float()
        out, mean, rstd = ops.ln_shift_fwd(x3, w, b, torch.float32, 0, 1, do_ln=True, do_shift=False, eps=eps)
        ctx.save_for_backward(x3, mean, rstd, w)
        return out.view(shp)

    @staticmethod
    def backward(ctx, d_out):
        x3, mean, rstd, w = ctx.saved_tensors
        d = x3.shape[-1]
        dw = torch.zeros(d, device=x3.device, dtype=torch.float32)
        db = torch.zeros(d, device=x3.device, dtype=torch.float32)
        dx = ops.ln_shift_bwd(d_out.contiguous().view(-1, d).float(), x3, mean, rstd, w, None, 0, 1, do_ln=True, do_shift=False,
                              dgamma=dw, dbeta=db)
        return dx.view(d_out.shape), dw, db, None


class HeadLossFn(torch.autograd.Function):
    """to_logits (LayerNorm + Linear) + the weighted text/image cross-entropy of dalle_pytorch.py:644-671 on the library
    kernels, restricted to the live vocabulary of each position class (see DALLE._loss_head): rows of `x` must be ordered
    [all text positions | all image positions].
        loss = (CE_text + w * CE_img) / (w + 1)
    forward : ln_shift_fwd -> 2 GEMMs (+bias) -> ce_fwd (one pass, keeps row log-sum-exp)
    backward: ce_bwd (in place over the logits) -> dgrad / wgrad GEMMs + colsum -> ln_shift_bwd"""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w, bias, labels_text, labels_img, ntt, w_img, dtype, eps):
        M, d = x.shape
        Rt, Ri = labels_text.numel(), labels_img.numel()
        assert Rt + Ri == M
        x3 = x.contiguous().view(1, M, d)
        h, mean, rstd = ops.ln_shift_fwd(x3, ln_w, ln_b, dtype, 0, 1, do_ln=True, do_shift=False, eps=eps)
        wc = _w(w, dtype)
        bias = bias.detach()
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        ct, ci = 1.0 / ((w_img + 1) * max(Rt, 1)), w_img / ((w_img + 1) * max(Ri, 1))
        lt = li = lse_t = lse_i = None
        if Rt:
            lt = ops.gemm_store(h[:Rt], wc[:ntt], bias=bias[:ntt])
            lse_t = ops.ce_fwd(lt, labels_text, ct, loss)
        if Ri:
            li = ops.gemm_store(h[Rt:], wc[ntt:], bias=bias[ntt:])
            lse_i = ops.ce_fwd(li, labels_img, ci, loss)
        ctx.saved = (x3, mean, rstd, h, wc, lt, li, lse_t, lse_i, labels_text, labels_img, ln_w)
        ctx.cfg = (ntt, ct, ci, Rt, Ri)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        x3, mean, rstd, h, wc, lt, li, lse_t, lse_i, labels_text, labels_img, ln_w = ctx.saved
        ctx.saved = None
        ntt, ct, ci, Rt, Ri = ctx.cfg
        M, d = h.shape
        V = wc.shape[0]
        up = dloss.detach().reshape(1).float().contiguous()
        dh = torch.empty(M, d, device=h.device, dtype=h.dtype)
        dw = torch.zeros(V, d, device=h.device, dtype=torch.float32) if (not Rt or not Ri) else torch.empty(V, d, device=h.device, dtype=torch.float32)
        db = torch.zeros(V, device=h.device, dtype=torch.float32)
        if Rt:
            ops.ce_bwd_(lt, labels_text, ct, lse_t, up)
            ops.gemm_store(lt, wc[:ntt], a_mn=False, b_mn=True, out=dh[:Rt])
            ops.gemm_store(lt, h[:Rt], a_mn=True, b_mn=True, out=dw[:ntt])
            db[:ntt] = ops.colsum(lt)
        if Ri:
            ops.ce_bwd_(li, labels_img, ci, lse_i, up)
            ops.gemm_store(li, wc[ntt:], a_mn=False, b_mn=True, out=dh[Rt:])
            ops.gemm_store(li, h[Rt:], a_mn=True, b_mn=True, out=dw[ntt:])
            db[ntt:] = ops.colsum(li)
        dln_w = torch.zeros(d, device=h.device, dtype=torch.float32)
        dln_b = torch.zeros(d, device=h.device, dtype=torch.float32)
        dx = ops.ln_shift_bwd(dh, x3, mean, rstd, ln_w, None, 0, 1, do_ln=True, do_shift=False, dgamma=dln_w, dbeta=dln_b)
        return dx.view(M, d), dln_w, dln_b, dw, db, None, None, None, None, None, None
