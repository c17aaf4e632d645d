"""Transformer stack with the reference's constructor, module tree and state-dict keys
(reference dalle_pytorch/transformer.py), executed as fused libdalle_b200 sub-layers.

The nesting LayerScale(PreNorm(CachedAs(PreShiftToken(CachedAs|NonCached(attn))))) is kept because it defines the
parameter names (`layers.layers.{i}.0.fn.fn.fn.fn.fn.to_qkv.weight` ...).  In training the
nest is never walked module by module: `LayerScale.plan()` resolves it once per call into a fused sub-layer
(functional.py) — LayerNorm + token shift -> GEMM(+rotary) -> attention -> GEMM(+bias, LayerScale, residual).
The module-by-module `forward`s remain for the inference-cache path (transformer.py:38-71, 138-153).
"""
from collections import deque
from collections.abc import Iterable
from functools import partial
from itertools import islice, cycle
from math import pi

import torch
from torch import nn
import torch.nn.functional as F

from . import ops, config
from .attention import (Attention, SparseAttention, SparseConvCausalAttention, SparseAxialCausalAttention, _AttentionBase,
                        rotary_tables, _key_mask_u8)
from .functional import (SublayerGeom, AttnSublayerFn, FFSublayerFn, LayerNormFn, attn_sublayer_forward, ff_sublayer_forward,
                         draw_dropout)
from .reversible import ReversibleSequence, SequentialSequence, _Sub


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def cast_tuple(val, depth=1):
    return val if isinstance(val, Iterable) else (val,) * depth


class DivideMax(nn.Module):
    """transformer.py:29-36"""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        maxes = x.amax(dim=self.dim, keepdim=True).detach()
        return x / maxes


class NonCached(nn.Module):
    """transformer.py:38-58: layers without cache support see the whole prefix again and return its suffix."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *, cache=None, cache_key=None, **kwargs):
        n = x.shape[-2]
        if exists(cache):
            if cache_key in cache:
                x = torch.cat([cache[cache_key], x], dim=-2)
            cache[cache_key] = x
        out = self.fn(x, **kwargs)
        return out[:, -n:]


class CachedAs(nn.Module):
    """transformer.py:60-71"""

    def __init__(self, cache_key, fn):
        super().__init__()
        self.cache_key = cache_key
        self.fn = fn

    def forward(self, x, *, cache=None, **kwargs):
        return self.fn(x, cache=cache, cache_key=self.cache_key, **kwargs)


class GEGLU(nn.Module):
    def forward(self, x):
        x, gates = x.chunk(2, dim=-1)
        return x * F.gelu(gates)


class FeedForward(nn.Module):
    """transformer.py:111-122.  `net` keeps the reference layout (0: Linear d->2*mult*d, 1: GEGLU, 2: Dropout,
    3: Linear mult*d->d); forward runs GEMM[bias+GEGLU] -> GEMM[bias] on the library kernels."""

    def __init__(self, dim, dropout=0., mult=4.):
        super().__init__()
        hidden = int(dim * mult)
        self.net = nn.Sequential(nn.Linear(dim, hidden * 2), GEGLU(), nn.Dropout(dropout), nn.Linear(hidden, dim))

    def dropout_active(self):
        return self.training and self.net[2].p > 0

    def geom(self, dtype, do_ln=False, do_shift=False, text_len=0, fmap=0):
        return SublayerGeom(dtype=dtype, text_len=text_len, fmap=fmap, do_ln=do_ln, do_shift=do_shift,
                            p_drop=self.net[2].p if self.dropout_active() else 0.0)

    def forward(self, x, cache=None, cache_key=None):
        dtype = config.compute_dtype()
        x = x.float()
        g = self.geom(dtype)
        w1, b1, w2, b2 = self.net[0].weight, self.net[0].bias, self.net[3].weight, self.net[3].bias
        if torch.is_grad_enabled() and (x.requires_grad or w1.requires_grad):
            return FFSublayerFn.apply(g, False, 1.0, x, None, None, None, w1, b1, w2, b2, None)
        out, _ = ff_sublayer_forward(g, x, None, None, None, w1, b1, w2, b2, None, 1.0, save=False,
                                     drop=draw_dropout(g, 'ff', x, w2.shape[1]))
        return out


class PreShiftToken(nn.Module):
    """transformer.py:126-200.  In training the shift is fused into the LayerNorm kernel (ln_shift_fwd); this
    module-level forward serves the inference cache (a deque of the last `image_size` tokens' shifted channels)."""

    def __init__(self, fn, image_size, seq_len):
        super().__init__()
        self.fn = fn
        self.image_size = image_size
        self.seq_len = seq_len
        self.img_seq_len = image_size ** 2
        self.text_len = seq_len - self.img_seq_len + 1

    def shift(self, x):
        """Token shift of a full prefix x [b,n,d] (n >= text_len), as a pure tensor op."""
        b, n, d = x.shape
        T, fm = self.text_len, self.image_size
        half, quarter = d // 2, d // 4
        out = x.clone()
        out[:, 1:T, :half] = x[:, :T - 1, :half]
        out[:, 0, :half] = 0
        n_img = n - T
        if n_img > 0:
            pos = torch.arange(n_img, device=x.device)
            row, col = pos // fm, pos % fm
            top = torch.zeros(b, n_img, quarter, device=x.device, dtype=x.dtype)
            has_top = row > 0
            top[:, has_top] = x[:, T:n][:, pos[has_top] - fm, :quarter]
            left = torch.zeros(b, n_img, half - quarter, device=x.device, dtype=x.dtype)
            has_left = col > 0
            left[:, has_left] = x[:, T:n][:, pos[has_left] - 1, quarter:half]
            out[:, T:, :quarter] = top
            out[:, T:, quarter:half] = left
        return out

    def forward(self, x, cache=None, cache_key=None, **kwargs):
        seq_len, image_size, text_len = self.seq_len, self.image_size, self.text_len
        if exists(cache) and cache_key in cache and cache.get('pos_t') is not None:
            # device-indexed form of the branch below (decode.GraphedDecoder): the position is a device tensor, the deque a ring
            ring = cache[cache_key]
            slot, prev, not_first = cache['shift_idx']
            x_top, x_left, *x_pass = x[:, -1].chunk(4, dim=-1)
            top_old = ring.top.index_select(0, slot)[0]                 # the token one row up (fmap steps ago)
            left_prev = ring.left.index_select(0, prev)[0]              # the previous token
            ring.top.index_copy_(0, slot, x_top[None])
            ring.left.index_copy_(0, slot, x_left[None])
            left_prev = torch.where(not_first, left_prev, torch.zeros_like(left_prev))
            x = torch.cat((top_old, left_prev, *x_pass), dim=-1)
            return self.fn(x[:, None], cache=cache, **kwargs)

        if exists(cache) and cache_key in cache:
            offset = cache['offset']
            assert offset >= text_len, "cached inference for text is not supported"
            q = cache[cache_key]
            assert isinstance(q, deque) and len(q) == image_size
            x_top, x_left, *x_pass = x[:, -1].chunk(4, dim=-1)
            q.append((x_top, x_left))
            x_top = q.popleft()[0]
            x_left = q[-2][1]
            if (offset - text_len) % image_size == 0:
                x_left = torch.zeros_like(x_left)
            x = torch.cat((x_top, x_left, *x_pass), dim=-1)
            return self.fn(x[:, None], cache=cache, **kwargs)

        n = x.shape[1]
        if n < text_len:
            return self.fn(x, **kwargs)
        x = self.shift(x)
        if exists(cache):
            dummy_top, dummy_left, *_ = x[:, -1].chunk(4, dim=-1)
            dummy_top, dummy_left = torch.zeros_like(dummy_top), torch.zeros_like(dummy_left)
            q = deque()
            x_img = x[:, text_len:][:, -image_size:]
            for _ in range(image_size - x_img.shape[1]):
                q.append((dummy_top, dummy_left))
            for i in range(x_img.shape[1]):
                q.append(x_img[:, i].chunk(4, dim=-1)[:2])
            cache[cache_key] = q
        return self.fn(x, cache=cache, **kwargs)


class PreNorm(nn.Module):
    """transformer.py:92-102"""

    def __init__(self, dim, fn, sandwich=False):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.norm_out = nn.LayerNorm(dim) if sandwich else nn.Identity()
        self.sandwich = sandwich
        self.fn = fn

    def _ln(self, ln, x):
        return LayerNormFn.apply(x.float(), ln.weight, ln.bias, ln.eps)

    def forward(self, x, **kwargs):
        x = self._ln(self.norm, x)
        x = self.fn(x, **kwargs)
        return self._ln(self.norm_out, x) if self.sandwich else x


class LayerScale(nn.Module):
    """transformer.py:74-88 (https://arxiv.org/abs/2103.17239)"""

    def __init__(self, dim, depth, fn):
        super().__init__()
        if depth <= 18:
            init_eps = 0.1
        elif depth > 18 and depth <= 24:
            init_eps = 1e-5
        else:
            init_eps = 1e-6
        scale = torch.zeros(1, 1, dim).fill_(init_eps)
        self.scale = nn.Parameter(scale)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(x, **kwargs) * self.scale

    # ---- fused execution -----------------------------------------------------------------------------
    def plan(self, x, cache=None, mask=None, rotary_pos_emb=None, **kwargs):
        """Resolve LayerScale(PreNorm([CachedAs(PreShiftToken(]CachedAs|NonCached(layer))) into one fused sub-layer, or
        None when the configuration needs the module-by-module path (inference cache, sandwich norm).  Active dropout is part of
        the fused sub-layer (geom.p_drop)."""
        if exists(cache) or kwargs:
            return None
        pre = self.fn
        if not isinstance(pre, PreNorm) or pre.sandwich:
            return None
        inner = pre.fn
        do_shift, text_len, fmap = False, 0, 0
        if isinstance(inner, CachedAs) and isinstance(inner.fn, PreShiftToken):
            ps = inner.fn
            do_shift, text_len, fmap = True, ps.text_len, ps.image_size
            inner = ps.fn
        if isinstance(inner, (CachedAs, NonCached)):
            inner = inner.fn
        dtype = config.compute_dtype()
        n = x.shape[1]
        p = _Sub()
        p.cos_t = p.sin_t = p.key_mask = None
        if isinstance(inner, _AttentionBase):
            p.kind = 'attn'
            p.geom = inner.geom(dtype, do_ln=True, do_shift=do_shift, text_len=text_len, fmap=fmap, n=n)
            p.geom.eps = pre.norm.eps
            p.params = dict(ln_w=pre.norm.weight, ln_b=pre.norm.bias, w_qkv=inner.to_qkv.weight, w_out=inner.to_out[0].weight,
                            b_out=inner.to_out[0].bias, scale=self.scale)
            p.cos_t, p.sin_t = rotary_tables(rotary_pos_emb, inner.dim_head)
            p.key_mask = _key_mask_u8(mask, n)
        elif isinstance(inner, FeedForward):
            p.kind = 'ff'
            p.geom = inner.geom(dtype, do_ln=True, do_shift=do_shift, text_len=text_len, fmap=fmap)
            p.geom.eps = pre.norm.eps
            p.params = dict(ln_w=pre.norm.weight, ln_b=pre.norm.bias, w1=inner.net[0].weight, b1=inner.net[0].bias,
                            w2=inner.net[3].weight, b2=inner.net[3].bias, scale=self.scale)
        else:
            return None
        return p

    def residual(self, x, **kwargs):
        """x + self(x) (reversible.py:139-140) as ONE fused sub-layer when possible."""
        p = self.plan(x, **kwargs)
        if p is None:
            return x + self(x, **kwargs)
        x = x.float()
        P = p.params
        grad = torch.is_grad_enabled()
        if p.kind == 'attn':
            if grad:
                return AttnSublayerFn.apply(p.geom, True, 1.0, p.cos_t, p.sin_t, p.key_mask, x, None, P['ln_w'], P['ln_b'],
                                            P['w_qkv'], P['w_out'], P['b_out'], P['scale'])
            out, _ = attn_sublayer_forward(p.geom, x, x, P['ln_w'], P['ln_b'], P['w_qkv'], P['w_out'], P['b_out'], P['scale'], 1.0,
                                           p.cos_t, p.sin_t, p.key_mask, save=False, drop=draw_dropout(p.geom, 'attn', x))
            return out
        if grad:
            return FFSublayerFn.apply(p.geom, True, 1.0, x, None, P['ln_w'], P['ln_b'], P['w1'], P['b1'], P['w2'], P['b2'], P['scale'])
        out, _ = ff_sublayer_forward(p.geom, x, x, P['ln_w'], P['ln_b'], P['w1'], P['b1'], P['w2'], P['b2'], P['scale'], 1.0, save=False,
                                     drop=draw_dropout(p.geom, 'ff', x, P['w2'].shape[1]))
        return out


def build_rotary_angle_table(text_len, image_fmap_size, dim_head):
    """The `pos_emb` buffer of transformer.py:304-328: angles [1, text_len + fmap^2, 6*(rot_dim//2)], per token
    [language(20) | axial row(20) | axial col(20)] with every frequency repeated on two adjacent dims.
    Frequencies follow rotary_embedding_torch.RotaryEmbedding ('lang': theta^(-2i/rot_dim); 'pixel': linspace(1, 5)*pi)."""
    rot_dim = dim_head // 3
    nf = rot_dim // 2
    lang = 1. / (10000 ** (torch.arange(0, rot_dim, 2)[:nf].float() / rot_dim))
    pixel = torch.linspace(1., 10 / 2, nf) * pi
    n_img = image_fmap_size ** 2

    def angles(pos, freqs):
        return (pos.float()[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)

    text_part = angles(torch.arange(text_len), lang)
    img_as_text = angles(torch.full((n_img,), 8192), lang)              # image tokens sit far away from the text
    lin = torch.linspace(-1, 1, steps=image_fmap_size)
    axis = angles(lin, pixel)                                           # [fmap, 2*nf]
    img_axial = torch.cat((axis[:, None, :].expand(-1, image_fmap_size, -1), axis[None, :, :].expand(image_fmap_size, -1, -1)),
                          dim=-1).reshape(n_img, -1)
    text_axial = angles(torch.full((text_len,), -10.), pixel)           # text sits at -10 on both image axes
    text_axial = torch.cat((text_axial, text_axial), dim=-1)
    table = torch.cat((torch.cat((text_part, img_as_text), dim=0), torch.cat((text_axial, img_axial), dim=0)), dim=-1)
    return table[None]


class Transformer(nn.Module):
    def __init__(self, *, dim, depth, seq_len, reversible=False, causal=True, heads=8, dim_head=64, ff_mult=4, attn_dropout=0.,
                 ff_dropout=0., attn_types=None, image_fmap_size=None, sparse_attn=False, stable=False, sandwich_norm=False,
                 shift_tokens=False, rotary_emb=True, shared_attn_ids=None, shared_ff_ids=None, optimize_for_inference=False):
        super().__init__()
        layers = nn.ModuleList([])
        sparse_layer = cast_tuple(sparse_attn, depth)
        self.seq_len = seq_len
        self.image_fmap_size = image_fmap_size

        attn_types = cast_tuple(default(attn_types, ('full',)))
        attn_type_layer = islice(cycle(attn_types), depth)
        shared_attn_ids = cycle(default(shared_attn_ids, range(depth)))
        shared_ff_ids = cycle(default(shared_ff_ids, range(depth)))
        shared_attn_layers, shared_ff_layers = {}, {}

        for ind, sparse_attn, attn_type, attn_id, ff_id in zip(range(depth), sparse_layer, attn_type_layer, shared_attn_ids,
                                                                shared_ff_ids):
            if attn_type == 'full':
                attn_class = partial(Attention, stable=stable)
            elif attn_type == 'sparse':
                attn_class = SparseAttention
            elif attn_type in ('axial_row', 'axial_col'):
                if optimize_for_inference:      # cache-friendly dense attention with a static mask (transformer.py:251-260)
                    attn_class = partial(Attention, stable=stable, static_mask=self._get_attention_mask(attn_type))
                else:
                    attn_class = partial(SparseAxialCausalAttention, seq_len=seq_len, axis=0 if attn_type == 'axial_row' else 1,
                                         image_size=image_fmap_size, stable=stable)
            elif attn_type == 'conv_like':
                attn_class = partial(SparseConvCausalAttention, seq_len=seq_len, image_size=image_fmap_size, stable=stable)
            else:
                raise ValueError(f'attention type "{attn_type}" is not valid')

            attn, reused_attn_type = shared_attn_layers.get(attn_id, (None, None))
            if not exists(attn):
                attn = attn_class(dim, causal=causal, seq_len=seq_len, heads=heads, dim_head=dim_head, dropout=attn_dropout)
                shared_attn_layers[attn_id] = (attn, attn_type)
            elif attn_type != reused_attn_type:
                raise ValueError('attn_types do not match shared_attn_ids '
                                 f'(ind = {ind}, attn_type = "{attn_type}", reused_attn_type = "{reused_attn_type}")')

            ff = shared_ff_layers.get(ff_id)
            if not exists(ff):
                ff = FeedForward(dim, mult=ff_mult, dropout=ff_dropout)
                shared_ff_layers[ff_id] = ff

            if isinstance(attn, Attention):
                attn = CachedAs(f'attn_{ind}', attn)
            else:
                attn = NonCached(attn)

            if shift_tokens:
                attn = CachedAs(f'preshift_attn_{ind}', PreShiftToken(attn, image_size=image_fmap_size, seq_len=seq_len))
                ff = CachedAs(f'preshift_ff_{ind}', PreShiftToken(ff, image_size=image_fmap_size, seq_len=seq_len))

            layers.append(nn.ModuleList([
                LayerScale(dim, ind + 1, PreNorm(dim, attn, sandwich=sandwich_norm)),
                LayerScale(dim, ind + 1, PreNorm(dim, ff, sandwich=sandwich_norm)),
            ]))

        execute_type = ReversibleSequence if reversible else SequentialSequence
        route_attn = ((True, False),) * depth
        route_all = ((True, True),) * depth
        attn_route_map = {'mask': route_attn, 'rotary_pos_emb': route_attn, 'cache': route_all}
        self.layers = execute_type(layers, args_route=attn_route_map)

        pos_emb = None
        if rotary_emb:
            img_seq_len = image_fmap_size ** 2
            text_len = seq_len - img_seq_len + 1
            pos_emb = build_rotary_angle_table(text_len, image_fmap_size, dim_head)
        self.register_buffer('pos_emb', pos_emb)

    def forward(self, x, **kwargs):
        return self.layers(x, rotary_pos_emb=self.pos_emb, **kwargs)

    def _get_attention_mask(self, attn_type):
        """transformer.py:333-350: static masks that reproduce the axial patterns for the cached dense attention."""
        fm = self.image_fmap_size
        img_seq_len = fm ** 2
        text_len = self.seq_len + 1 - img_seq_len
        static_mask = torch.zeros(self.seq_len, self.seq_len, dtype=torch.bool)
        static_mask[:, :text_len] = True
        pos = torch.arange(self.seq_len - text_len)
        row, col = pos // fm, pos % fm
        if attn_type == 'axial_row':
            static_mask[text_len:, text_len:] = row[:, None] == row[None, :]
        elif attn_type == 'axial_col':
            static_mask[text_len:, text_len:] = col[:, None] == col[None, :]
        else:
            raise ValueError(f'attention type "{attn_type}" can\'t be simulated with a static mask')
        return static_mask
