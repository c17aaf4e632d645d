"""Attention modules with the reference's constructor / forward signatures and parameter names
(reference dalle_pytorch/attention.py), executed by the libdalle_b200 kernels.

    Attention                    attention.py:39-99    dense (causal / non-causal / static mask / KV cache)
    SparseAxialCausalAttention   attention.py:225-335  axial row / column
    SparseConvCausalAttention    attention.py:103-221  conv-like window
    SparseAttention(Attention)   attention.py:339-398  DeepSpeed block-sparse layout, evaluated as a static mask

All of them are one kernel family: q/k/v projection GEMM (rotary + scale + head split fused in the epilogue), a
flash-style attention kernel whose sparsity pattern is a predicate evaluated in registers, and the output
projection GEMM.  No [n,n] tensor is ever materialised.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from . import ops, config
from .functional import SublayerGeom, AttnSublayerFn, attn_sublayer_forward, _w, draw_dropout
from ._lib import ATTN_FULL, ATTN_AXIAL_ROW, ATTN_AXIAL_COL, ATTN_CONV_LIKE, ATTN_STATIC


def exists(val):
    return val is not None


_rot_cache = {}


def rotary_tables(rotary_pos_emb, dim_head):
    """Angle table [..., n, R] (transformer.py:304-328: each frequency repeated on adjacent pairs) ->
    (cos, sin) fp32 [n, dim_head/2] with (1, 0) on the pass-through pairs.  cos/sin are evaluated in float64
    on the stored fp32 angles and rounded once (image tokens sit at angle 8192*f, SURVEY.md §7 hard parts)."""
    if rotary_pos_emb is None:
        return None, None
    key = (rotary_pos_emb.data_ptr(), rotary_pos_emb._version, tuple(rotary_pos_emb.shape), str(rotary_pos_emb.device), dim_head)
    hit = _rot_cache.get(key)
    if hit is not None:
        return hit
    ang = rotary_pos_emb.detach().reshape(-1, rotary_pos_emb.shape[-1])
    R = ang.shape[-1]
    if R % 2 or R > dim_head or not torch.equal(ang[:, 0::2], ang[:, 1::2]):
        raise NotImplementedError('rotary table must hold pair-repeated angles with rot_dim <= dim_head')
    a = ang[:, 0::2].double()
    n = a.shape[0]
    cos = torch.ones(n, dim_head // 2, dtype=torch.float64, device=ang.device)
    sin = torch.zeros(n, dim_head // 2, dtype=torch.float64, device=ang.device)
    cos[:, :R // 2] = a.cos()
    sin[:, :R // 2] = a.sin()
    out = (cos.float().contiguous(), sin.float().contiguous())
    if len(_rot_cache) > 64:
        _rot_cache.clear()
    _rot_cache[key] = out
    return out


def _key_mask_u8(mask, n_k):
    if mask is None:
        return None
    m = mask[:, :n_k]
    if m.shape[1] < n_k:       # keys beyond the supplied mask are kept (sparse classes only mask text keys, attention.py:260)
        m = F.pad(m, (0, n_k - m.shape[1]), value=True)
    return m.to(torch.uint8).contiguous()


class _AttentionBase(nn.Module):
    """Parameters + the pieces shared by every pattern."""

    def _init_params(self, dim, heads, dim_head, dropout):
        inner_dim = dim_head * heads
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))

    def attn_spec(self, n):
        raise NotImplementedError

    def dropout_active(self):
        return self.training and self.to_out[1].p > 0

    def geom(self, dtype, do_ln=False, do_shift=False, text_len=0, fmap=0, n=None):
        return SublayerGeom(dtype=dtype, text_len=text_len, fmap=fmap, do_ln=do_ln, do_shift=do_shift, heads=self.heads,
                            dim_head=self.dim_head, attn_spec=self.attn_spec(n), q_scale=self.scale,
                            p_drop=self.to_out[1].p if self.dropout_active() else 0.0)

    def _plain_forward(self, x, mask, rotary_pos_emb):
        """to_out(attend(to_qkv(x))) without norm / shift / LayerScale / residual."""
        dtype = config.compute_dtype()
        b, n, _ = x.shape
        x = x.float()
        cos_t, sin_t = rotary_tables(rotary_pos_emb, self.dim_head)
        g = self.geom(dtype, n=n)
        km = _key_mask_u8(mask, n)
        w_qkv, w_out, b_out = self.to_qkv.weight, self.to_out[0].weight, self.to_out[0].bias
        # (the Dropout of to_out, attention.py:53-56, is applied inside the fused sub-layer when it is active: g.p_drop)
        if torch.is_grad_enabled() and (x.requires_grad or w_qkv.requires_grad):
            return AttnSublayerFn.apply(g, False, 1.0, cos_t, sin_t, km, x, None, None, None, w_qkv, w_out, b_out, None)
        out, _ = attn_sublayer_forward(g, x, None, None, None, w_qkv, w_out, b_out, None, 1.0, cos_t, sin_t, km, save=False,
                                       drop=draw_dropout(g, 'attn', x))
        return out


class Attention(_AttentionBase):
    def __init__(self, dim, seq_len, causal=True, heads=8, dim_head=64, dropout=0., stable=False, static_mask=None):
        super().__init__()
        self.seq_len = seq_len
        self.stable = stable
        self.causal = causal
        self.register_buffer('static_mask', static_mask, persistent=False)
        self._init_params(dim, heads, dim_head, dropout)
        self._static_u8 = None

    def _static(self):
        if self.static_mask is None:
            return None
        sm = self._static_u8
        if sm is None or sm.device != self.static_mask.device or sm.shape != self.static_mask.shape:
            sm = self.static_mask.to(torch.uint8).contiguous()
            self._static_u8 = sm
        return sm

    def attn_spec(self, n=None):
        sm = self._static()
        if sm is not None:
            return ops.AttnSpec(ATTN_STATIC, causal=self.causal, stable=self.stable, static_mask=sm)
        return ops.AttnSpec(ATTN_FULL, causal=self.causal, stable=self.stable)

    def forward(self, x, mask=None, rotary_pos_emb=None, cache=None, cache_key=None):
        if not exists(cache):
            return self._plain_forward(x, mask, rotary_pos_emb)
        return self._cached_forward(x, mask, rotary_pos_emb, cache, cache_key)

    @torch.no_grad()
    def _cached_forward(self, x, mask, rotary_pos_emb, cache, cache_key):
        """KV-cache decoding (attention.py:61, 71-76, 84-92): queries are the last n positions of the cached keys.  The cache is
        IN PLACE: one [b, h, seq_len, dh] buffer pair per layer, allocated at the first call; every step writes its new rows behind
        the valid ones and the attention kernel is told how many rows are valid (db200_attn_fwd_params::kv_rows) -- the reference
        re-allocates and copies the whole cache with torch.cat at every token."""
        dtype = config.compute_dtype()
        b, n, d = x.shape
        if cache.get('pos_t') is not None:
            # device-indexed step (decode.GraphedDecoder, capturable in a CUDA graph): rotary row, cache write position and the
            # allowed keys are selected on the device; attention runs over the whole buffer with the key mask of this position
            assert n == 1 and mask is None, 'graph-replayed decoding feeds one token per step and no padding mask'
            cos_r, sin_r = cache['rot_row']
            ent = cache[cache_key]
            a, _, _ = ops.ln_shift_fwd(x.float().contiguous(), None, None, dtype, 0, 1, do_ln=False, do_shift=False)
            wq, wo = _w(self.to_qkv.weight, dtype), _w(self.to_out[0].weight, dtype)
            q, k, v = ops.gemm_qkv(a, wq, b, 1, self.heads, self.dim_head, cos_r, sin_r, self.scale, pos_offset=0)
            ent.k.index_copy_(2, cache['pos_t'], k)
            ent.v.index_copy_(2, cache['pos_t'], v)
            o, _ = ops.attn_fwd(ops.AttnSpec(ATTN_FULL, causal=False, stable=self.stable), q, ent.k, ent.v, cache['key_mask'][cache_key],
                                n_k=cache.get('n_k'))
            out, _ = ops.gemm_resid(o.view(b, -1), wo, self.to_out[0].bias.detach(), None, None, 1.0)
            return out.view(b, 1, d)
        offset = cache.get('offset', 0)
        cos_t, sin_t = rotary_tables(rotary_pos_emb, self.dim_head)
        a, _, _ = ops.ln_shift_fwd(x.float().contiguous(), None, None, dtype, 0, 1, do_ln=False, do_shift=False)
        wq, wo = _w(self.to_qkv.weight, dtype), _w(self.to_out[0].weight, dtype)        # cached bf16 copies in bf16 mode
        q, k, v = ops.gemm_qkv(a, wq, b, n, self.heads, self.dim_head, cos_t, sin_t, self.scale, pos_offset=offset)
        n_k = offset + n
        ent = cache.get(cache_key)
        cap = max(self.seq_len + 1, n_k)
        if (not isinstance(ent, _KVCache) or ent.k.shape[0] != b or ent.k.shape[2] < n_k or ent.k.dtype != k.dtype or ent.k.device != k.device
                or offset == 0):
            # (zero-filled once: the tensor-core kernels load whole 64-row tiles and multiply the rows behind n_k by exact zeros)
            ent = _KVCache(torch.zeros(b, self.heads, cap, self.dim_head, device=k.device, dtype=k.dtype),
                           torch.zeros(b, self.heads, cap, self.dim_head, device=k.device, dtype=k.dtype))
            cache[cache_key] = ent
        ent.k[:, :, offset:n_k].copy_(k)
        ent.v[:, :, offset:n_k].copy_(v)
        o, _ = ops.attn_fwd(self.attn_spec(n_k), q, ent.k, ent.v, _key_mask_u8(mask, n_k), n_k=n_k)
        out, _ = ops.gemm_resid(o.view(b * n, -1), wo, self.to_out[0].bias.detach(), None, None, 1.0)
        return out.view(b, n, d)


class _KVCache:
    """In-place key/value cache of one attention layer (the `cache[cache_key]` entry of attention.py:71-76)."""
    __slots__ = ('k', 'v')

    def __init__(self, k, v):
        self.k, self.v = k, v


class SparseAxialCausalAttention(_AttentionBase):
    def __init__(self, dim, seq_len, image_size=32, axis=0, heads=8, dim_head=64, dropout=0., stable=False, **kwargs):
        super().__init__()
        assert axis in {0, 1}, 'axis must be either 0 (along height) or 1 (along width)'
        self.axis = axis
        self.seq_len = seq_len
        self.image_size = image_size
        self.stable = stable
        self._init_params(dim, heads, dim_head, dropout)

    @property
    def text_len(self):
        return self.seq_len + 1 - self.image_size ** 2          # attention.py:251

    def attn_spec(self, n=None):
        return ops.AttnSpec(ATTN_AXIAL_ROW if self.axis == 0 else ATTN_AXIAL_COL, causal=True, stable=self.stable,
                            text_len=self.text_len, fmap=self.image_size)

    def forward(self, x, mask=None, rotary_pos_emb=None):
        # the reference zero-pads to seq_len+1 and slices the result back (attention.py:255-258, 335); padded keys are
        # never visible to a real query, so the predicate kernel simply runs on the n real tokens.
        return self._plain_forward(x, mask, rotary_pos_emb)


class SparseConvCausalAttention(_AttentionBase):
    def __init__(self, dim, seq_len, image_size=32, kernel_size=5, dilation=1, heads=8, dim_head=64, dropout=0., stable=False,
                 **kwargs):
        super().__init__()
        assert kernel_size % 2 == 1, 'kernel size must be odd'
        self.seq_len = seq_len
        self.image_size = image_size
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.stable = stable
        self._init_params(dim, heads, dim_head, dropout)

    @property
    def text_len(self):
        return self.seq_len + 1 - self.image_size ** 2

    def attn_spec(self, n=None):
        return ops.AttnSpec(ATTN_CONV_LIKE, causal=True, stable=self.stable, text_len=self.text_len, fmap=self.image_size,
                            kernel_size=self.kernel_size, dilation=self.dilation)

    def forward(self, x, mask=None, rotary_pos_emb=None):
        return self._plain_forward(x, mask, rotary_pos_emb)


class SparseAttention(Attention):
    """Block-sparse attention (attention.py:339-398).  The reference delegates to DeepSpeed's
    `SparseSelfAttention(VariableSparsityConfig(...))` + Triton kernels, which are un-vendored, unpinned and draw their random
    blocks from Python's RNG ("parity unpinned").  This class pins the LAYOUT FAMILY with a written spec and
    evaluates it with the library's attention kernels (static-mask pattern); no DeepSpeed import.

    Layout spec (block = `block_size` = 16 tokens, nb = ceil(seq_len / block) block rows / columns; the same layout for every
    head, as VariableSparsityConfig's default `different_layout_per_head=False`):
      * global: every block row attends the first ceil(text_seq_len / block) block columns  (`global_block_indices`,
        attention.py:355) -- the text tokens;
      * local: block row r attends block column r (VariableSparsityConfig's default `local_window_blocks=[4]` widens this to the
        window of 4 blocks containing r: columns 4*(r//4) .. r);
      * random: block row r > 0 additionally attends `num_random_blocks` (default seq_len // block // 4, attention.py:353) block
        columns drawn uniformly from [0, r] (unidirectional) -- from a generator seeded with `layout_seed` (default 0) so that the
        layout is reproducible, where DeepSpeed uses the unseeded `random` module;
      * causal ('unidirectional', attention.py:361): inside the allowed blocks token i attends token j <= i.
    `block_layout()` returns the [nb, nb] boolean layout; the token-level mask is its Kronecker expansion cropped to seq_len."""

    def __init__(self, *args, block_size=16, text_seq_len=256, num_random_blocks=None, layout_seed=0, **kwargs):
        super().__init__(*args, **kwargs)
        self.block_size = block_size
        num_random_blocks = num_random_blocks if exists(num_random_blocks) else self.seq_len // block_size // 4
        self.num_random_blocks = num_random_blocks
        self.num_global_blocks = math.ceil(text_seq_len / block_size)
        nb = math.ceil(self.seq_len / block_size)
        gen = torch.Generator().manual_seed(layout_seed)
        layout = torch.zeros(nb, nb, dtype=torch.bool)
        layout[:, :self.num_global_blocks] = True                  # global attention to the text blocks
        for r in range(nb):
            layout[r, 4 * (r // 4):r + 1] = True                   # local window of 4 blocks, up to the row's own block
            if num_random_blocks > 0 and r > 0:
                idx = torch.randint(0, r + 1, (num_random_blocks,), generator=gen)
                layout[r, idx] = True                              # random blocks (unidirectional: at or before the row)
        if self.causal:
            layout &= torch.ones(nb, nb, dtype=torch.bool).tril()
        self._layout = layout
        mask = layout.repeat_interleave(block_size, 0).repeat_interleave(block_size, 1)[:self.seq_len, :self.seq_len]
        self.register_buffer('static_mask', mask, persistent=False)
        self._static_u8 = None

    def block_layout(self):
        return self._layout.clone()

    def forward(self, x, mask=None, rotary_pos_emb=None):
        return self._plain_forward(x, mask, rotary_pos_emb)
