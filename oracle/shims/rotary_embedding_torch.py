"""Restatement of the two entry points of the un-vendored PyPI package `rotary-embedding-torch`
that the reference calls (attention.py:9,35 `apply_rotary_emb`; transformer.py:14,310-318
`RotaryEmbedding`, `broadcat`).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified reference
under /root/reference can be imported in the dev container to generate golden vectors.  The package is
unpinned in the reference's setup.py:29, so this follows its published algorithm:
interleaved-pair rotation, 'lang' freqs 1/theta^(2i/dim), 'pixel' freqs linspace(1, max_freq/2)*pi.
"""
from math import pi
import torch
from torch import nn
from einops import rearrange, repeat


def broadcat(tensors, dim=-1):
    num_tensors = len(tensors)
    shape_lens = set(len(t.shape) for t in tensors)
    assert len(shape_lens) == 1, 'tensors must all have the same number of dimensions'
    shape_len = list(shape_lens)[0]
    dim = (dim + shape_len) if dim < 0 else dim
    dims = list(zip(*(list(t.shape) for t in tensors)))
    expandable = [(i, val) for i, val in enumerate(dims) if i != dim]
    assert all(len(set(v)) <= 2 for _, v in expandable), 'invalid dimensions for broadcastable concat'
    max_dims = [(i, max(v)) for i, v in expandable]
    expanded = [(i, (v,) * num_tensors) for i, v in max_dims]
    expanded.insert(dim, (dim, dims[dim]))
    shapes = list(zip(*(v for _, v in expanded)))
    tensors = [t.expand(*s) for t, s in zip(tensors, shapes)]
    return torch.cat(tensors, dim=dim)


def rotate_half(x):
    x = rearrange(x, '... (d r) -> ... d r', r=2)
    x1, x2 = x.unbind(dim=-1)
    x = torch.stack((-x2, x1), dim=-1)
    return rearrange(x, '... d r -> ... (d r)')


def apply_rotary_emb(freqs, t, start_index=0):
    freqs = freqs.to(t)
    rot_dim = freqs.shape[-1]
    end_index = start_index + rot_dim
    assert rot_dim <= t.shape[-1]
    t_left, t_mid, t_right = t[..., :start_index], t[..., start_index:end_index], t[..., end_index:]
    t_mid = (t_mid * freqs.cos()) + (rotate_half(t_mid) * freqs.sin())
    return torch.cat((t_left, t_mid, t_right), dim=-1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, custom_freqs=None, freqs_for='lang', theta=10000, max_freq=10, num_freqs=1,
                 learned_freq=False):
        super().__init__()
        if custom_freqs is not None:
            freqs = custom_freqs
        elif freqs_for == 'lang':
            freqs = 1. / (theta ** (torch.arange(0, dim, 2)[:(dim // 2)].float() / dim))
        elif freqs_for == 'pixel':
            freqs = torch.linspace(1., max_freq / 2, dim // 2) * pi
        elif freqs_for == 'constant':
            freqs = torch.ones(num_freqs).float()
        else:
            raise ValueError(f'unknown modality {freqs_for}')
        self.cache = dict()
        if learned_freq:
            self.freqs = nn.Parameter(freqs)
        else:
            self.register_buffer('freqs', freqs)

    def forward(self, t, cache_key=None):
        freqs = self.freqs
        freqs = torch.einsum('..., f -> ... f', t.type(freqs.dtype), freqs)
        freqs = repeat(freqs, '... n -> ... (n r)', r=2)
        return freqs
