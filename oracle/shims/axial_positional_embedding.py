"""Stub of the un-vendored `axial_positional_embedding` package (dalle_pytorch.py:7), used by the
reference only when rotary_emb=False (dalle_pytorch.py:389).  Restated: one learned [1,s,1,..,d]
parameter per axis, broadcast-summed over the axial grid, flattened and sliced to the input length.
TEST INFRASTRUCTURE ONLY (lets the reference import in the dev container)."""
import torch
from torch import nn
from functools import reduce
from operator import mul


class AxialPositionalEmbedding(nn.Module):
    def __init__(self, dim, axial_shape, axial_dims=None):
        super().__init__()
        self.dim, self.shape = dim, axial_shape
        self.max_seq_len = reduce(mul, axial_shape, 1)
        self.summed = axial_dims is None
        axial_dims = ((dim,) * len(axial_shape)) if self.summed else axial_dims
        self.weights = nn.ParameterList()
        for ind, (shape, axial_dim) in enumerate(zip(self.shape, axial_dims)):
            ax_shape = [1] * len(self.shape)
            ax_shape[ind] = shape
            ax_shape = (1, *ax_shape, axial_dim)
            self.weights.append(nn.Parameter(torch.zeros(ax_shape).normal_(0, 1)))

    def forward(self, x):
        b, t, e = x.shape
        embs = []
        for ax_emb in self.weights:
            axial_dim = ax_emb.shape[-1]
            expand_shape = (b, *self.shape, axial_dim)
            emb = ax_emb.expand(expand_shape).reshape(b, self.max_seq_len, axial_dim)
            embs.append(emb)
        pos_emb = sum(embs) if self.summed else torch.cat(embs, dim=-1)
        return pos_emb[:, :t].to(x)
