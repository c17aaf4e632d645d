"""Precision mode of the hot path.

    fp32  (default)  every tensor fp32, FFMA GEMMs / fp32 attention: the PARITY mode (rtol 1e-3 / atol 1e-5
                     against the reference's CPU path, BASELINE.json north_star)
    bf16             bf16 activations + weights copies, fp32 accumulation, fp32 residual stream / LayerNorm /
                     softmax statistics / gradients: the SPEED mode (tcgen05 GEMMs).  Selected explicitly with
                     set_compute_dtype(torch.bfloat16) / `with compute_dtype_ctx(torch.bfloat16)`, or implicitly
                     inside `torch.autocast('cuda', dtype=torch.bfloat16)`.
"""
import contextlib

import torch

_dtype = torch.float32


def set_compute_dtype(dtype):
    global _dtype
    if isinstance(dtype, str):
        dtype = {'fp32': torch.float32, 'float32': torch.float32, 'bf16': torch.bfloat16, 'bfloat16': torch.bfloat16}[dtype]
    assert dtype in (torch.float32, torch.bfloat16)
    _dtype = dtype


def compute_dtype():
    if torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16:
        return torch.bfloat16
    return _dtype


@contextlib.contextmanager
def compute_dtype_ctx(dtype):
    global _dtype
    old = _dtype
    set_compute_dtype(dtype)
    try:
        yield
    finally:
        _dtype = old
