"""Precision mode of the hot path.

    fp32  (default)  every tensor fp32: the PARITY mode (rtol 1e-3 / atol 1e-5 against the reference's CPU path,
                     BASELINE.json north_star).  Its GEMMs run on the SAME tcgen05 kernel as the speed mode, as a
                     "bf16x6" product (each fp32 operand split into three bf16 pieces, six partial products
                     accumulated in fp32 in tensor memory, ops.py::_gemm_x6); `set_fp32_gemm('simt')` /
                     DALLE_B200_FP32_GEMM=simt selects the CUDA-core FFMA kernel instead (cross-check).  Attention
                     uses the fp32-arithmetic kernels.
    bf16             bf16 activations + weights copies, fp32 accumulation, fp32 residual stream / LayerNorm /
                     softmax statistics / gradients: the SPEED mode (tcgen05 GEMMs).  Selected explicitly with
                     set_compute_dtype(torch.bfloat16) / `with compute_dtype_ctx(torch.bfloat16)`, or implicitly
                     inside `torch.autocast('cuda', dtype=torch.bfloat16)`.
"""
import contextlib
import os

import torch

_dtype = torch.float32
_fp32_gemm = os.environ.get('DALLE_B200_FP32_GEMM', 'bf16x6')


def set_fp32_gemm(kind):
    """'bf16x6' (tcgen05, default) or 'simt' (CUDA-core fp32): how GEMMs of the fp32 parity mode are evaluated."""
    global _fp32_gemm
    assert kind in ('bf16x6', 'simt')
    _fp32_gemm = kind


def fp32_gemm():
    return _fp32_gemm


@contextlib.contextmanager
def fp32_gemm_ctx(kind):
    global _fp32_gemm
    old = _fp32_gemm
    set_fp32_gemm(kind)
    try:
        yield
    finally:
        _fp32_gemm = old


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
