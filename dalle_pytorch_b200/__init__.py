"""dalle_pytorch_b200 — the DALL-E transformer block hot path on B200 (sm_100a).

Same module API as lucidrains/DALLE-pytorch for the path (DALLE, Transformer, Attention,
SparseAxialCausalAttention, SparseConvCausalAttention, SparseAttention, FeedForward, ...), executed by the
hand-written CUDA kernels of libdalle_b200.so through a C ABI (include/dalle_b200.h).
"""
from .config import set_compute_dtype, compute_dtype, compute_dtype_ctx, set_fp32_gemm, fp32_gemm, fp32_gemm_ctx
from .attention import Attention, SparseAxialCausalAttention, SparseConvCausalAttention, SparseAttention
from .transformer import (Transformer, FeedForward, GEGLU, LayerScale, PreNorm, PreShiftToken, CachedAs, NonCached, DivideMax)
from .reversible import SequentialSequence, ReversibleSequence
from .dalle import DALLE, TokenVAE
from . import ops, functional
from .optim import FusedAdam
from .patch import patch_dalle_pytorch
from .graph import GraphedStep

__version__ = '0.1.0'
