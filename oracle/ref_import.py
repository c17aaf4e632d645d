"""Import the UNMODIFIED reference (lucidrains/DALLE-pytorch) from /root/reference (dev container) or from the
offline install under baseline/_ref (`pip install --no-deps --target baseline/_ref`, which travels to the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product package.  Users: oracle/make_golden.py (generates
the committed fixtures under tests/golden/), the `not gpu` tests that pin oracle/dalle_oracle.py against the live
reference, the drop-in test of `patch_dalle_pytorch()`, and bench.py's reference legs (`--impl reference`, `cpu_baseline`,
`gpu_eager_baseline`), which time the reference's own `DALLE(...)` stock code path.

Mechanism: register a synthetic package object `dalle_pytorch` whose __path__ points
at /root/reference/dalle_pytorch so that its __init__.py (which pulls tokenizers / vae deps that are not
installed) is skipped, and put oracle/shims (restated rotary_embedding_torch, stubs for
axial_positional_embedding / omegaconf / taming) on sys.path.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, 'shims')


def _find_root():
    cands = [os.environ.get('DALLE_REFERENCE_ROOT'), '/root/reference', os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, 'dalle_pytorch', 'dalle_pytorch.py')):
            return c
    return cands[1]


REF_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'dalle_pytorch', 'dalle_pytorch.py'))


def import_reference():
    """Returns a namespace with the reference's DALLE, DiscreteVAE, Transformer, Attention, ... classes."""
    if not reference_available():
        raise RuntimeError(f'reference not found at {REF_ROOT}')
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if 'dalle_pytorch' not in sys.modules or not getattr(sys.modules['dalle_pytorch'], '_is_ref_shim', False):
        pkg = types.ModuleType('dalle_pytorch')
        pkg.__path__ = [os.path.join(REF_ROOT, 'dalle_pytorch')]
        pkg._is_ref_shim = True
        sys.modules['dalle_pytorch'] = pkg
    ns = types.SimpleNamespace()
    ns.attention = importlib.import_module('dalle_pytorch.attention')
    ns.transformer = importlib.import_module('dalle_pytorch.transformer')
    ns.reversible = importlib.import_module('dalle_pytorch.reversible')
    ns.dalle = importlib.import_module('dalle_pytorch.dalle_pytorch')
    ns.DALLE = ns.dalle.DALLE
    ns.DiscreteVAE = ns.dalle.DiscreteVAE
    ns.Transformer = ns.transformer.Transformer
    ns.Attention = ns.attention.Attention
    ns.SparseAxialCausalAttention = ns.attention.SparseAxialCausalAttention
    ns.SparseConvCausalAttention = ns.attention.SparseConvCausalAttention
    return ns
