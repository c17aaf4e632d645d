"""Import the UNMODIFIED reference (lucidrains/DALLE-pytorch @ /root/reference) in the dev container.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; nothing under tests -m gpu,
smoke() or bench.py may call this at run time.  It is used by oracle/make_golden.py (to generate the
committed fixtures under tests/golden/) and by the `not gpu` tests that pin oracle/dalle_oracle.py
against the live reference when it is present.

Mechanism: register a synthetic package object `dalle_pytorch` whose __path__ points
at /root/reference/dalle_pytorch so that its __init__.py (which pulls tokenizers / vae deps that are not
installed) is skipped, and put oracle/shims (restated rotary_embedding_torch, stubs for
axial_positional_embedding / omegaconf / taming) on sys.path.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get('DALLE_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'dalle_pytorch'))


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
