"""Drop-in installer: rebinds the hot-path class names of an imported `dalle_pytorch` (the reference) to the
libdalle_b200-backed classes of this package, so that an UNMODIFIED `dalle_pytorch.DALLE(...)` /
`dalle_pytorch.transformer.Transformer(...)` builds the fused sm_100a blocks.

The reference's `Transformer.__init__` looks the names up in its module globals at call time
(reference transformer.py:245-292: `Attention`, `SparseAttention`, `SparseAxialCausalAttention`,
`SparseConvCausalAttention`, `FeedForward`, `LayerScale`, `PreNorm`, `PreShiftToken`, `CachedAs`, `NonCached`,
`SequentialSequence`, `ReversibleSequence`), so rebinding them is all that is needed; parameter names and shapes are
identical, so reference checkpoints load unchanged.

    import dalle_pytorch, dalle_pytorch_b200
    undo = dalle_pytorch_b200.patch_dalle_pytorch()      # models built from here on run on libdalle_b200
    dalle = dalle_pytorch.DALLE(dim=1024, vae=vae, ...).cuda()
    undo()                                               # restores the reference classes (already-built models keep theirs)
"""
import importlib
import sys

_ATTN_NAMES = ('Attention', 'SparseAxialCausalAttention', 'SparseConvCausalAttention', 'SparseAttention')
_BLOCK_NAMES = ('FeedForward', 'GEGLU', 'LayerScale', 'PreNorm', 'PreShiftToken', 'CachedAs', 'NonCached')
_EXEC_NAMES = ('SequentialSequence', 'ReversibleSequence')


def patch_dalle_pytorch(package='dalle_pytorch'):
    """Rebinds the names listed above in `<package>.attention`, `<package>.transformer` and `<package>.reversible`.
    Returns a callable that restores the previous bindings.  Raises ImportError if the reference is not importable and
    AttributeError if it does not define one of the names (an incompatible version must not be patched half-way)."""
    import dalle_pytorch_b200 as b200
    mods = {m: (sys.modules.get(f'{package}.{m}') or importlib.import_module(f'{package}.{m}'))
            for m in ('attention', 'transformer', 'reversible')}
    plan = []
    for name in _ATTN_NAMES:
        plan.append((mods['attention'], name))
        plan.append((mods['transformer'], name))
    for name in _BLOCK_NAMES:
        plan.append((mods['transformer'], name))
    for name in _EXEC_NAMES:
        plan.append((mods['reversible'], name))
        plan.append((mods['transformer'], name))
    for mod, name in plan:                               # validate everything before touching anything
        if not hasattr(mod, name):
            raise AttributeError(f'{mod.__name__} has no attribute {name!r}: not the dalle_pytorch layout this package mirrors')
        if not hasattr(b200, name):
            raise AttributeError(f'dalle_pytorch_b200 has no replacement for {name!r}')
    saved = [(mod, name, getattr(mod, name)) for mod, name in plan]
    for mod, name in plan:
        setattr(mod, name, getattr(b200, name))

    def undo():
        for mod, name, old in saved:
            setattr(mod, name, old)

    return undo
