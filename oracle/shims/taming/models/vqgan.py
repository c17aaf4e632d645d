"""Import stub (vae.py:16). TEST INFRASTRUCTURE ONLY."""
class VQModel:  # pragma: no cover
    pass
class GumbelVQ:  # pragma: no cover
    pass
