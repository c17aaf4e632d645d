"""Import stub (vae.py:15). TEST INFRASTRUCTURE ONLY."""
class OmegaConf:
    @staticmethod
    def load(*a, **k):
        raise RuntimeError('omegaconf is not available in this environment')
