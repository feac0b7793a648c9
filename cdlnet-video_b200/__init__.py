"""cdlnet-video_b200 — B200-native (sm_100a) implementation of the CDLNet K-iteration ISTA forward pass.

Importable as `cdlnet_video_b200` (see the alias package next to this directory) or, with this
directory first on sys.path, as the reference's own `model` package (`from model.net import CDLNet`).
"""
from . import _lib                        # noqa: F401
from .plan import Plan                    # noqa: F401
from .model.net import CDLNet, CDLNetVideo, GDLNet, CDLNet_CSR, CDLNet_CSRf2, ST, prox_CSR, prox_CSR_f2      # noqa: F401

__all__ = ["CDLNet", "CDLNetVideo", "GDLNet", "CDLNet_CSR", "CDLNet_CSRf2", "ST", "prox_CSR", "prox_CSR_f2", "Plan", "build", "load_library"]


def build(force=False, verbose=False):
    """Compile libcdl_b200.so in-tree for sm_100a."""
    return _lib.build(force=force, verbose=verbose)


def load_library():
    return _lib.load()
