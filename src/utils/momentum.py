"""Drop-in for src/utils/momentum.py of nicoboou/chadavit (``MomentumUpdater``, ``initialize_momentum_params``): the EMA is one
launch over the flat parameter arenas (chadavit_b200/utils/momentum.py)."""
from chadavit_b200.utils.momentum import MomentumUpdater, initialize_momentum_params  # noqa: F401

__all__ = ["MomentumUpdater", "initialize_momentum_params"]
