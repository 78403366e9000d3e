"""Drop-in for src/utils/lars.py of nicoboou/chadavit (``LARS``): same constructor and ``torch.optim.Optimizer`` surface, the
step runs on the flat-arena kernels (chadavit_b200/utils/lars.py)."""
from chadavit_b200.utils.lars import LARS  # noqa: F401

__all__ = ["LARS"]
