"""``DINOHead`` (src/methods/dino.py:32-111 of nicoboou/chadavit) and the DINO training-step engine behind the reference's
method interface, implemented by chadavit_b200.

NOTE for the overlay: the reference's src/methods/dino.py also holds its LightningModule.  To keep Lightning, do not copy
this file; replace the reference's ``class DINOHead`` by ``from chadavit_b200.methods.dino import DINOHead`` instead
(INTEGRATION.md, step 2).  Copying this file swaps the LightningModule for the engine class ``DINO`` (same sub-module
names, hooks and cfg fields; ``fused_train_step`` instead of Lightning's loop)."""
from chadavit_b200.methods.dino import DINO, DINOHead  # noqa: F401

__all__ = ["DINO", "DINOHead"]
