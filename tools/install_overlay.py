#!/usr/bin/env python
"""Copy the overlay tree (src/**, leaf modules only) over a checkout of nicoboou/chadavit.

    python tools/install_overlay.py /path/to/chadavit [--with-methods-dino] [--dry-run]

Leaves every ``__init__.py`` of the reference alone.  ``src/methods/dino.py`` is copied only with --with-methods-dino (it
replaces the reference's LightningModule by the engine class; the default is the two-line patch of INTEGRATION.md).
chadavit_b200 itself must be importable in the target environment (``pip install -e`` this repository or add it to PYTHONPATH)."""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEAVES = ["src/backbones/vit/chada_vit.py", "src/losses/dino.py", "src/utils/momentum.py", "src/utils/lars.py"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("target")
    ap.add_argument("--with-methods-dino", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    leaves = LEAVES + (["src/methods/dino.py"] if a.with_methods_dino else [])
    for rel in leaves:
        dst = os.path.join(a.target, rel)
        if not os.path.exists(dst):
            raise SystemExit(f"{dst} does not exist: is {a.target} a checkout of nicoboou/chadavit?")
        print(("would copy " if a.dry_run else "copy ") + rel)
        if not a.dry_run:
            shutil.copy2(dst, dst + ".orig")
            shutil.copy2(os.path.join(ROOT, rel), dst)


if __name__ == "__main__":
    main()
