"""CPU: the C-ABI shared library loads without a GPU/driver and exports exactly the symbols include/chadavit_b200.h declares;
the ctypes table in chadavit_b200/_lib.py lists each of them with the declared number of arguments."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "chadavit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(cb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_library_builds_loads_and_exports_every_declared_symbol():
    from chadavit_b200 import _lib
    from chadavit_b200.build import build
    path = build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)                      # must not need libcuda / a GPU to load
    decls = _declared()
    assert len(decls) >= 25
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        if name == "cb_last_error":
            continue
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes table"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, ctypes table {len(_lib.SIGNATURES[name])}"
    for name in _lib.SIGNATURES:
        assert name in decls, f"{name} bound in _lib.py but not declared in the header"
    assert _lib.load().cb_version() == 1


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import pytest
    import torch
    from chadavit_b200.backbones import chada_vit
    from chadavit_b200.losses import DINOLoss
    from chadavit_b200.methods import DINOHead
    m = chada_vit(patch_size=16, embed_dim=32, return_all_tokens=False, max_number_channels=10)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 1, 32, 32), 0, [[2]])
    with pytest.raises(RuntimeError):
        DINOHead(32, 64, use_bn=False)(torch.zeros(2, 32))
    with pytest.raises(RuntimeError):
        DINOLoss(64, 0.04, 0.07, 0, 10)(torch.zeros(4, 64), torch.zeros(4, 64))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "chadavit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"
