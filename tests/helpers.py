"""Shared helpers for the parity tests: deterministic weights/inputs (oracle.det) and the golden fixtures."""
import json
import os

import numpy as np
import torch

from oracle import chada_oracle as O
from oracle import det

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


def cases():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        return json.load(f)


def det_params(shapes, salt):
    return {k: torch.from_numpy(v) for k, v in det.det_state_dict(shapes, salt).items()}


def backbone_case(c):
    """(params dict, pixels, nhead, final_eps) of a golden backbone case."""
    P = det_params(O.backbone_shapes(c["D"], max_ch=c["max_ch"]), c["seed"])
    x = torch.from_numpy(det.det_pixels(sum(c["counts"]), c["hw"], c["hw"], c["seed"]))
    nhead, eps = (2, 1e-6) if c["ctor"] == "factory" else (12, 1e-5)
    return P, x, nhead, eps


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
