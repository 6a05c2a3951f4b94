"""Shared helpers for the parity tests (fixtures under tests/golden/)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    npz = os.path.join(GOLDEN, name + ".npz")
    arrays = dict(np.load(npz)) if os.path.exists(npz) else {}
    return meta, arrays


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def relu_kink_margin(sd, cfg, choice, int_x, cat_x):
    """Smallest |pre-activation| over every ReLU of the oracle forward, relative to the median
    |pre-activation| of the same call.  fp32 implementations agree on a ReLU's side only up to
    the rounding error of its input (~1e-6 relative after a K~1e3 contraction and a LayerNorm):
    a unit closer to zero than that may legitimately fall on either side, which changes every
    upstream gradient by O(1/sqrt(#active units)) ~ 1e-2 on the 5-sample fixtures.  Tests use this
    to decide whether a case can be held to the strict gradient tolerance."""
    import torch
    from oracle import nasrec_oracle as orc
    rec = []
    orig = torch.relu

    def spy(x):
        a = x.detach().abs()
        med = float(a.median())
        if med > 0:
            rec.append(float(a.min()) / med)
        return orig(x)

    torch.relu = spy
    try:
        with torch.no_grad():
            orc.supernet_forward(sd, cfg, choice, int_x, cat_x)
    finally:
        torch.relu = orig
    return min(rec) if rec else 1.0
