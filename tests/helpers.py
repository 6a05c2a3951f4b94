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


def relu_kink_candidates(sd, cfg, choice, int_x, cat_x, margin=6e-6, most=6):
    """ReLU inputs of the oracle forward that lie within `margin` (relative to the median |input| of their call) of
    zero: [(relative distance, relu call index, flat element index, value)], closest first.  These are the units whose
    side two correct fp32 implementations may legitimately disagree on."""
    import torch
    from oracle import nasrec_oracle as orc
    rec = []
    orig = torch.relu

    def spy(x):
        a = x.detach().abs().flatten()
        med = float(a.median())
        if med > 0:
            near = torch.nonzero(a < margin * med).flatten().tolist()
            for i in near[:most]:
                rec.append((float(a[i]) / med, spy.calls, int(i), float(x.detach().flatten()[i])))
        spy.calls += 1
        return orig(x)

    spy.calls = 0
    torch.relu = spy
    try:
        with torch.no_grad():
            orc.supernet_forward(sd, cfg, choice, int_x, cat_x)
    finally:
        torch.relu = orig
    return sorted(rec)[:most]


class flipped_relu:
    """Context manager: inside it, the oracle's ReLU call number `call` sees element `idx` moved across zero (a constant
    offset of minus twice its value on that one element -- everything stays differentiable)."""

    def __init__(self, call, idx, val):
        self.call, self.idx, self.val = call, idx, val

    def __enter__(self):
        import torch
        self.orig = torch.relu
        count = [0]

        def relu(x):
            c = count[0]
            count[0] += 1
            if c == self.call:
                off = torch.zeros_like(x).flatten()
                off[self.idx] = -2.0 * self.val
                x = x + off.view_as(x)
            return self.orig(x)

        torch.relu = relu
        return self

    def __exit__(self, *exc):
        import torch
        torch.relu = self.orig
        return False
