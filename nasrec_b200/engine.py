"""Host-side execution engine: a tiny tape over the C-ABI kernels.

Every operator of the supernet hot path is a pair (forward launch sequence,
backward closure) working on ``Var`` handles.  A ``Var`` is a device tensor plus
a lazily-allocated gradient buffer; the first writer of a gradient overwrites,
later writers accumulate in the kernel epilogue (no separate add kernels, no
zero-padded concats).  Parameters are ``PVar``s whose gradient is the dense
tensor autograd / the optimizers expect.

The operators mirror the reference modules; citations are relative to the NasRec
repository root.  All device work goes through ``nasrec_b200._lib.call`` --
there is no torch arithmetic on this path.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import query


def call(name: str, *args):
    # late-bound so that bench.py's per-kernel timer can wrap _lib.call
    _lib.call(name, *args)


E = 16                 # embedding dim (supernet.py:224)
LN_EPS = 1e-5
FUSED_CALLS = True     # one C call per operator direction (nasrec_linear_ln_* / nasrec_sproj_ln_*)
# When the library forks weight-gradient GEMMs onto a side stream (nasrec_set_side_stream), the scratch
# buffers those GEMMs read must outlive the closure that allocated them: they are parked here until the
# caller has joined the side stream (FusedTrainer.forward_backward).
OVERLAP_KEEP: Optional[list] = None


class Var:
    """Activation handle: data tensor ``t`` (contiguous fp32), gradient ``g``."""
    __slots__ = ("t", "g", "req")

    def __init__(self, t: torch.Tensor, req: bool = False):
        self.t = t
        self.g: Optional[torch.Tensor] = None
        self.req = req

    def grad_buf(self) -> Tuple[torch.Tensor, int]:
        """(gradient tensor, accumulate flag): first toucher overwrites."""
        if self.g is None:
            self.g = torch.empty_like(self.t)
            return self.g, 0
        return self.g, 1


class PVar:
    """Parameter handle. ``g`` is a dense gradient of the parameter's shape."""
    __slots__ = ("p", "t", "g", "req")

    def __init__(self, p: torch.Tensor, req: bool):
        self.p = p
        self.t = p.detach()
        self.g: Optional[torch.Tensor] = None
        self.req = req

    def grad(self, full_overwrite: bool) -> torch.Tensor:
        if self.g is None:
            self.g = torch.empty_like(self.t) if full_overwrite else torch.zeros_like(self.t)
        return self.g


class Tape:
    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.ops: List[Callable[[], None]] = []

    def record(self, fn: Callable[[], None]):
        if self.enabled:
            self.ops.append(fn)

    def backward(self):
        for fn in reversed(self.ops):
            fn()
        self.ops = []


class Seg:
    """One source of a (virtual) zero-padded concat; see nasrec_seg_t."""
    __slots__ = ("v", "off", "ld", "width", "w_off")

    def __init__(self, v: Var, off: int, ld: int, width: int, w_off: int):
        self.v, self.off, self.ld, self.width, self.w_off = v, off, ld, width, w_off


def _p(t: torch.Tensor, off: int = 0) -> int:
    return t.data_ptr() + 4 * off


def _pack(seg_list: Sequence[Seg], grad: bool = False):
    return _lib.segs([(_p(s.v.g if grad else s.v.t, s.off), s.ld, s.width, s.w_off) for s in seg_list])


def _new(*shape, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _any_req(seg_list: Sequence[Seg]) -> bool:
    return any(s.v.req for s in seg_list)


def _unique_woff_groups(seg_list: Sequence[Seg]) -> List[List[Seg]]:
    """Split so that no two segments of a group meet the same weight columns
    (Sum may feed the same source as left and right, modules.py:470-487)."""
    groups: List[List[Seg]] = []
    for s in seg_list:
        if s.width == 0:
            continue
        for g in groups:
            if all(o.w_off != s.w_off for o in g):
                g.append(s)
                break
        else:
            groups.append([s])
    return groups


def _distinct_woffs(seg_list: Sequence[Seg]) -> bool:
    offs = [s.w_off for s in seg_list if s.width > 0]
    return len(offs) == len(set(offs))


def _grad_targets(seg_list: Sequence[Seg]):
    """(nasrec_seg_t[] of gradient targets, accumulate flags) for the fused backward entry points,
    or None when two segments share a target (then the grouped general path is used).  Targets are
    allocated on first touch; a segment whose source needs no gradient gets a null pointer."""
    keys = [(id(s.v), s.off) for s in seg_list if s.v.req and s.width > 0]
    if len(keys) != len(set(keys)):
        return None               # (checked before anything is allocated)
    fresh = set()
    items, flags = [], []
    for s in seg_list:
        if s.v.req and s.width > 0:
            if s.v.g is None:
                s.v.g = torch.empty_like(s.v.t)
                fresh.add(id(s.v))
            flags.append(0 if id(s.v) in fresh else 1)
            items.append((_p(s.v.g, s.off), s.ld, s.width, s.w_off))
        else:
            flags.append(0)
            items.append((0, s.ld, s.width, s.w_off))
    return _lib.segs(items)[0], _lib.i32_array(flags)


def _dgrad_groups(seg_list: Sequence[Seg]):
    """Segments needing a gradient, grouped into launches: (segments, accumulate).
    Within a launch all targets are distinct tensors or disjoint slices."""
    fresh: List[Seg] = []
    acc: List[List[Seg]] = []
    seen_fresh = set()
    for s in seg_list:
        if not s.v.req or s.width == 0:
            continue
        key = (id(s.v), s.off)
        if s.v.g is None or (id(s.v) in seen_fresh and key not in {(id(o.v), o.off) for o in fresh}):
            # first contribution into this tensor in this op: make sure it exists
            if s.v.g is None:
                s.v.g = torch.empty_like(s.v.t)
                seen_fresh.add(id(s.v))
            fresh.append(s)
        else:
            for g in acc:
                if all((id(o.v), o.off) != key for o in g):
                    g.append(s)
                    break
            else:
                acc.append([s])
    out = []
    if fresh:
        out.append((fresh, 0))
    for g in acc:
        out.append((g, 1))
    return out


# Weight planes for the TMA-fed GEMM (nasrec_b200/planes.py).  The trainers that keep planes in step with the weights
# (FusedTrainer and its subclasses) publish their PlaneCache here for the duration of a forward / backward; every linear
# then announces its weight's planes to the library right before its call (a one-entry hint, ignored when it does not
# match the W pointer of the call).  None -> LDG-producer kernel, as with plain autograd.
PLANES = None


def _hint(W: "PVar"):
    c = PLANES
    if c is None:
        return
    pl = c.planes(W.p)
    if pl is None:
        return
    hi, lo, ldp, first = pl
    from ._lib import LIB
    LIB.set_weight_planes(W.t.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, W.t.shape[0], W.t.shape[1], first)


# --------------------------------------------------------------------------- 2-D linear (+LN/act)
def linear_ln(tape: Tape, seg_list: Sequence[Seg], M: int, W: PVar, b: Optional[PVar],
              ln: Optional[Tuple[PVar, PVar]], relu: bool, d_out: int, n_off: int = 0, n_full: Optional[int] = None,
              out: Optional[Var] = None, out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0,
              w_full_support: bool = False) -> Var:
    """y[:, :d_out] = mask_d . act(LN(concat(segs) @ W[n_off:n_off+n_full].T + b))
    (ElasticLinear / Sum / SigmoidGating projection / FM / merger: modules.py:171-181,
    489-499, 584-593, 740-749; supernet.py:1140-1142).  Without LayerNorm only the
    d_out live output columns are computed at all."""
    ref = seg_list[0].v.t
    ldw = W.t.shape[1]
    n_full = (W.t.shape[0] - n_off) if n_full is None else n_full
    N = n_full if ln is not None else d_out
    z = _new(M, N, like=ref)
    sp, ns = _pack(seg_list)
    if out is None:
        out = Var(_new(M, d_out, like=ref))
        ldy = d_out
    if ln is not None:
        mean, rstd = _new(M, like=ref), _new(M, like=ref)
        gam, bet, pm, pr = _p(ln[0].t), _p(ln[1].t), _p(mean), _p(rstd)
    else:
        mean = rstd = None
        gam = bet = pm = pr = None
    if FUSED_CALLS:
        _hint(W)
        call("nasrec_linear_ln_fwd", sp, ns, _p(W.t), ldw, n_off, N, _p(b.t) if b is not None else None, gam, bet,
             LN_EPS, int(relu), d_out, _p(z), _p(out.t, out_off), ldy, pm, pr, accumulate, M)
    else:
        _hint(W)
        call("nasrec_seg_linear_fwd", sp, ns, _p(W.t), ldw, n_off, N, _p(b.t) if b is not None else None, _p(z), N, M)
        if ln is not None:
            call("nasrec_ln_fwd", _p(z), N, M, N, gam, bet, LN_EPS, int(relu), d_out, _p(out.t, out_off), ldy, pm, pr,
                 accumulate)
        else:
            call("nasrec_act_fwd", _p(z), N, M, N, int(relu), _p(out.t, out_off), ldy, accumulate)
    req = _any_req(seg_list) or W.req or (b is not None and b.req) or (ln is not None and (ln[0].req or ln[1].req))
    out.req = out.req or req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        pm, pr = (_p(mean), _p(rstd)) if mean is not None else (None, None)   # (keeps the saved stats alive)
        dz = _new(M, N, like=ref)
        want_ln = ln is not None and (ln[0].req or ln[1].req)
        gw = W.grad(w_full_support and N == W.t.shape[0]) if W.req else None
        gb = b.grad(N == b.t.shape[0]) if (b is not None and b.req) else None
        # (_grad_targets allocates the gradient buffers it hands out: only call it when its result is used)
        targets = _grad_targets(seg_list) if (FUSED_CALLS and _distinct_woffs(seg_list)) else None
        if targets is not None:
            dsp, flags = targets
            _hint(W)
            call("nasrec_linear_ln_bwd", _p(out.g, out_off), ldy, d_out, _p(z), M, N, gam, bet, pm, pr, int(relu), sp,
                 dsp, flags, ns, _p(W.t), ldw, n_off, _p(gw) if gw is not None else None,
                 _p(gb) if gb is not None else None, _p(ln[0].grad(True)) if want_ln else None,
                 _p(ln[1].grad(True)) if want_ln else None, _p(dz))
            if OVERLAP_KEEP is not None:
                OVERLAP_KEEP.append(dz)
            return
        # general path: the same source feeds two segments (Sum with left == right, modules.py:470-487)
        if ln is not None:
            call("nasrec_ln_bwd", _p(out.g, out_off), ldy, d_out, _p(z), N, M, N, gam, bet, pm, pr, int(relu), _p(dz), N,
                 _p(ln[0].grad(True)) if want_ln else None, _p(ln[1].grad(True)) if want_ln else None, 0)
        else:
            call("nasrec_act_bwd", _p(out.g, out_off), ldy, _p(z), N, M, N, int(relu), _p(dz), N)
        if gw is not None:
            for gi, grp in enumerate(_unique_woff_groups(seg_list)):
                spk, nsk = _pack(grp)
                call("nasrec_seg_linear_wgrad", _p(dz), N, N, spk, nsk, _p(gw), ldw, n_off, M, 1 if gi else 0)
        if gb is not None:
            call("nasrec_colsum", _p(dz), N, M, N, _p(gb, n_off), 0)
        for grp, acc in _dgrad_groups(seg_list):
            spk, nsk = _pack(grp, grad=True)
            _hint(W)
            call("nasrec_seg_linear_dgrad", _p(dz), N, N, _p(W.t), ldw, n_off, spk, nsk, M, acc)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- sparse-axis projection (+LN/act)
def sproj_ln(tape: Tape, seg_list: Sequence[Seg], B: int, W: PVar, b: Optional[PVar],
             ln: Optional[Tuple[PVar, PVar]], relu: bool, p_out: int, out: Optional[Var] = None, out_off: int = 0,
             out_bstride: Optional[int] = None, accumulate: int = 0, w_full_support: bool = False) -> Var:
    """y[b, p<p_out, :] = mask_p . act(LN_P(W @ concat_rows(segs)[b] + bias))
    (ElasticLinear3D modules.py:222-235; DotProduct._sparse_inp_proj :358-361;
    Transformer._linear_proj :648-662)."""
    ref = seg_list[0].v.t
    ldw = W.t.shape[1]
    P_full = W.t.shape[0]
    P = P_full if ln is not None else p_out
    z = _new(B, P, E, like=ref)
    sp, ns = _pack(seg_list)
    if out is None:
        out = Var(_new(B, p_out, E, like=ref))
        out_bstride = p_out * E
    if ln is not None:
        mean, rstd = _new(B, E, like=ref), _new(B, E, like=ref)
        gam, bet, pm, pr = _p(ln[0].t), _p(ln[1].t), _p(mean), _p(rstd)
    else:
        mean = rstd = None
        gam = bet = pm = pr = None
    if FUSED_CALLS:
        _hint(W)
        call("nasrec_sproj_ln_fwd", sp, ns, _p(W.t), ldw, P, _p(b.t) if b is not None else None, gam, bet, LN_EPS,
             int(relu), p_out, _p(z), _p(out.t, out_off), out_bstride, pm, pr, accumulate, B)
    else:
        _hint(W)
        call("nasrec_sproj_fwd", sp, ns, _p(W.t), ldw, P, _p(b.t) if b is not None else None, _p(z), P * E, B)
        if ln is not None:
            call("nasrec_ln3_fwd", _p(z), P * E, B, P, gam, bet, LN_EPS, int(relu), p_out, _p(out.t, out_off),
                 out_bstride, pm, pr, accumulate)
        else:
            call("nasrec_act_fwd", _p(z), P * E, B, p_out * E, int(relu), _p(out.t, out_off), out_bstride, accumulate)
    req = _any_req(seg_list) or W.req or (b is not None and b.req) or (ln is not None and (ln[0].req or ln[1].req))
    out.req = out.req or req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        pm, pr = (_p(mean), _p(rstd)) if mean is not None else (None, None)   # (keeps the saved stats alive)
        dz = _new(B, P, E, like=ref)
        want_ln = ln is not None and (ln[0].req or ln[1].req)
        gw = W.grad(w_full_support and P == P_full) if W.req else None
        gb = b.grad(P == P_full) if (b is not None and b.req) else None
        # (_grad_targets allocates the gradient buffers it hands out: only call it when its result is used)
        targets = _grad_targets(seg_list) if (FUSED_CALLS and _distinct_woffs(seg_list)) else None
        if targets is not None:
            dsp, flags = targets
            ws = None
            if gw is not None:
                ws = _new(query("nasrec_sproj_wgrad_ws_floats", P, sum(s.width for s in seg_list), B), like=ref)
            _hint(W)
            call("nasrec_sproj_ln_bwd", _p(out.g, out_off), out_bstride, p_out, _p(z), B, P, gam, bet, pm, pr, int(relu),
                 sp, dsp, flags, ns, _p(W.t), ldw, _p(gw) if gw is not None else None,
                 _p(gb) if gb is not None else None, _p(ln[0].grad(True)) if want_ln else None,
                 _p(ln[1].grad(True)) if want_ln else None, _p(dz), _p(ws) if ws is not None else None)
            if OVERLAP_KEEP is not None:
                OVERLAP_KEEP.extend((dz, ws))
            return
        if ln is not None:
            call("nasrec_ln3_bwd", _p(out.g, out_off), out_bstride, p_out, _p(z), P * E, B, P, gam, bet, pm, pr,
                 int(relu), _p(dz), P * E, _p(ln[0].grad(True)) if want_ln else None,
                 _p(ln[1].grad(True)) if want_ln else None, 0)
        else:
            call("nasrec_act_bwd", _p(out.g, out_off), out_bstride, _p(z), P * E, B, P * E, int(relu), _p(dz), P * E)
        if gw is not None:
            for gi, grp in enumerate(_unique_woff_groups(seg_list)):
                tw = sum(s.width for s in grp)
                ws = _new(query("nasrec_sproj_wgrad_ws_floats", P, tw, B), like=ref)
                spk, nsk = _pack(grp)
                call("nasrec_sproj_wgrad", _p(dz), P * E, P, spk, nsk, _p(gw), ldw, B, 1 if gi else 0, _p(ws))
        if gb is not None:
            call("nasrec_sproj_bias_grad", _p(dz), P * E, P, B, _p(gb), 0)
        for grp, acc in _dgrad_groups(seg_list):
            spk, nsk = _pack(grp, grad=True)
            _hint(W)
            call("nasrec_sproj_dgrad", _p(dz), P * E, P, _p(W.t), ldw, spk, nsk, B, acc)

    tape.record(bwd)
    return out


def ln3_only(tape: Tape, x: Var, B: int, P: int, ln: Tuple[PVar, PVar], relu: bool, p_out: int) -> Var:
    """LayerNorm over the row axis of an existing [B,P,16] tensor (no projection)."""
    ref = x.t
    out = Var(_new(B, p_out, E, like=ref))
    mean, rstd = _new(B, E, like=ref), _new(B, E, like=ref)
    call("nasrec_ln3_fwd", _p(x.t), P * E, B, P, _p(ln[0].t), _p(ln[1].t), LN_EPS, int(relu), p_out, _p(out.t),
         p_out * E, _p(mean), _p(rstd), 0)
    req = x.req or ln[0].req or ln[1].req
    out.req = req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        want = ln[0].req or ln[1].req
        dz = _new(B, P, E, like=ref) if x.req else None
        call("nasrec_ln3_bwd", _p(out.g), p_out * E, p_out, _p(x.t), P * E, B, P, _p(ln[0].t), _p(ln[1].t), _p(mean),
             _p(rstd), int(relu), _p(dz) if dz is not None else None, P * E,
             _p(ln[0].grad(True)) if want else None, _p(ln[1].grad(True)) if want else None, 0)
        if dz is not None:
            _accumulate_into(x, dz)

    tape.record(bwd)
    return out


def _accumulate_into(v: Var, d: torch.Tensor):
    """v.g (+)= d for same-shaped contiguous tensors."""
    if v.g is None:
        v.g = d
        return
    n = d.numel()
    call("nasrec_act_fwd", _p(d), n, 1, n, 0, _p(v.g), n, 1)


# --------------------------------------------------------------------------- row LayerNorm on an existing tensor
def ln_rows(tape: Tape, x: Var, M: int, N: int, ln: Optional[Tuple[PVar, PVar]], relu: bool, d_out: int) -> Var:
    """out = mask . act(LN(x)) for a materialised [M,N] input (no projection in front:
    the R == dims / width == dims corners, modules.py:386-392, 488-493, 583-589)."""
    ref = x.t
    out = Var(_new(M, d_out, like=ref))
    if ln is not None:
        mean, rstd = _new(M, like=ref), _new(M, like=ref)
        call("nasrec_ln_fwd", _p(x.t), N, M, N, _p(ln[0].t), _p(ln[1].t), LN_EPS, int(relu), d_out, _p(out.t), d_out,
             _p(mean), _p(rstd), 0)
    else:
        mean = rstd = None
        call("nasrec_act_fwd", _p(x.t), N, M, d_out, int(relu), _p(out.t), d_out, 0)
    req = x.req or (ln is not None and (ln[0].req or ln[1].req))
    out.req = req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        if ln is not None:
            want = ln[0].req or ln[1].req
            dx = _new(M, N, like=ref) if x.req else None
            call("nasrec_ln_bwd", _p(out.g), d_out, d_out, _p(x.t), N, M, N, _p(ln[0].t), _p(ln[1].t), _p(mean),
                 _p(rstd), int(relu), _p(dx) if dx is not None else None, N,
                 _p(ln[0].grad(True)) if want else None, _p(ln[1].grad(True)) if want else None, 0)
        else:
            dx = torch.zeros(M, N, dtype=torch.float32, device=ref.device) if d_out < N else _new(M, N, like=ref)
            call("nasrec_act_bwd", _p(out.g), d_out, _p(x.t), N, M, d_out, int(relu), _p(dx), N)
        if dx is not None and x.req:
            _accumulate_into(x, dx)

    tape.record(bwd)
    return out


def concat2d(tape: Tape, seg_list: Sequence[Seg], M: int, width: int, extra: Sequence[Seg] = ()) -> Var:
    """Materialise a zero-padded concat [M,width] (plus `extra` added on top, the
    left+right of Sum); only used where the reference skips its projection."""
    ref = seg_list[0].v.t
    out = Var(torch.zeros(M, width, dtype=torch.float32, device=ref.device))
    for lst, acc in ((seg_list, 0), (extra, 1)):
        for grp in _unique_woff_groups(lst):
            sp, ns = _pack(grp)
            call("nasrec_concat_segs", sp, ns, _p(out.t), width, M, acc)
            acc = 1
    req = _any_req(seg_list) or _any_req(extra)
    out.req = req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        for s in list(seg_list) + list(extra):
            if not s.v.req or s.width == 0:
                continue
            g, acc = s.v.grad_buf()
            # dsrc[m, k] (+)= dout[m, w_off + k]
            sp, ns = _lib.segs([(_p(out.g, s.w_off), width, s.width, 0)])
            call("nasrec_concat_segs", sp, ns, _p(g, s.off), s.ld, M, acc)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- DotProduct interaction
def dot_tril(tape: Tape, x: Var, y: Var, B: int, P: int) -> Var:
    """R = strict lower triangle of [x;y][x;y]^T, modules.py:366-383."""
    R = (P + 1) * P // 2
    out = Var(_new(B, R, like=x.t))
    call("nasrec_dot_tril_fwd", _p(x.t), E, _p(y.t), P * E, P, _p(out.t), R, B)
    out.req = x.req or y.req
    if not (tape.enabled and out.req):
        return out

    def bwd():
        if out.g is None:
            return
        dx = _new(B, E, like=x.t) if x.req else None
        dy = _new(B, P, E, like=x.t) if y.req else None
        call("nasrec_dot_tril_bwd", _p(out.g), R, _p(x.t), E, _p(y.t), P * E, P,
             _p(dx) if dx is not None else None, E, _p(dy) if dy is not None else None, P * E, B)
        if dx is not None:
            _accumulate_into(x, dx)
        if dy is not None:
            _accumulate_into(y, dy)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- SigmoidGating core
def gate(tape: Tape, pre: Var, right: Sequence[Seg], M: int, K: int) -> Var:
    """out = sigmoid(pre) * right  (modules.py:580-582) on the live columns of right."""
    out = Var(_new(M, K, like=pre.t))
    sp, ns = _pack(right)
    call("nasrec_gate_fwd", _p(pre.t), K, sp, ns, _p(out.t), K, M)
    out.req = pre.req or _any_req(right)
    if not (tape.enabled and out.req):
        return out

    def bwd():
        if out.g is None:
            return
        dpre = _new(M, K, like=pre.t)
        # right gradients: fresh targets first, then accumulating ones (separate launches)
        fresh, acc = [], []
        for s in right:
            if s.v.req and s.v.g is None:
                s.v.g = torch.empty_like(s.v.t)
                fresh.append(s)
            elif s.v.req:
                acc.append(s)
        null = (0, 0, 0, 0)
        for accf, chosen in ((0, fresh), (1, acc), (0, None)):
            if chosen is None:
                if fresh or acc:
                    continue
                chosen = []          # nobody wants dright: still need dpre
            elif not chosen:
                continue
            ids = {id(s) for s in chosen}
            d_items = [(_p(s.v.g, s.off), s.ld, s.width, s.w_off) if id(s) in ids else null for s in right]
            dsp, _ = _lib.segs(d_items)
            call("nasrec_gate_bwd", _p(out.g), K, _p(pre.t), K, sp, dsp, ns, _p(dpre), K, M, accf)
        if pre.req:
            _accumulate_into(pre, dpre)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- FM
def fm_ix(tape: Tape, x: Var, B: int, rows: int, bstride: int) -> Var:
    """ix = (sum_r x)^2 - sum_r x^2 over the live rows, modules.py:736-738."""
    out = Var(_new(B, E, like=x.t))
    call("nasrec_fm_fwd", _p(x.t), bstride, rows, _p(out.t), B)
    out.req = x.req
    if not (tape.enabled and out.req):
        return out

    def bwd():
        if out.g is None:
            return
        if x.g is None:
            x.g = torch.zeros_like(x.t) if rows * E < bstride else torch.empty_like(x.t)
            call("nasrec_fm_bwd", _p(out.g), _p(x.t), bstride, rows, None, 0, _p(x.g), bstride, B)
        else:
            call("nasrec_fm_bwd", _p(out.g), _p(x.t), bstride, rows, _p(x.g), bstride, _p(x.g), bstride, B)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- attention core
def attention(tape: Tape, x: Var, B: int, L: int, s_live: int, params: Sequence[PVar], out: Var, out_bstride: int,
              accumulate_out: bool = False) -> Var:
    """MHA(16, 8 heads) + residual + LN + FC-ReLU-FC + residual + LN on [B, s_live, 16]
    (tokens s_live..L-1 are the reference's zero rows), modules.py:664-688."""
    pp = _lib.ptr_array([_p(p.t) for p in params])
    if accumulate_out:
        tmp = Var(_new(B, s_live, E, like=x.t))
        call("nasrec_attn_fwd", _p(x.t), s_live * E, L, s_live, pp, _p(tmp.t), s_live * E, B)
        call("nasrec_act_fwd", _p(tmp.t), s_live * E, B, s_live * E, 0, _p(out.t), out_bstride, 1)
    else:
        call("nasrec_attn_fwd", _p(x.t), s_live * E, L, s_live, pp, _p(out.t), out_bstride, B)
    req = x.req or any(p.req for p in params)
    out.req = out.req or req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        dx = _new(B, s_live, E, like=x.t) if x.req else None
        want = any(p.req for p in params)
        dpar = _new(_lib.ATTN_PARAMS, like=x.t) if want else None
        ws = _new(query("nasrec_attn_bwd_ws_floats", B), like=x.t)
        call("nasrec_attn_bwd", _p(out.g), out_bstride, _p(x.t), s_live * E, L, s_live, pp,
             _p(dx) if dx is not None else None, s_live * E, _p(dpar) if dpar is not None else None, 0, _p(ws), B)
        if dpar is not None:
            o = 0
            for p in params:
                n = p.t.numel()
                if p.req:
                    p.g = dpar[o:o + n].view(p.t.shape)
                o += n
        if dx is not None:
            _accumulate_into(x, dx)

    tape.record(bwd)
    return out


# --------------------------------------------------------------------------- copies
def copy2d(tape: Tape, src: Var, src_off: int, lds: int, M: int, N: int, dst: Var, dst_off: int, ldd: int,
           accumulate: int = 0):
    """dst[m, :N] (+)= src[m, :N]; gradient flows back to src."""
    call("nasrec_act_fwd", _p(src.t, src_off), lds, M, N, 0, _p(dst.t, dst_off), ldd, accumulate)
    dst.req = dst.req or src.req
    if not (tape.enabled and src.req):
        return

    def bwd():
        if dst.g is None:
            return
        if src.g is None:
            full = (src_off == 0 and N == lds and src.t.numel() == M * N)
            src.g = torch.empty_like(src.t) if full else torch.zeros_like(src.t)
            call("nasrec_act_fwd", _p(dst.g, dst_off), ldd, M, N, 0, _p(src.g, src_off), lds, 0)
        else:
            call("nasrec_act_fwd", _p(dst.g, dst_off), ldd, M, N, 0, _p(src.g, src_off), lds, 1)

    tape.record(bwd)


# --------------------------------------------------------------------------- embedding stem
class EmbeddingTables:
    """Device-side pointer tables for the fused gather (rebuilt when storage moves)."""

    def __init__(self):
        self.key = None
        self.ptrs = None
        self.rows = None
        self.err = None

    def refresh(self, weights: Sequence[torch.Tensor]):
        key = tuple((w.data_ptr(), w.shape[0]) for w in weights)
        if key != self.key:
            dev = weights[0].device
            self.ptrs = torch.tensor([k[0] for k in key], dtype=torch.int64, device=dev)
            self.rows = torch.tensor([k[1] for k in key], dtype=torch.int64, device=dev)
            self.err = torch.zeros(1, dtype=torch.int32, device=dev)
            self.key = key
        return self

    def check(self):
        """Raise IndexError if any embedding id seen so far was outside its table (synchronises; call it where the
        caller synchronises anyway).  The reference fails through nn.Embedding's index assert (supernet.py:407)."""
        if self.err is not None and int(self.err.item()) != 0:
            self.err.zero_()
            raise IndexError("embedding index out of range (cat_feats must lie in [0, num_embeddings[f]))")


class SparseEmbGrad:
    """Result of the deterministic sorted-row reduction for one step."""
    __slots__ = ("uniq", "nuniq", "row_grad", "sumsq", "B", "F", "ws")


class SparseSink(list):
    """Receives the embedding gradient of a step instead of dense [N_f,16] tensors.
    defer=True stores the raw (cat_x, d_out) pair (data-parallel training gathers
    them across ranks before the sorted-row reduction)."""
    defer = False


def reduce_sparse(cat_x: torch.Tensor, gout: torch.Tensor, tables: Optional["EmbeddingTables"] = None) -> SparseEmbGrad:
    """Deterministic sorted-row reduction of d(sparse)[B,F,16] (nasrec_emb_grad_sort_reduce).  With ``tables`` the ids are
    bounds-checked as in the forward gather: out-of-range ids are dropped and ``tables.err`` is raised."""
    B, F = cat_x.shape
    dev = cat_x.device
    sg = SparseEmbGrad()
    sg.B, sg.F = B, F
    sg.uniq = torch.empty(F, B, dtype=torch.int64, device=dev)
    sg.nuniq = torch.empty(F, dtype=torch.int32, device=dev)
    sg.row_grad = torch.empty(F, B, E, dtype=torch.float32, device=dev)
    sg.sumsq = torch.empty(F, dtype=torch.float32, device=dev)
    scratch = torch.empty(F, B + 1, dtype=torch.int32, device=dev)
    if B > 2048 and F <= 31:
        # large batches (data-parallel global batch, KDD): multi-CTA radix-sort reduction (csrc/emb_big.cu)
        from ._lib import query
        nb = query("nasrec_emb_grad_sort_reduce_big_ws_bytes", B, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        has = tables is not None and tables.rows is not None
        call("nasrec_emb_grad_sort_reduce_big", cat_x.data_ptr(), _p_i(tables.rows) if has else None,
             _p_i(tables.err) if has else None, _p(gout), B, F, sg.uniq.data_ptr(), sg.nuniq.data_ptr(), _p(sg.row_grad),
             _p(sg.sumsq), ws.data_ptr(), nb)
        sg.ws = ws
    elif tables is not None and tables.rows is not None:
        call("nasrec_emb_grad_sort_reduce_checked", cat_x.data_ptr(), _p_i(tables.rows), _p_i(tables.err), _p(gout), B, F,
             sg.uniq.data_ptr(), sg.nuniq.data_ptr(), _p(sg.row_grad), _p(sg.sumsq), scratch.data_ptr())
    else:
        call("nasrec_emb_grad_sort_reduce", cat_x.data_ptr(), _p(gout), B, F, sg.uniq.data_ptr(), sg.nuniq.data_ptr(),
             _p(sg.row_grad), _p(sg.sumsq), scratch.data_ptr())
    return sg


def embedding(tape: Tape, tables: EmbeddingTables, weights: Sequence[PVar], cat_x: torch.Tensor,
              sparse_sink: Optional[list] = None, cache: Optional[dict] = None) -> Var:
    """sparse[b,f,:] = W_f[cat[b,f],:]  (supernet.py:412-430).  Backward: sorted-row
    reduction; dense [N_f,16] grads (reference layout) unless `sparse_sink` is given,
    in which case the reduced rows are handed to the fused optimizer instead."""
    B, F = cat_x.shape
    req = any(w.req for w in weights)
    if cache is not None and not req:
        # frozen tables (one-shot candidate scoring): one gather per batch, shared by every candidate
        hit = cache.get(cat_x.data_ptr())
        if hit is not None:
            return Var(hit)
    tables.refresh([w.t for w in weights])
    out = Var(torch.empty(B, F, E, dtype=torch.float32, device=cat_x.device))
    call("nasrec_emb_gather_fwd", _p_i(tables.ptrs), _p_i(tables.rows), _p_i(cat_x), _p(out.t), B, F,
         _p_i(tables.err))
    if cache is not None and not req:
        cache[cat_x.data_ptr()] = out.t
    out.req = req
    if not (tape.enabled and req):
        return out

    def bwd():
        if out.g is None:
            return
        dev = cat_x.device
        if sparse_sink is not None and getattr(sparse_sink, "defer", False):
            sparse_sink.append((cat_x, out.g))
            return
        sg = reduce_sparse(cat_x, out.g, tables)
        if sparse_sink is not None:
            sparse_sink.append(sg)
            return
        grads = [w.grad(False) for w in weights]
        gp = torch.tensor([g.data_ptr() for g in grads], dtype=torch.int64, device=dev)
        call("nasrec_emb_grad_to_dense", _p_i(sg.uniq), _p_i(sg.nuniq), _p(sg.row_grad), _p_i(gp), B, F)

    tape.record(bwd)
    return out


def _p_i(t: torch.Tensor) -> int:
    return t.data_ptr()


# --------------------------------------------------------------------------- loss
def bce_with_logits(logits: torch.Tensor, y: torch.Tensor, grad_scale: float = 1.0, want_grad: bool = True):
    """mean BCE-with-logits and d(loss)/d(logits) in one launch (train_utils.py:266)."""
    B = logits.numel()
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    call("nasrec_bce_fwd_bwd", _p(logits), _p(y), B, float(grad_scale), _p(loss), _p(dl) if dl is not None else None)
    return loss, dl
