"""B200-native mirror of ``nasrec/supernet/modules.py``.

Same class names, constructor kwargs, sub-module names (hence state-dict keys),
lazy-materialisation life cycle and ``forward(tensor..., dims_in_use)``
signatures as the reference, but the arithmetic is the fused CUDA path of
``nasrec_b200.engine`` (C ABI in include/nasrec_b200.h).  Modules are parameter
containers plus two entry points:

* ``forward(...)``   -- the reference's public API on ordinary (zero-padded) tensors;
* ``_run(run, ...)`` -- the segment-list form used by SuperNet, which never
  materialises the zero padding (compact tensors, K-support skipping).

Citations name the reference lines each piece follows (NasRec repo root).
"""
from __future__ import annotations

from math import sqrt
from typing import List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .. import engine as eng
from ..engine import PVar, Seg, Tape, Var

NUM_MHA_HEADS = 8            # modules.py:26
LN_INIT = 0.17               # modules.py:598
EMB = 16


def apply_activation_fn(x, activation):      # modules.py:35-36 (host convenience only)
    raise NotImplementedError("activations are fused into the CUDA epilogues; call a module instead")


def _relu_flag(activation: str) -> bool:
    if activation == "relu":
        return True
    if activation == "identity":
        return False
    raise NotImplementedError("activation '%s' has no fused sm_100a epilogue (relu|identity)" % activation)


class FLAGS:                                   # modules.py:41-54
    def __init__(self):
        self.DEBUG = False

    def config_debug(self, debug: bool = False):
        self.DEBUG = debug


flags = FLAGS()


class CleverMaskGenerator:
    """modules.py:57-96.  Kept for API compatibility; the CUDA path applies prefix
    masks as an epilogue predicate (col < dims_in_use) and never builds them."""

    def __init__(self):
        self.cached_mask = {}

    def __call__(self, max_dims_or_dims: int, dims_in_use: int, device=None):
        assert max_dims_or_dims >= dims_in_use, \
            "'max_dims_or_dims' should be larger than 'dims_in_use' to successfully generate a mask."
        token = "{}_{}_{}".format(max_dims_or_dims, dims_in_use, device)
        if token in self.cached_mask and not flags.DEBUG:
            return self.cached_mask[token]
        mask = torch.zeros(max_dims_or_dims, device=device)
        mask[:dims_in_use] = 1.0
        self.cached_mask[token] = mask
        return mask


class CleverZeroTensorGenerator:
    """modules.py:99-127 (cache keyed by shape AND device, fixing the reference's
    shape-only key)."""

    def __init__(self):
        self.cached_zeros = {}

    def __call__(self, size, device=None):
        token = "_".join(str(x) for x in size) + "@" + str(device)
        if token in self.cached_zeros and not flags.DEBUG:
            return self.cached_zeros[token]
        z = torch.zeros(size, dtype=torch.float, device=device)
        self.cached_zeros[token] = z
        return z


_mask_generator = CleverMaskGenerator()
_zeros_generator = CleverZeroTensorGenerator()


# --------------------------------------------------------------------------- run context
class Run:
    """Per-forward context: the tape plus the parameter handles."""

    def __init__(self, tape: Tape, grad_ok=None, sparse_sink: Optional[list] = None,
                 emb_cache: Optional[dict] = None):
        self.tape = tape
        self._pv = {}
        self._grad_ok = grad_ok          # None: honour requires_grad; else set of id(param) allowed
        self.sparse_sink = sparse_sink
        self.emb_cache = emb_cache

    def pv(self, p: Optional[torch.Tensor]) -> Optional[PVar]:
        if p is None:
            return None
        h = self._pv.get(id(p))
        if h is None:
            req = self.tape.enabled and bool(p.requires_grad)
            if self._grad_ok is not None:
                req = req and (id(p) in self._grad_ok)
            h = PVar(p, req)
            self._pv[id(p)] = h
        return h

    def ln(self, m: Optional[nn.LayerNorm]):
        return None if m is None else (self.pv(m.weight), self.pv(m.bias))

    def grad_of(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        h = self._pv.get(id(p))
        return None if h is None else h.g

    def touched(self) -> List[PVar]:
        return list(self._pv.values())


def _materialize(lin: nn.Module, in_features: int):
    """nn.LazyLinear -> nn.Linear exactly as its first forward would (class swap,
    hook removal, reset_parameters) but without running a torch forward
    (reference behaviour at modules.py:154 etc.)."""
    if isinstance(lin, nn.LazyLinear) and lin.has_uninitialized_params():
        fake = torch.empty(1, int(in_features), device="meta")
        lin._infer_parameters(lin, (fake,))
    elif getattr(lin, "in_features", in_features) != in_features:
        raise ValueError("linear layer was materialised for %d input features, got %d"
                         % (lin.in_features, in_features))


def _whole(v: Var, width: int, w_off: int = 0) -> Seg:
    return Seg(v, 0, v.t.shape[-1] if v.t.dim() == 2 else v.t.shape[1] * EMB, width, w_off)


class _TapeFn(torch.autograd.Function):
    """Bridges a tape-recorded launch sequence into torch.autograd."""

    @staticmethod
    def forward(ctx, body, n_in, params, *tensors):
        inputs, ptensors = tensors[:n_in], tensors[n_in:]
        need = ctx.needs_input_grad[3:]
        tape = Tape(any(need))
        ok = {id(p) for p, nd_ in zip(params, need[n_in:]) if nd_}
        run = Run(tape, grad_ok=ok)
        in_vars = []
        for t, nd_ in zip(inputs, need[:n_in]):
            if t.is_floating_point():
                t = t.detach().contiguous().float()
            in_vars.append(Var(t, req=bool(nd_)))
        outs = body(run, in_vars)
        ctx.tape, ctx.run, ctx.in_vars, ctx.outs, ctx.params = tape, run, in_vars, outs, params
        return tuple(o.t for o in outs)

    @staticmethod
    def backward(ctx, *gouts):
        for o, g in zip(ctx.outs, gouts):
            o.g = None if g is None else g.contiguous()
        ctx.tape.backward()
        gi = [v.g if v.req else None for v in ctx.in_vars]
        gp = [ctx.run.grad_of(p) for p in ctx.params]
        return (None, None, None, *gi, *gp)


def run_with_autograd(module: nn.Module, inputs: Sequence[torch.Tensor], body):
    """Execute ``body(run, in_vars) -> [Var]`` with `module`'s parameters tracked."""
    if not inputs[0].is_cuda:
        raise RuntimeError("nasrec_b200 runs on CUDA tensors only (no CPU fallback); got device %s"
                           % inputs[0].device)
    params = [p for p in module.parameters() if not isinstance(p, nn.parameter.UninitializedParameter)]
    # parameters materialised inside body are picked up on the next call; the
    # warm-up forward is a no-grad pass in every reference entry point.
    return _TapeFn.apply(body, len(inputs), params, *inputs, *params)


def _pad_out(v: Var, M: int, width: int, full: int, like: torch.Tensor) -> Var:
    """Reference-shaped (zero-padded) output buffer for the standalone module API."""
    return Var(torch.zeros(M, full, dtype=torch.float32, device=like.device)) if full != width else v


# --------------------------------------------------------------------------- FC
class ElasticLinear(nn.Module):
    """modules.py:134-181."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._activation = kwargs["activation"]
        self._use_layernorm = kwargs["use_layernorm"]
        self._fixed = fixed
        self._linear = nn.LazyLinear(self._max_dims_or_dims, bias=not self._use_layernorm)
        self._layernorm = nn.LayerNorm([self._max_dims_or_dims]) if self._use_layernorm else None

    def _prepare(self, in_features: int):
        _materialize(self._linear, in_features)

    def _run(self, run: Run, segs: Sequence[Seg], K: int, M: int, dims_in_use: int, out: Optional[Var] = None,
             out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0) -> Var:
        self._prepare(K)
        d = self._max_dims_or_dims if self._fixed else dims_in_use
        return eng.linear_ln(run.tape, segs, M, run.pv(self._linear.weight), run.pv(self._linear.bias),
                             run.ln(self._layernorm), _relu_flag(self._activation), d, out=out, out_off=out_off,
                             ldy=ldy, accumulate=accumulate, w_full_support=self._fixed)

    def forward(self, tensor, dims_in_use):
        if not self._fixed:
            assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
        M, K = tensor.shape
        full = self._max_dims_or_dims
        self._prepare(K)

        def body(run, iv):
            out = Var(torch.zeros(M, full, dtype=torch.float32, device=tensor.device))
            self._run(run, [_whole(iv[0], K)], K, M, dims_in_use, out=out, ldy=full)
            return [out]

        return run_with_autograd(self, [tensor], body)[0]


# --------------------------------------------------------------------------- EFC
class ElasticLinear3D(nn.Module):
    """modules.py:184-235."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._activation = kwargs["activation"]
        self._use_layernorm = kwargs["use_layernorm"]
        self._fixed = fixed
        self._linear = nn.LazyLinear(self._max_dims_or_dims, bias=not self._use_layernorm)
        self._layernorm = nn.LayerNorm([self._max_dims_or_dims]) if self._use_layernorm else None

    def _prepare(self, in_rows: int):
        _materialize(self._linear, in_rows)

    def _run(self, run: Run, segs: Sequence[Seg], S: int, B: int, dims_in_use: int, out: Optional[Var] = None,
             out_off: int = 0, out_bstride: Optional[int] = None, accumulate: int = 0) -> Var:
        self._prepare(S)
        p_out = self._max_dims_or_dims if self._fixed else dims_in_use
        return eng.sproj_ln(run.tape, segs, B, run.pv(self._linear.weight), run.pv(self._linear.bias),
                            run.ln(self._layernorm), _relu_flag(self._activation), p_out, out=out, out_off=out_off,
                            out_bstride=out_bstride, accumulate=accumulate, w_full_support=self._fixed)

    def forward(self, tensor, dims_in_use):
        assert len(tensor.size()) == 3, "Tensor should be 3D!"
        if not self._fixed:
            assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
        B, S, _ = tensor.shape
        full = self._max_dims_or_dims
        self._prepare(S)

        def body(run, iv):
            out = Var(torch.zeros(B, full, EMB, dtype=torch.float32, device=tensor.device))
            self._run(run, [Seg(iv[0], 0, S * EMB, S, 0)], S, B, dims_in_use, out=out, out_bstride=full * EMB)
            return [out]

        return run_with_autograd(self, [tensor], body)[0]


# --------------------------------------------------------------------------- zeros
class Zeros2D(nn.Module):
    """modules.py:238-270 (contributes nothing to the node sum)."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._fixed = fixed

    def forward(self, dense_t: torch.Tensor, dims_in_use: int):
        assert len(dense_t.size()) == 2, ValueError("Input tensor to 'Zeros2D' should have a 2D shape.")
        if not self._fixed:
            assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
            return _zeros_generator(torch.Size((dense_t.size(0), self._max_dims_or_dims)), dense_t.device)
        return _zeros_generator(torch.Size((dense_t.size(0), dims_in_use)), dense_t.device)


class Zeros3D(nn.Module):
    """modules.py:691-718."""

    def __init__(self, **kwargs):
        super().__init__()
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]

    def forward(self, sparse_t: torch.Tensor, dims_in_use: int):
        assert len(sparse_t.size()) == 3, ValueError("Input must have a shape of 3D!")
        assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
        return _zeros_generator(torch.Size((sparse_t.size(0), self._max_dims_or_dims, sparse_t.size(2))),
                                sparse_t.device)


# --------------------------------------------------------------------------- DotProduct
class DotProduct(nn.Module):
    """modules.py:273-401."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._use_layernorm = kwargs["use_layernorm"]
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._embedding_dim = kwargs["embedding_dim"]
        self._fixed = fixed
        ln = self._use_layernorm
        self._dense_proj = nn.LazyLinear(self._embedding_dim, bias=not ln)
        self._sparse_proj = nn.LazyLinear(self._embedding_dim, bias=not ln)
        self.sparse_inp_proj_dim = round(sqrt(2 * self._max_dims_or_dims))
        self._sparse_inp_proj = nn.LazyLinear(self.sparse_inp_proj_dim, bias=not ln)
        self._linear_proj = nn.LazyLinear(self._max_dims_or_dims, bias=not ln)
        self._dense_layernorm = nn.LayerNorm(self._embedding_dim) if ln else None
        self._sparse_layernorm = nn.LayerNorm(self._embedding_dim) if ln else None
        self._sparse_inp_proj_layernorm = nn.LayerNorm(self.sparse_inp_proj_dim) if ln else None
        self._linear_layernorm = nn.LayerNorm(self._max_dims_or_dims) if ln else None

    def _prepare(self, dense_width: int, sparse_rows: int, emb: int = EMB):
        """Materialise / delete sub-layers exactly where the reference's first forward does
        (modules.py:339-364, 384-389)."""
        if emb != self._embedding_dim:
            raise NotImplementedError("sparse width must equal embedding_dim (always true in NASRec)")
        if dense_width != self._embedding_dim:
            _materialize(self._dense_proj, dense_width)
        else:
            self._dense_proj, self._dense_layernorm = None, None
        self._sparse_proj, self._sparse_layernorm = None, None
        P = self.sparse_inp_proj_dim
        if sparse_rows != P:
            _materialize(self._sparse_inp_proj, sparse_rows)
        else:
            self._sparse_inp_proj, self._sparse_inp_proj_layernorm = None, None
        R = (P + 1) * P // 2
        if R != self._max_dims_or_dims:
            _materialize(self._linear_proj, R)
        else:
            self._linear_proj = None

    def _run(self, run: Run, dsegs: Sequence[Seg], Kd: int, ssegs: Sequence[Seg], S: int, B: int, dims_in_use: int,
             out: Optional[Var] = None, out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0) -> Var:
        self._prepare(Kd, S)
        tape = run.tape
        P = self.sparse_inp_proj_dim
        full = self._max_dims_or_dims
        d = full if self._fixed else dims_in_use
        if self._dense_proj is not None:
            x = eng.linear_ln(tape, dsegs, B, run.pv(self._dense_proj.weight), run.pv(self._dense_proj.bias),
                              run.ln(self._dense_layernorm), False, EMB, w_full_support=self._fixed)
        else:
            x = eng.concat2d(tape, dsegs, B, EMB)
        if self._sparse_inp_proj is not None:
            y = eng.sproj_ln(tape, ssegs, B, run.pv(self._sparse_inp_proj.weight),
                             run.pv(self._sparse_inp_proj.bias), run.ln(self._sparse_inp_proj_layernorm), False, P,
                             w_full_support=self._fixed)
        else:
            y = _concat_rows(tape, ssegs, B, P)
        R = eng.dot_tril(tape, x, y, B, P)
        nR = (P + 1) * P // 2
        if self._linear_proj is not None:
            return eng.linear_ln(tape, [Seg(R, 0, nR, nR, 0)], B, run.pv(self._linear_proj.weight),
                                 run.pv(self._linear_proj.bias), run.ln(self._linear_layernorm), False, d, out=out,
                                 out_off=out_off, ldy=ldy, accumulate=accumulate, w_full_support=True)
        y2 = eng.ln_rows(tape, R, B, nR, run.ln(self._linear_layernorm), False, d)
        return _deliver(tape, y2, B, d, out, out_off, ldy, accumulate)

    def forward(self, dense_t: torch.Tensor, sparse_t: torch.Tensor, dims_in_use: int):
        assert len(dense_t.size()) == 2, ValueError("Dense tensor should be 2D!")
        assert len(sparse_t.size()) == 3, ValueError("Sparse tensor should be 3D!")
        if not self._fixed:
            assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
        B, Kd = dense_t.shape
        S = sparse_t.shape[1]
        full = self._max_dims_or_dims
        self._prepare(Kd, S)

        def body(run, iv):
            out = Var(torch.zeros(B, full, dtype=torch.float32, device=dense_t.device))
            self._run(run, [_whole(iv[0], Kd)], Kd, [Seg(iv[1], 0, S * EMB, S, 0)], S, B, dims_in_use, out=out,
                      ldy=full)
            return [out]

        return run_with_autograd(self, [dense_t, sparse_t], body)[0]


def _concat_rows(tape: Tape, ssegs: Sequence[Seg], B: int, rows: int) -> Var:
    """[B, rows, 16] materialisation of a sparse segment list (a 2-D concat on the
    flattened rows); only for the S == P corner of DotProduct (modules.py:362-364)."""
    flat = [Seg(s.v, s.off, s.ld, s.width * EMB, s.w_off * EMB) for s in ssegs]
    v = eng.concat2d(tape, flat, B, rows * EMB)
    v.t = v.t.view(B, rows, EMB)
    return v


def _deliver(tape: Tape, y: Var, M: int, d: int, out: Optional[Var], out_off: int, ldy: Optional[int],
             accumulate: int) -> Var:
    if out is None:
        return y
    eng.copy2d(tape, y, 0, d, M, d, out, out_off, ldy, accumulate)
    return out


def _pad_2Dtensors_if_needed(left_2d_tensor: torch.Tensor, right_2d_tensor: torch.Tensor):
    """modules.py:403-430 (public helper; the CUDA path pads virtually through segment offsets)."""
    assert len(left_2d_tensor.size()) == 2 and len(right_2d_tensor.size()) == 2
    sl, sr = left_2d_tensor.size(-1), right_2d_tensor.size(-1)
    if sl == sr:
        return left_2d_tensor, right_2d_tensor
    z = _zeros_generator(torch.Size((left_2d_tensor.size(0), abs(sl - sr))), device=left_2d_tensor.device)
    if sl < sr:
        return torch.cat([left_2d_tensor, z], dim=1), right_2d_tensor
    return left_2d_tensor, torch.cat([right_2d_tensor, z], dim=1)


# --------------------------------------------------------------------------- Sum
class Sum(nn.Module):
    """modules.py:432-501."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._use_layernorm = kwargs["use_layernorm"]
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._linear_proj = nn.LazyLinear(self._max_dims_or_dims, bias=not self._use_layernorm)
        self._layernorm = nn.LayerNorm(self._max_dims_or_dims) if self._use_layernorm else None
        self._fixed = fixed

    def _prepare(self, width: int):
        if width != self._max_dims_or_dims:
            _materialize(self._linear_proj, width)
        else:
            self._linear_proj = None

    def _run(self, run: Run, lsegs: Sequence[Seg], Kl: int, rsegs: Sequence[Seg], Kr: int, M: int, dims_in_use: int,
             out: Optional[Var] = None, out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0) -> Var:
        D = max(Kl, Kr)                               # _pad_2Dtensors_if_needed
        self._prepare(D)
        full = self._max_dims_or_dims
        d = full if self._fixed else dims_in_use
        if self._linear_proj is not None:
            return eng.linear_ln(run.tape, list(lsegs) + list(rsegs), M, run.pv(self._linear_proj.weight),
                                 run.pv(self._linear_proj.bias), run.ln(self._layernorm), False, d, out=out,
                                 out_off=out_off, ldy=ldy, accumulate=accumulate,
                                 w_full_support=self._fixed and Kl == Kr)
        z = eng.concat2d(run.tape, lsegs, M, D, extra=rsegs)
        y = eng.ln_rows(run.tape, z, M, D, run.ln(self._layernorm), False, d)
        return _deliver(run.tape, y, M, d, out, out_off, ldy, accumulate)

    def forward(self, left_2d: torch.Tensor, right_2d: torch.Tensor, dims_in_use: int):
        assert len(left_2d.size()) == 2, ValueError("Left tensor should have a shape of 2D!")
        assert len(right_2d.size()) == 2, ValueError("Right tensor should have a shape of 2D!")
        M, Kl = left_2d.shape
        Kr = right_2d.shape[1]
        full = self._max_dims_or_dims
        self._prepare(max(Kl, Kr))

        def body(run, iv):
            out = Var(torch.zeros(M, full, dtype=torch.float32, device=left_2d.device))
            self._run(run, [_whole(iv[0], Kl)], Kl, [_whole(iv[1], Kr)], Kr, M, dims_in_use, out=out, ldy=full)
            return [out]

        return run_with_autograd(self, [left_2d, right_2d], body)[0]


# --------------------------------------------------------------------------- SigmoidGating
class LazySelfLinear(nn.Module):
    """modules.py:504-519."""

    def __init__(self):
        super().__init__()
        self._linear = None
        self._linear_size: int = -1

    def _prepare(self, size: int, device):
        if self._linear is None:
            self._linear = nn.Linear(size, size, bias=True).to(device)
            self._linear_size = size
        assert size == self._linear_size, "'LazySelfLinear' inconsistent size: {} vs {}".format(
            self._linear_size, size)

    def forward(self, x):
        self._prepare(x.size(-1), x.device)
        M, K = x.shape

        def body(run, iv):
            return [eng.linear_ln(run.tape, [_whole(iv[0], K)], M, run.pv(self._linear.weight),
                                  run.pv(self._linear.bias), None, False, K, w_full_support=True)]

        return run_with_autograd(self, [x], body)[0]


class SigmoidGating(nn.Module):
    """modules.py:521-595: linear(right * sigmoid(self_linear(left)))."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._use_layernorm = kwargs["use_layernorm"]
        self._fixed = fixed
        self._left_self_linear = LazySelfLinear()
        self._linear_proj = nn.LazyLinear(self._max_dims_or_dims, bias=True)
        self._layernorm = nn.LayerNorm(self._max_dims_or_dims) if self._use_layernorm else None

    def _prepare(self, width: int, device):
        self._left_self_linear._prepare(width, device)
        if width != self._max_dims_or_dims:
            _materialize(self._linear_proj, width)
        else:
            self._linear_proj = None

    def _run(self, run: Run, lsegs: Sequence[Seg], Kl: int, rsegs: Sequence[Seg], Kr: int, M: int, dims_in_use: int,
             out: Optional[Var] = None, out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0) -> Var:
        D = max(Kl, Kr)
        dev = lsegs[0].v.t.device
        self._prepare(D, dev)
        tape = run.tape
        full = self._max_dims_or_dims
        d = full if self._fixed else dims_in_use
        sl = self._left_self_linear._linear
        Wself, bself = run.pv(sl.weight), run.pv(sl.bias)
        live = [s for s in rsegs if s.width > 0]
        Kg = sum(s.width for s in live)
        # sigmoid(left @ Wself^T + b) is only needed where `right` can be non-zero:
        # rows of Wself in supp(right), columns in supp(left).
        pre = Var(torch.empty(M, Kg, dtype=torch.float32, device=dev))
        ko = 0
        for s in live:
            eng.linear_ln(tape, lsegs, M, Wself, bself, None, False, s.width, n_off=s.w_off, n_full=s.width,
                          out=pre, out_off=ko, ldy=Kg, w_full_support=False)
            ko += s.width
        gated = eng.gate(tape, pre, live, M, Kg)
        gsegs, ko = [], 0
        for s in live:
            gsegs.append(Seg(gated, ko, Kg, s.width, s.w_off))
            ko += s.width
        if self._linear_proj is not None:
            return eng.linear_ln(tape, gsegs, M, run.pv(self._linear_proj.weight), run.pv(self._linear_proj.bias),
                                 run.ln(self._layernorm), False, d, out=out, out_off=out_off, ldy=ldy,
                                 accumulate=accumulate, w_full_support=self._fixed and Kr == D)
        z = eng.concat2d(tape, gsegs, M, D)
        y = eng.ln_rows(tape, z, M, D, run.ln(self._layernorm), False, d)
        return _deliver(tape, y, M, d, out, out_off, ldy, accumulate)

    def forward(self, left_2d: torch.Tensor, right_2d: torch.Tensor, dims_in_use: int):
        assert len(left_2d.size()) == 2, ValueError("Left tensor should have a shape of 2D!")
        assert len(right_2d.size()) == 2, ValueError("Right tensor should have a shape of 2D!")
        if not self._fixed:
            assert dims_in_use <= self._max_dims_or_dims, ValueError("'dims_in_use' > 'max_dims_or_dims'")
        M, Kl = left_2d.shape
        Kr = right_2d.shape[1]
        full = self._max_dims_or_dims
        self._prepare(max(Kl, Kr), left_2d.device)

        def body(run, iv):
            out = Var(torch.zeros(M, full, dtype=torch.float32, device=left_2d.device))
            self._run(run, [_whole(iv[0], Kl)], Kl, [_whole(iv[1], Kr)], Kr, M, dims_in_use, out=out, ldy=full)
            return [out]

        return run_with_autograd(self, [left_2d, right_2d], body)[0]


# --------------------------------------------------------------------------- Transformer
class Transformer(nn.Module):
    """modules.py:599-688."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._use_layernorm = kwargs["use_layernorm"]
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._activation = kwargs["activation"]
        self._embedding_dim = kwargs["embedding_dim"]
        if self._embedding_dim != EMB or self._max_dims_or_dims > 64:
            raise NotImplementedError("fused attention kernel is built for embedding_dim=16, <=64 tokens")
        self._linear_proj = nn.LazyLinear(self._max_dims_or_dims, bias=not self._use_layernorm)
        self._proj_ln = nn.LayerNorm(self._max_dims_or_dims) if self._use_layernorm else None
        self._mha = nn.MultiheadAttention(self._embedding_dim, num_heads=NUM_MHA_HEADS, batch_first=True)
        self._attn_ln = nn.LayerNorm(self._embedding_dim, eps=1e-5)
        self.attn_fc1 = nn.LazyLinear(self._embedding_dim)
        self.attn_fc2 = nn.LazyLinear(self._embedding_dim)
        self._attn_fc_ln = nn.LayerNorm(self._embedding_dim, eps=1e-5)
        self._dropout = kwargs["dropout"] if "dropout" in kwargs else 0.0     # stored, never applied (:632)
        self._fixed = fixed
        torch.nn.init.constant_(self._attn_ln.weight, LN_INIT)
        torch.nn.init.constant_(self._attn_fc_ln.weight, LN_INIT)

    def _prepare(self, in_rows: int):
        _materialize(self._linear_proj, in_rows)
        _materialize(self.attn_fc1, self._embedding_dim)
        _materialize(self.attn_fc2, self._embedding_dim)

    def _attn_params(self, run: Run):
        m = self._mha
        return [run.pv(m.in_proj_weight), run.pv(m.in_proj_bias), run.pv(m.out_proj.weight),
                run.pv(m.out_proj.bias), run.pv(self._attn_ln.weight), run.pv(self._attn_ln.bias),
                run.pv(self.attn_fc1.weight), run.pv(self.attn_fc1.bias), run.pv(self.attn_fc2.weight),
                run.pv(self.attn_fc2.bias), run.pv(self._attn_fc_ln.weight), run.pv(self._attn_fc_ln.bias)]

    def _run(self, run: Run, segs: Sequence[Seg], S: int, B: int, dims_in_use: int, out: Optional[Var] = None,
             out_off: int = 0, out_bstride: Optional[int] = None, accumulate: int = 0) -> Var:
        self._prepare(S)
        L = self._max_dims_or_dims
        s_live = L if self._fixed else dims_in_use
        xa = eng.sproj_ln(run.tape, segs, B, run.pv(self._linear_proj.weight), run.pv(self._linear_proj.bias),
                          run.ln(self._proj_ln), False, s_live, w_full_support=self._fixed)
        if out is None:
            out = Var(torch.empty(B, s_live, EMB, dtype=torch.float32, device=xa.t.device))
            out_off, out_bstride = 0, s_live * EMB
        if out_off != 0:
            raise ValueError("attention output must start at row 0 of its buffer")
        return eng.attention(run.tape, xa, B, L, s_live, self._attn_params(run), out, out_bstride,
                             accumulate_out=bool(accumulate))

    def forward(self, sparse_t: torch.Tensor, dims_in_use: int):
        assert len(sparse_t.size()) == 3, ValueError("Input must have a shape of 3D!")
        B, S, _ = sparse_t.shape
        full = self._max_dims_or_dims
        self._prepare(S)

        def body(run, iv):
            out = Var(torch.zeros(B, full, EMB, dtype=torch.float32, device=sparse_t.device))
            self._run(run, [Seg(iv[0], 0, S * EMB, S, 0)], S, B, dims_in_use, out=out, out_bstride=full * EMB)
            return [out]

        return run_with_autograd(self, [sparse_t], body)[0]


# --------------------------------------------------------------------------- FM
class FactorizationMachine3D(nn.Module):
    """modules.py:720-750."""

    def __init__(self, fixed: bool = False, **kwargs):
        super().__init__()
        self._use_layernorm = kwargs["use_layernorm"]
        self._max_dims_or_dims = kwargs["max_dims_or_dims"]
        self._linear_proj = nn.LazyLinear(self._max_dims_or_dims, bias=not self._use_layernorm)
        self._fixed = fixed
        if self._use_layernorm:
            self._linear_layernorm = nn.LayerNorm(self._max_dims_or_dims, eps=1e-5)

    def _prepare(self, emb: int = EMB):
        if emb != self._max_dims_or_dims:
            _materialize(self._linear_proj, emb)
        else:
            self._linear_proj, self._use_layernorm = None, None      # modules.py:743

    def _run(self, run: Run, x: Var, rows: int, bstride: int, B: int, dims_in_use: int, out: Optional[Var] = None,
             out_off: int = 0, ldy: Optional[int] = None, accumulate: int = 0) -> Var:
        self._prepare()
        tape = run.tape
        full = self._max_dims_or_dims
        d = full if self._fixed else dims_in_use
        ix = eng.fm_ix(tape, x, B, rows, bstride)
        if self._linear_proj is not None:
            ln = run.ln(self._linear_layernorm) if self._use_layernorm else None
            return eng.linear_ln(tape, [Seg(ix, 0, EMB, EMB, 0)], B, run.pv(self._linear_proj.weight),
                                 run.pv(self._linear_proj.bias), ln, False, d, out=out, out_off=out_off, ldy=ldy,
                                 accumulate=accumulate, w_full_support=True)
        if not self._fixed and d < EMB:
            y = eng.ln_rows(tape, ix, B, EMB, None, False, d)
            return _deliver(tape, y, B, d, out, out_off, ldy, accumulate)
        return _deliver(tape, ix, B, EMB, out, out_off, ldy, accumulate)

    def forward(self, sparse_t: torch.Tensor, dims_in_use: int):
        assert len(sparse_t.size()) == 3, "Tensor must be a sparse tensor!"
        B, S, _ = sparse_t.shape
        full = self._max_dims_or_dims
        self._prepare()

        def body(run, iv):
            out = Var(torch.zeros(B, full, dtype=torch.float32, device=sparse_t.device))
            self._run(run, iv[0], S, S * EMB, B, dims_in_use, out=out, ldy=full)
            return [out]

        return run_with_autograd(self, [sparse_t], body)[0]
