"""Native step executor binding (nasrec_b200/csrc/net.cu).

The Python engine is the reference-shaped, autograd-compatible path; at B=512 it is host-bound
(~3 ms of interpreter time per step against ~2.2 ms of device work).  ``NativeNet`` describes a
materialised weight-sharing ``SuperNet`` to the C++ executor once (parameter table + per-block node
tables) and then drives whole steps through it: same operators, same order, same kernels -- results
are bit-identical to the Python engine (tests/test_gpu_native.py) -- with activations and gradients in
two caller-owned arenas.  Fixed (standalone) models and unusual corners stay on the Python engine:
``NativeNet.unsupported_reason(model)`` says why.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .planes import PlaneCache
from .supernet.supernet import EMB, SuperNet, _ints
from .utils.train_utils import FusedTrainer

_NODE_TYPES = {"linear-2d": 0, "dot-product": 1, "sum": 2, "sigmoid-gating": 3, "linear-3d": 4, "transformer": 5,
               "zeros-2d": 6, "zeros-3d": 7}
_CHOICE_STRIDE = 5 * 9 + 4
_ENOSPACE = -3
_vp, _i, _l, _fl = C.c_void_p, C.c_int, C.c_int64, C.c_float

_PROTOS = {
    "nasrec_net_create": ([_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _vp),
    "nasrec_net_destroy": ([_vp], None),
    "nasrec_net_set_arenas": ([_vp, _vp, _l, _vp, _l], _i),
    "nasrec_net_set_requires_grad": ([_vp, _vp, _i], _i),
    "nasrec_net_set_planes": ([_vp, _vp, _vp, _vp, _vp, _i], _i),
    "nasrec_net_set_overlap": ([_vp, _i], _i),
    "nasrec_net_set_defer_wgrad": ([_vp, _i], _i),
    "nasrec_net_set_reserve": ([_vp, _i], _i),
    "nasrec_net_set_seal_callback": ([_vp, _vp], _i),
    "nasrec_net_forward": ([_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp], _i),
    "nasrec_multi_subnet_eval": ([_vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp], _i),
    "nasrec_net_forward_backward": ([_vp, _vp, _vp, _vp, _vp, _i, _fl, _vp, _vp, _vp], _i),
    "nasrec_net_grad_bucket": ([_vp, _vp, _vp], _i),
    "nasrec_net_sparse_raw": ([_vp, _vp, _vp], _i),
    "nasrec_net_sparse_reduce": ([_vp, _vp, _vp, _i, _vp], _i),
    "nasrec_net_apply": ([_vp, _fl, _fl, _fl, _vp, _vp], _i),
    "nasrec_net_launches": ([], _l),
    "nasrec_net_arena_high_water": ([_vp, _i], _l),
    "nasrec_set_side_stream": ([_vp], _i),
}
EXPORTS = list(_PROTOS)
_fns: Dict[str, Any] = {}


def _fn(name: str):
    f = _fns.get(name)
    if f is None:
        f = getattr(_lib.LIB.load().cdll, name)
        f.argtypes, f.restype = _PROTOS[name]
        _fns[name] = f
    return f


def _check(rc: int, what: str):
    if rc != 0:
        if rc > 0:
            raise RuntimeError("%s failed: CUDA error %d" % (what, rc))
        raise ValueError("%s rejected its arguments (code %d)" % (what, rc))


class NativeNet:
    """One materialised weight-sharing SuperNet as seen by the C++ executor."""

    @staticmethod
    def unsupported_reason(model: SuperNet) -> Optional[str]:
        if not isinstance(model, SuperNet):
            return "not a SuperNet"
        if model._fixed:
            return "fixed (standalone) models run on the Python engine / CUDA graphs"
        if model._needs_materialize():
            return "model is not materialised yet"
        for blk in model._blocks:
            if getattr(blk, "_activation", "relu") != "relu":
                return "activation other than relu"
            if blk._max_dims_or_dims_sparse > 64:
                return "more than 64 sparse rows"
            if blk.deep_fm is None or blk.deep_fm._linear_proj is None:
                return "FM without projection (max dense width 16)"
            for i, node in enumerate(blk._nodes):
                name = blk._node_names[i]
                if name not in _NODE_TYPES:
                    return "node '%s'" % name
                if name == "dot-product" and (node._dense_proj is None or node._sparse_inp_proj is None
                                              or node._linear_proj is None):
                    return "DotProduct corner without a projection"
                if name in ("sum", "sigmoid-gating") and node._linear_proj is None:
                    return "%s corner without a projection" % name
        return None

    def __init__(self, model: SuperNet, state_of=None, act_bytes: int = 256 << 20, pgrad_bytes: int = 128 << 20):
        why = self.unsupported_reason(model)
        if why is not None:
            raise NotImplementedError("native executor: " + why)
        self.model = model
        self.dev = model._final.weight.device
        self.params: List[torch.Tensor] = [p for _, p in model.named_parameters()]
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._state_of = state_of
        self.states = [state_of(p) for p in self.params] if state_of is not None else None
        self._build()
        self._alloc_arenas(act_bytes, pgrad_bytes)
        self._launch_mark = int(_fn("nasrec_net_launches")())

    # ---------------------------------------------------------------- description
    def _pi(self, p) -> int:
        return -1 if p is None else self._index[id(p)]

    def _lin(self, m) -> List[int]:
        return [-1, -1] if m is None else [self._pi(m.weight), self._pi(m.bias)]

    def _ln(self, m) -> List[int]:
        return [-1, -1] if m is None else [self._pi(m.weight), self._pi(m.bias)]

    def _node(self, name: str, node) -> List[int]:
        t = _NODE_TYPES[name]
        if name in ("linear-2d", "linear-3d"):
            p = self._lin(node._linear) + self._ln(node._layernorm)
        elif name == "dot-product":
            p = (self._lin(node._dense_proj) + self._ln(node._dense_layernorm) + self._lin(node._sparse_inp_proj)
                 + self._ln(node._sparse_inp_proj_layernorm) + self._lin(node._linear_proj)
                 + self._ln(node._linear_layernorm))
        elif name == "sum":
            p = self._lin(node._linear_proj) + self._ln(node._layernorm)
        elif name == "sigmoid-gating":
            p = self._lin(node._left_self_linear._linear) + self._lin(node._linear_proj) + self._ln(node._layernorm)
        elif name == "transformer":
            m = node._mha
            attn = [m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias, node._attn_ln.weight,
                    node._attn_ln.bias, node.attn_fc1.weight, node.attn_fc1.bias, node.attn_fc2.weight,
                    node.attn_fc2.bias, node._attn_fc_ln.weight, node._attn_fc_ln.bias]
            p = self._lin(node._linear_proj) + self._ln(node._proj_ln) + [self._pi(q) for q in attn]
        else:
            p = []
        return [t] + p + [-1] * (24 - len(p))

    def _build(self):
        m = self.model
        F = m._sparse_input_size
        desc = [m._num_blocks, int(m._num_dense_features), F, self._pi(m._final.weight), self._pi(m._final.bias)]
        desc += [self._pi(e.weight) for e in m._embedding]
        for blk in m._blocks:
            dp_P = 0
            for i, node in enumerate(blk._nodes):
                if blk._node_names[i] == "dot-product":
                    dp_P = node.sparse_inp_proj_dim
            fm = blk.deep_fm
            fm_ln = getattr(fm, "_linear_layernorm", None) if fm._use_layernorm else None
            desc += [len(blk._nodes), blk._max_dims_or_dims_dense, blk._max_dims_or_dims_sparse, dp_P]
            desc += self._lin(blk.project_emb_dim) + self._ln(blk.project_emb_dim_layernorm)
            desc += self._lin(fm._linear_proj) + self._ln(fm_ln)
            for i, node in enumerate(blk._nodes):
                desc += self._node(blk._node_names[i], node)
        self._desc = np.asarray(desc, dtype=np.int32)
        n = len(self.params)
        self._w = (C.c_void_p * n)(*[p.data_ptr() for p in self.params])
        self._s = (C.c_void_p * n)(*[s.data_ptr() for s in self.states]) if self.states is not None else None
        self._numel = np.asarray([p.numel() for p in self.params], dtype=np.int64)
        self._rows = np.asarray([p.shape[0] if p.dim() >= 1 else 1 for p in self.params], dtype=np.int32)
        self._cols = np.asarray([p.shape[1] if p.dim() == 2 else 1 for p in self.params], dtype=np.int32)
        self._req_list = [p.requires_grad for p in self.params]
        self._req = np.asarray(self._req_list, dtype=np.int32)
        m._tables.refresh([e.weight.detach() for e in m._embedding])
        tb = m._tables
        emb_states = [self.states[self._pi(e.weight)] for e in m._embedding] if self.states is not None else None
        self._emb_w = torch.tensor([e.weight.data_ptr() for e in m._embedding], dtype=torch.int64, device=self.dev)
        self._emb_s = (torch.tensor([s.data_ptr() for s in emb_states], dtype=torch.int64, device=self.dev)
                       if emb_states is not None else None)
        self._key = self._signature()
        self.handle = _fn("nasrec_net_create")(
            self._desc.ctypes.data, len(desc), n, C.cast(self._w, C.c_void_p),
            C.cast(self._s, C.c_void_p) if self._s is not None else None, self._numel.ctypes.data,
            self._rows.ctypes.data, self._cols.ctypes.data, self._req.ctypes.data, tb.ptrs.data_ptr(),
            tb.rows.data_ptr(), self._emb_w.data_ptr(), self._emb_s.data_ptr() if self._emb_s is not None else None,
            tb.err.data_ptr())
        if not self.handle:
            raise RuntimeError("nasrec_net_create failed")
        self._plane_gen = -1
        self._sync_planes()

    def _sync_planes(self):
        """Weight planes for the TMA-fed GEMM (nasrec_b200/planes.py): rebuild what changed behind the executor's
        back and (re)announce the pointers when plane storage was (re)allocated."""
        cache = PlaneCache.of(self.model)
        emb = {id(e.weight) for e in self.model._embedding}
        if self._plane_gen < 0:
            self._plane_params = [p for p in self.params if id(p) not in emb and PlaneCache.wanted(p)]
            self._vsum = -1
        vsum = 0
        for p in self._plane_params:          # cheap per-step check; storage moves are caught by refresh()'s signature
            vsum += p._version
        if vsum == self._vsum and cache.epoch == _lib.LIB.weights_epoch and cache.generation == self._plane_gen:
            return
        cache.sync(self._plane_params)
        self._vsum = vsum
        if cache.generation != self._plane_gen:
            n = len(self.params)
            hi, lo, ld, first = [None] * n, [None] * n, [0] * n, [0] * n
            for p in self._plane_params:
                h, l, ldp, f = cache.planes(p)
                i = self._index[id(p)]
                hi[i], lo[i], ld[i], first[i] = h.data_ptr(), l.data_ptr(), ldp, f
            self._ldp = np.asarray(ld, dtype=np.int64)
            self._first = np.asarray(first, dtype=np.int32)
            _check(_fn("nasrec_net_set_planes")(self.handle, C.cast((C.c_void_p * n)(*hi), C.c_void_p),
                                                C.cast((C.c_void_p * n)(*lo), C.c_void_p), self._ldp.ctypes.data,
                                                self._first.ctypes.data, n),
                   "nasrec_net_set_planes")
            self._plane_gen = cache.generation

    def _signature(self):
        return (self.params[0].data_ptr(), self.params[-1].data_ptr(), len(self.params))

    def _alloc_arenas(self, act_bytes: int, pgrad_bytes: int):
        self.act = torch.empty(int(act_bytes), dtype=torch.uint8, device=self.dev)
        self.pg = torch.empty(int(pgrad_bytes), dtype=torch.uint8, device=self.dev)
        _check(_fn("nasrec_net_set_arenas")(self.handle, self.act.data_ptr(), self.act.numel(), self.pg.data_ptr(),
                                            self.pg.numel()), "nasrec_net_set_arenas")

    def _grow(self):
        """An arena overflowed: the executor's high-water marks say how much the step needs."""
        hi_a = int(_fn("nasrec_net_arena_high_water")(self.handle, 0))
        hi_p = int(_fn("nasrec_net_arena_high_water")(self.handle, 1))
        a, p = self.act.numel(), self.pg.numel()
        na = max(2 * a, int(1.25 * hi_a)) if hi_a > a else a
        np_ = max(2 * p, int(1.25 * hi_p)) if hi_p > p else p
        if na == a and np_ == p:
            na, np_ = 2 * a, 2 * p
        if max(na, np_) > (64 << 30):
            raise MemoryError("native executor arena would exceed 64 GiB")
        torch.cuda.synchronize()                # also NCCL work that may still be reading the old gradient bucket
        self.act = self.pg = None
        self._alloc_arenas(na, np_)

    def refresh(self):
        """Re-describe the model if parameter storage moved; push requires_grad flags."""
        if self._signature() != self._key:
            _fn("nasrec_net_destroy")(self.handle)
            if self._state_of is not None:
                self.states = [self._state_of(p) for p in self.params]
            self._build()
            _check(_fn("nasrec_net_set_arenas")(self.handle, self.act.data_ptr(), self.act.numel(), self.pg.data_ptr(),
                                                self.pg.numel()), "nasrec_net_set_arenas")
        self._sync_planes()
        req = [p.requires_grad for p in self.params]
        if req != self._req_list:
            self._req_list = req
            self._req = np.asarray(req, dtype=np.int32)
            _check(_fn("nasrec_net_set_requires_grad")(self.handle, self._req.ctypes.data, len(req)), "set_requires_grad")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _fn("nasrec_net_destroy")(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---------------------------------------------------------------- per step
    @staticmethod
    def encode_choice(macro: Sequence[Dict[str, Any]], micro: Sequence[Dict[str, Any]]) -> np.ndarray:
        flat: List[int] = []
        for mac, mic in zip(macro, micro):
            for key in ("dense_idx", "sparse_idx", "dense_left_idx", "dense_right_idx"):
                v = mac[key]
                if type(v) is not list:
                    v = _ints(v)
                n = len(v)
                if n > 8:
                    raise ValueError("more than 8 sources in '%s'" % key)
                flat.append(n)
                flat += v
                flat += (0,) * (8 - n)
            act = mic["active_nodes"]
            if type(act) is not list:
                act = _ints(act)
            n = len(act)
            if n > 8:
                raise ValueError("more than 8 active nodes")
            flat.append(n)
            flat += act
            flat += (0,) * (8 - n)
            flat += (mic["dense_in_dims"], mic["sparse_in_dims"], mic["dense_sparse_interact"], mic["deep_fm"])
        return np.asarray(flat, dtype=np.int32)

    def _count(self):
        now = int(_fn("nasrec_net_launches")())
        _lib.LIB.launches += now - self._launch_mark
        self._launch_mark = now

    def _retrying(self, what: str, fn):
        for _ in range(48):
            rc = fn()
            if rc != _ENOSPACE:
                _check(rc, what)
                self._count()
                return
            self._grow()
        raise MemoryError("native executor: arenas keep overflowing")

    def forward(self, choice: np.ndarray, int_x: torch.Tensor, cat_x: Optional[torch.Tensor],
                emb_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = int_x.shape[0]
        logits = torch.empty(B, 1, dtype=torch.float32, device=self.dev)
        st = _lib.stream_ptr()
        _lib.LIB.ensure_workspace()
        self._retrying("nasrec_net_forward", lambda: _fn("nasrec_net_forward")(
            self.handle, choice.ctypes.data, int_x.data_ptr(), cat_x.data_ptr() if cat_x is not None else None,
            emb_rows.data_ptr() if emb_rows is not None else None, B, logits.data_ptr(), st))
        return logits

    def forward_multi(self, choices: Sequence[np.ndarray], int_x: torch.Tensor, cat_x: Optional[torch.Tensor],
                      emb_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Logits [C, B] of C candidates on one batch (nasrec_multi_subnet_eval): blocks that candidates share -- same
        choice, same upstream -- are computed once.  ``self.last_multi_stats`` = (blocks computed, blocks reused)."""
        B = int_x.shape[0]
        C_ = len(choices)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32).reshape(-1) for c in choices]))
        logits = torch.empty(C_, B, dtype=torch.float32, device=self.dev)
        stats = (C.c_int * 2)()
        st = _lib.stream_ptr()
        _lib.LIB.ensure_workspace()
        self._retrying("nasrec_multi_subnet_eval", lambda: _fn("nasrec_multi_subnet_eval")(
            self.handle, flat.ctypes.data, C_, int_x.data_ptr(), cat_x.data_ptr() if cat_x is not None else None,
            emb_rows.data_ptr() if emb_rows is not None else None, B, logits.data_ptr(), C.cast(stats, C.c_void_p), st))
        self.last_multi_stats = (int(stats[0]), int(stats[1]))
        return logits

    def forward_backward(self, choice: np.ndarray, int_x, cat_x, y, grad_scale: float = 1.0):
        B = int_x.shape[0]
        logits = torch.empty(B, 1, dtype=torch.float32, device=self.dev)
        loss = torch.empty(1, dtype=torch.float32, device=self.dev)
        st = _lib.stream_ptr()
        _lib.LIB.ensure_workspace()
        self._retrying("nasrec_net_forward_backward", lambda: _fn("nasrec_net_forward_backward")(
            self.handle, choice.ctypes.data, int_x.data_ptr(), cat_x.data_ptr(), y.data_ptr(), B, float(grad_scale),
            logits.data_ptr(), loss.data_ptr(), st))
        return logits, loss

    SEAL_CB = C.CFUNCTYPE(None, C.c_int64, C.c_int64)

    def set_seal_callback(self, fn):
        """fn(offset_bytes, nbytes) is called during forward_backward each time a block's parameter gradients are
        final (see nasrec_net_set_seal_callback); None switches it off."""
        self._seal_cb = self.SEAL_CB(fn) if fn is not None else None          # keep the thunk alive
        _check(_fn("nasrec_net_set_seal_callback")(self.handle, C.cast(self._seal_cb, C.c_void_p) if fn else None),
               "nasrec_net_set_seal_callback")

    def grad_bucket(self) -> torch.Tensor:
        """The step's dense parameter gradients as ONE flat fp32 view of the gradient arena
        (what data-parallel training all-reduces; padding between tensors is zero)."""
        ptr, n = C.c_void_p(), C.c_int64()
        _check(_fn("nasrec_net_grad_bucket")(self.handle, C.byref(ptr), C.byref(n)), "nasrec_net_grad_bucket")
        return self.pg[: n.value * 4].view(torch.float32)

    def sparse_raw(self, cat_x: torch.Tensor) -> Optional[torch.Tensor]:
        """d(loss)/d(embedding rows) [B,F,16] of the step as a view of the activation arena."""
        cp, gp = C.c_void_p(), C.c_void_p()
        _check(_fn("nasrec_net_sparse_raw")(self.handle, C.byref(cp), C.byref(gp)), "nasrec_net_sparse_raw")
        if not gp.value:
            return None
        B, F = cat_x.shape
        off = gp.value - self.act.data_ptr()
        return self.act[off: off + B * F * EMB * 4].view(torch.float32).view(B, F, EMB)

    def reserve(self, rows: int):
        """Rows the sparse reduction will see (world_size x B under data parallelism): forward_backward reserves the
        reduction's and the clip's scratch for them, so reduce / apply cannot overflow an arena afterwards."""
        if rows != getattr(self, "_reserved", None):
            _check(_fn("nasrec_net_set_reserve")(self.handle, int(rows)), "nasrec_net_set_reserve")
            self._reserved = rows

    def _after_step(self, what: str, rc: int):
        # These two run on the gradients forward_backward left in the arenas: growing the arenas here would free them.
        # forward_backward reserved their scratch, so an overflow means the caller under-announced the batch (reserve()).
        if rc == _ENOSPACE:
            raise MemoryError("%s ran out of arena after the step's gradients were produced; call NativeNet.reserve(rows) "
                              "with the size of the all-gathered batch before forward_backward" % what)
        _check(rc, what)
        self._count()

    def sparse_reduce(self, cat_all: Optional[torch.Tensor] = None, gout_all: Optional[torch.Tensor] = None):
        B_all = cat_all.shape[0] if cat_all is not None else 0
        self._after_step("nasrec_net_sparse_reduce", _fn("nasrec_net_sparse_reduce")(
            self.handle, cat_all.data_ptr() if cat_all is not None else None,
            gout_all.data_ptr() if gout_all is not None else None, B_all, _lib.stream_ptr()))

    def apply(self, lr: float, eps: float, clip: Optional[float]) -> torch.Tensor:
        norm = torch.empty(2, dtype=torch.float32, device=self.dev)
        self._after_step("nasrec_net_apply", _fn("nasrec_net_apply")(
            self.handle, float(lr), float(eps), float(clip) if clip is not None else 0.0, norm.data_ptr(),
            _lib.stream_ptr()))
        return norm


class NativeTrainer(FusedTrainer):
    """FusedTrainer whose steps run on the C++ executor when the model allows it (weight-sharing
    supernet, every parameter's Adagrad state pre-allocated); otherwise it is a FusedTrainer."""

    def __init__(self, model: SuperNet, lr: float, eps: float = 1e-2, clip: Optional[float] = 5.0,
                 overlap_wgrad: bool = True, defer_wgrad: bool = True):
        super().__init__(model, lr, eps, clip)
        # dense weight gradients queue up during backward and run as one batched launch at its end (nasrec_wgrad_flush);
        # off = one launch per operator, the Python engine's launch sequence (bit-identical to it)
        self.defer_wgrad = defer_wgrad
        self.net: Optional[NativeNet] = None
        self.fallback_reason: Optional[str] = None
        # weight-gradient GEMMs on a second stream, off the dY -> dX critical path (joined before the optimizer)
        self.overlap_wgrad = overlap_wgrad
        self.late_join = True
        self._side: Optional[torch.cuda.Stream] = None

    def _fork_on(self, net: "NativeNet"):
        if not self.overlap_wgrad:
            return
        if self._side is None:
            self._side = torch.cuda.Stream()
            # 2 = the final join moves into nasrec_net_apply (step() below always calls it next); a data-parallel owner
            # reads the gradient bucket in between and sets late_join = False
            _check(_fn("nasrec_net_set_overlap")(net.handle, 2 if self.late_join else 1), "nasrec_net_set_overlap")
        _fn("nasrec_set_side_stream")(self._side.cuda_stream)

    def _fork_off(self):
        if self.overlap_wgrad:
            _fn("nasrec_set_side_stream")(None)       # the Python engine's entry points must not fork

    def _native(self, int_x) -> Optional[NativeNet]:
        if self.net is None and self.fallback_reason is None:
            m = self.model
            if m._needs_materialize():
                m.materialize(int_x.shape[1])
            self.fallback_reason = NativeNet.unsupported_reason(m)
            if self.fallback_reason is None:
                try:
                    self.net = NativeNet(m, state_of=self._state_of)
                    _check(_fn("nasrec_net_set_defer_wgrad")(self.net.handle, 1 if self.defer_wgrad else 0),
                           "nasrec_net_set_defer_wgrad")
                except (NotImplementedError, ValueError) as e:      # a model shape the executor does not describe
                    self.fallback_reason = str(e)
        return self.net

    def _choice(self) -> np.ndarray:
        macro, micro = self.model._sample()
        return NativeNet.encode_choice(macro, micro)

    def step(self, int_x, cat_x, y, lr: Optional[float] = None):
        net = self._native(int_x)
        if net is None:
            return super().step(int_x, cat_x, y, lr)
        net.refresh()
        cat = cat_x if cat_x.dtype == torch.int64 else cat_x.long()
        with _lib.pin_stream():
            self._fork_on(net)
            try:
                logits, loss = net.forward_backward(self._choice(), int_x.contiguous(), cat.contiguous(),
                                                    y.contiguous())
            finally:
                self._fork_off()
            net.sparse_reduce()
            norm = net.apply(self.lr if lr is None else lr, self.eps, self.clip)
        self.last_total_norm = norm[0:1]
        return logits, loss
