"""ctypes binding of include/nasrec_b200.h (the C-ABI drop-in boundary).

There is deliberately NO fallback: if the CUDA library is missing or a call
fails, this module raises.  Nothing here (or anywhere in nasrec_b200) imports
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnasrec_b200.so")

MAX_SEGS = 16
ATTN_PARAMS = 1696

_f = C.c_void_p      # device pointers travel as integers
_i = C.c_int
_l = C.c_int64
_fl = C.c_float

# name -> (argtypes, kernels launched per call) ; restype is int unless noted
_SIGS = {
    "nasrec_emb_gather_fwd": ([_f, _f, _f, _f, _i, _i, _f, _f], 1),
    "nasrec_emb_grad_sort_reduce": ([_f, _f, _i, _i, _f, _f, _f, _f, _f, _f], 1),
    "nasrec_emb_grad_sort_reduce_checked": ([_f, _f, _f, _f, _i, _i, _f, _f, _f, _f, _f, _f], 1),
    "nasrec_emb_grad_sort_reduce_big": ([_f, _f, _f, _f, _i, _i, _f, _f, _f, _f, _f, _l, _f], 10),
    "nasrec_emb_grad_to_dense": ([_f, _f, _f, _f, _i, _i, _f], 1),
    "nasrec_emb_rowwise_adagrad": ([_f, _f, _f, _f, _f, _i, _i, _fl, _fl, _f, _f], 1),
    "nasrec_seg_linear_fwd": ([_f, _i, _f, _l, _i, _i, _f, _f, _l, _i, _f], 1),
    "nasrec_seg_linear_dgrad": ([_f, _l, _i, _f, _l, _i, _f, _i, _i, _i, _f], 1),
    "nasrec_seg_linear_wgrad": ([_f, _l, _i, _f, _i, _f, _l, _i, _i, _i, _f], 1),
    "nasrec_wgrad_flush": ([_f], 1),
    "nasrec_colsum": ([_f, _l, _i, _i, _f, _i, _f], 1),
    "nasrec_sproj_fwd": ([_f, _i, _f, _l, _i, _f, _f, _l, _i, _f], 1),
    "nasrec_sproj_dgrad": ([_f, _l, _i, _f, _l, _f, _i, _i, _i, _f], 1),
    "nasrec_sproj_wgrad": ([_f, _l, _i, _f, _i, _f, _l, _i, _i, _f, _f], 2),
    "nasrec_sproj_bias_grad": ([_f, _l, _i, _i, _f, _i, _f], 1),
    "nasrec_linear_ln_fwd": ([_f, _i, _f, _l, _i, _i, _f, _f, _f, _fl, _i, _i, _f, _f, _l, _f, _f, _i, _i, _f], 2),
    "nasrec_linear_ln_bwd": ([_f, _l, _i, _f, _i, _i, _f, _f, _f, _f, _i, _f, _f, _f, _i, _f, _l, _i, _f, _f, _f, _f,
                              _f, _f], 5),
    "nasrec_sproj_ln_fwd": ([_f, _i, _f, _l, _i, _f, _f, _f, _fl, _i, _i, _f, _f, _l, _f, _f, _i, _i, _f], 2),
    "nasrec_sproj_ln_bwd": ([_f, _l, _i, _f, _i, _i, _f, _f, _f, _f, _i, _f, _f, _f, _i, _f, _l, _f, _f, _f, _f, _f,
                             _f, _f], 6),
    "nasrec_ln_fwd": ([_f, _l, _i, _i, _f, _f, _fl, _i, _i, _f, _l, _f, _f, _i, _f], 1),
    "nasrec_ln_bwd": ([_f, _l, _i, _f, _l, _i, _i, _f, _f, _f, _f, _i, _f, _l, _f, _f, _i, _f], 2),
    "nasrec_ln3_fwd": ([_f, _l, _i, _i, _f, _f, _fl, _i, _i, _f, _l, _f, _f, _i, _f], 1),
    "nasrec_ln3_bwd": ([_f, _l, _i, _f, _l, _i, _i, _f, _f, _f, _f, _i, _f, _l, _f, _f, _i, _f], 2),
    "nasrec_act_fwd": ([_f, _l, _i, _i, _i, _f, _l, _i, _f], 1),
    "nasrec_act_bwd": ([_f, _l, _f, _l, _i, _i, _i, _f, _l, _f], 1),
    "nasrec_dot_tril_fwd": ([_f, _l, _f, _l, _i, _f, _l, _i, _f], 1),
    "nasrec_dot_tril_bwd": ([_f, _l, _f, _l, _f, _l, _i, _f, _l, _f, _l, _i, _f], 1),
    "nasrec_gate_fwd": ([_f, _l, _f, _i, _f, _l, _i, _f], 1),
    "nasrec_gate_bwd": ([_f, _l, _f, _l, _f, _f, _i, _f, _l, _i, _i, _f], 1),
    "nasrec_concat_segs": ([_f, _i, _f, _l, _i, _i, _f], 1),
    "nasrec_fm_fwd": ([_f, _l, _i, _f, _i, _f], 1),
    "nasrec_fm_bwd": ([_f, _f, _l, _i, _f, _l, _f, _l, _i, _f], 1),
    "nasrec_attn_fwd": ([_f, _l, _i, _i, _f, _f, _l, _i, _f], 1),
    "nasrec_attn_bwd": ([_f, _l, _f, _l, _i, _i, _f, _f, _l, _f, _i, _f, _i, _f], 2),
    "nasrec_bce_fwd_bwd": ([_f, _f, _i, _fl, _f, _f, _f], 1),
    "nasrec_grad_norm_clip": ([_f, _f, _i, _f, _i, _fl, _f, _f, _f], 2),
    "nasrec_adagrad_multi": ([_f, _f, _f, _f, _i, _fl, _fl, _f, _f], 1),
    "nasrec_adagrad_multi_planes": ([_f, _f, _f, _f, _i, _fl, _fl, _f, _f, _f, _f, _f, _f, _f], 1),
    "nasrec_planes_refresh": ([_f, _l, _i, _i, _i, _f, _f, _l, _f], 1),
    "nasrec_binary_metrics": ([_f, _f, _l, _f, _l, _f, _f], 6),     # prepare, sort (3 passes), scan, pairs, final
    "nasrec_input_transform": ([_f, _l, _l, _i, _f, _l, _l, _i, _i, _f, _l, _f, _f, _f, _f], 1),
}
_SIGS_I64 = {
    "nasrec_sproj_wgrad_ws_floats": [_i, _l, _i],
    "nasrec_attn_bwd_ws_floats": [_i],
    "nasrec_sumsq_ws_floats": [_f, _i],
    "nasrec_emb_grad_sort_reduce_big_ws_bytes": [_i, _i],
    "nasrec_binary_metrics_ws_bytes": [_l],
}
_SIGS_I64["nasrec_tensor_map_stats"] = [_i]
_SIGS_I64["nasrec_host_prof"] = [_i]
_SIGS_I64["nasrec_wgrad_pending"] = []
EXPORTS = ["nasrec_version", "nasrec_set_gemm_mode", "nasrec_get_gemm_mode", "nasrec_set_workspace",
           "nasrec_set_side_stream", "nasrec_side_join", "nasrec_set_gemm_tma", "nasrec_set_small_k", "nasrec_wgrad_defer", "nasrec_gemm_plan",
           "nasrec_set_weight_planes", "nasrec_gemm_prof"] + list(_SIGS) + list(_SIGS_I64)


class _Lib:
    def __init__(self):
        self.cdll = None
        self.launches = 0          # kernels launched through this binding (bench.py's gpu_launches)
        self.workspace = None
        self.fn = {}
        # bumped whenever dense weights are rewritten through raw pointers by a call that does NOT keep the
        # pre-split weight planes in step (nasrec_b200/planes.py refreshes them on a mismatch)
        self.weights_epoch = 0

    def load(self):
        if self.cdll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "nasrec_b200: CUDA library %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nasrec_b200/csrc/build.sh). There is no CPU fallback." % LIB_PATH)
        self.cdll = C.CDLL(LIB_PATH)
        for name, (argtypes, _n) in _SIGS.items():
            fn = getattr(self.cdll, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
            self.fn[name] = fn
        for name, argtypes in _SIGS_I64.items():
            fn = getattr(self.cdll, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int64
            self.fn[name] = fn
        self.cdll.nasrec_set_gemm_mode.argtypes = [C.c_int]
        self.cdll.nasrec_set_gemm_mode.restype = C.c_int
        self.cdll.nasrec_set_workspace.argtypes = [C.c_void_p, C.c_int64]
        self.cdll.nasrec_set_workspace.restype = C.c_int
        self.cdll.nasrec_get_gemm_mode.argtypes = []
        self.cdll.nasrec_get_gemm_mode.restype = C.c_int
        self.cdll.nasrec_set_gemm_tma.argtypes = [C.c_int]
        self.cdll.nasrec_set_gemm_tma.restype = C.c_int
        self.cdll.nasrec_gemm_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.cdll.nasrec_gemm_plan.restype = C.c_int
        self.cdll.nasrec_wgrad_defer.argtypes = [C.c_int]
        self.cdll.nasrec_wgrad_defer.restype = C.c_int
        self.cdll.nasrec_set_small_k.argtypes = [C.c_int]
        self.cdll.nasrec_set_small_k.restype = C.c_int
        self.cdll.nasrec_set_weight_planes.argtypes = [_f, _f, _f, _l, _i, _i, _i]
        self.cdll.nasrec_set_weight_planes.restype = C.c_int
        tma = os.environ.get("NASREC_GEMM_TMA")
        if tma is not None:
            self.cdll.nasrec_set_gemm_tma(int(tma))
        mode = os.environ.get("NASREC_GEMM_MODE")
        if mode is not None:
            if self.cdll.nasrec_set_gemm_mode(int(mode)) != 0:
                raise ValueError("NASREC_GEMM_MODE=%s is not one of 0..4" % mode)
        self.cdll.nasrec_version.argtypes = [C.POINTER(C.c_int)]
        self.cdll.nasrec_version.restype = C.c_int
        return self

    def set_gemm_mode(self, mode: int):
        """0 = fp32 FFMA, 3/4 = tcgen05 3xTF32 / 4xTF32, 2 = bf16 operands (one product, fp32 accumulate), 1 = single-pass
        tf32 (diagnostics)."""
        self.load()
        before = int(self.cdll.nasrec_get_gemm_mode())
        if self.cdll.nasrec_set_gemm_mode(int(mode)) != 0:
            raise ValueError("unsupported GEMM mode %r" % (mode,))
        if (before == 2) != (int(mode) == 2):
            self.weights_epoch += 1          # the weight planes hold a different rounding now: rebuild them

    def set_gemm_tma(self, on: bool):
        """TMA-fed operand path of the tensor-core GEMM on/off (both paths agree bit for bit)."""
        self.load()
        self.cdll.nasrec_set_gemm_tma(1 if on else 0)

    def gemm_plan(self, kind: int, M: int, N: int, K: int, nprob: int = 1) -> Tuple[int, int]:
        """(tile width, split-K factor) the planner picks for a launch of `nprob` [M x N x K] problems (host only)."""
        self.load()
        bn, ns = C.c_int(0), C.c_int(0)
        if self.cdll.nasrec_gemm_plan(kind, M, N, K, nprob, C.byref(bn), C.byref(ns)) != 0:
            raise ValueError("nasrec_gemm_plan rejected its arguments")
        return bn.value, ns.value

    def set_small_k(self, k: int) -> int:
        """Largest contraction length served by the CUDA-core kernel instead of the tensor-core pipeline (0 = never)."""
        self.load()
        return int(self.cdll.nasrec_set_small_k(int(k)))

    def set_weight_planes(self, W: int, hi: int, lo: int, ldp: int, rows: int, cols: int, first: int = 0):
        self.load()
        if self.cdll.nasrec_set_weight_planes(W, hi, lo, ldp, rows, cols, first) != 0:
            raise ValueError("nasrec_set_weight_planes rejected its arguments")

    def gemm_prof_start(self):
        self.load()
        self.cdll.nasrec_gemm_prof.argtypes = [C.c_int, C.c_void_p]
        self.cdll.nasrec_gemm_prof(1, None)

    def gemm_prof_stop(self):
        """(total ms, launches, algorithmic flops) of the GEMM launches since gemm_prof_start; synchronises."""
        out = (C.c_double * 3)()
        self.cdll.nasrec_gemm_prof.argtypes = [C.c_int, C.c_void_p]
        self.cdll.nasrec_gemm_prof(2, C.cast(out, C.c_void_p))
        return float(out[0]), int(out[1]), float(out[2])

    def ensure_workspace(self, nfloats: int = 16 * 1024 * 1024):
        """Attach a split-K scratch buffer on the current CUDA device (kept alive here)."""
        if self.workspace is None:
            self.load()
            self.workspace = torch.empty(nfloats, dtype=torch.float32, device="cuda")
            if self.cdll.nasrec_set_workspace(self.workspace.data_ptr(), nfloats) != 0:
                raise RuntimeError("nasrec_set_workspace failed")

    def gemm_mode(self) -> int:
        self.load()
        return int(self.cdll.nasrec_get_gemm_mode())

    def version(self) -> Tuple[int, int]:
        self.load()
        sm = C.c_int(0)
        v = self.cdll.nasrec_version(C.byref(sm))
        return v, sm.value


LIB = _Lib()


_STREAM = [None]      # set by pin_stream() for the duration of a step; None -> ask torch on every call


def stream_ptr() -> int:
    s = _STREAM[0]
    return torch.cuda.current_stream().cuda_stream if s is None else s


class pin_stream:
    """Context manager: resolve torch's current stream once for a whole launch sequence."""

    def __enter__(self):
        self.prev = _STREAM[0]
        _STREAM[0] = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        _STREAM[0] = self.prev


def call(name: str, *args):
    """Invoke a C-ABI entry point on torch's current stream; raise on any error."""
    lib = LIB.load()
    if lib.workspace is None:
        lib.ensure_workspace()
    rc = lib.fn[name](*args, stream_ptr())
    if rc != 0:
        if rc > 0:
            raise RuntimeError("%s failed: CUDA error %d" % (name, rc))
        raise ValueError("%s rejected its arguments (code %d)" % (name, rc))
    lib.launches += _SIGS[name][1]
    if name == "nasrec_adagrad_multi":
        lib.weights_epoch += 1


def query(name: str, *args) -> int:
    return int(LIB.load().fn[name](*args))


def segs(items: Sequence[Tuple[int, int, int, int]]):
    """Pack (device_ptr, ld, width, w_off) tuples as a nasrec_seg_t[] (4 x int64 each)."""
    n = len(items)
    if n == 0 or n > MAX_SEGS:
        raise ValueError("segment list must hold 1..%d entries, got %d" % (MAX_SEGS, n))
    flat: List[int] = []
    for it in items:
        flat.extend(it)
    return (C.c_int64 * (4 * n))(*flat), n


def i32_array(vals: Sequence[int]):
    return (C.c_int * len(vals))(*vals)


def i64_array(vals: Sequence[int]):
    return (C.c_int64 * len(vals))(*vals)


def ptr_array(vals: Sequence[int]):
    return (C.c_void_p * len(vals))(*vals)
