"""Pre-split weight planes for the TMA-fed tensor-core GEMM (nasrec_b200/csrc/gemm_tma.cuh).

The model's ``nn.Linear`` weights keep the reference's state-dict layout ``[out, D_i]`` (``D_i = nd + 1024*i``, e.g.
1037 floats per row), which no TMA tensor map can describe (row strides must be multiples of 16 bytes).  For every
2-D weight the cache keeps two derived device tensors ``hi = rn_tf32(W)`` and ``lo = W - hi`` of shape
``[out, ldp]``.  A TMA box must also *start* on a 16-byte boundary, and the reference layout puts the second source
of a concat at column ``nd`` (13) or ``F`` (26): the planes shift every column ``>= first`` right by
``(4 - first % 4) % 4`` zero pad columns (``first`` = width of the first concat source), so every segment start is a
multiple of 4 floats.  They are what forward and dgrad GEMMs fetch; the weights themselves stay the single source of
truth:

* the native optimizer step rewrites the planes of every tensor it updates in the same kernel
  (``nasrec_adagrad_multi_planes``);
* anything else that changes a weight is noticed here -- in-place torch ops through ``Tensor._version``, storage
  moves through ``data_ptr()``, raw-pointer updates by the Python engine's fused trainer through
  ``_lib.LIB.weights_epoch`` -- and the affected planes are rebuilt by ``nasrec_planes_refresh`` before the next use.

One cache per model (``PlaneCache.of(model)``), shared by every executor built over it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


class PlaneCache:
    ATTR = "_nasrec_plane_cache"

    def __init__(self, first_widths=(0,)):
        # candidate widths of the first concat source (num_dense_features, sparse_input_size): a weight whose column
        # count is  first + k * 1024  (dense inputs) or  first + k * 72  (sparse-axis inputs, 64 + 8 merger rows)
        # gets its columns >= first shifted; a wrong guess only sends that weight's GEMMs down the LDG path
        self.first_widths = tuple(first_widths)
        self.hi: Dict[int, torch.Tensor] = {}
        self.lo: Dict[int, torch.Tensor] = {}
        self.seen: Dict[int, Tuple[int, int]] = {}       # id(param) -> (data_ptr, _version) the planes were built from
        self.epoch = -1
        self.generation = 0                              # bumped when plane storage is (re)allocated

    @classmethod
    def of(cls, model) -> "PlaneCache":
        c = model.__dict__.get(cls.ATTR)
        if c is None:
            c = cls((int(getattr(model, "_num_dense_features", 0) or 0), int(getattr(model, "_sparse_input_size", 0) or 0)))
            model.__dict__[cls.ATTR] = c
        return c

    def first(self, p: torch.Tensor) -> int:
        cols = p.shape[1]
        nd, F = (self.first_widths + (0, 0))[:2]
        if nd % 4 and cols > nd and (cols - nd) % 1024 == 0:
            return nd
        if F % 4 and cols > F and (cols - F) % 72 == 0:
            return F
        return 0

    @staticmethod
    def wanted(p: torch.Tensor) -> bool:
        return p.dim() == 2 and p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()

    def ldp(self, p: torch.Tensor) -> int:
        f = self.first(p)
        return (p.shape[1] + ((4 - f % 4) % 4) + 3) & ~3

    def _rebuild(self, p: torch.Tensor):
        k = id(p)
        rows, cols = p.shape
        ldp = self.ldp(p)
        hi = self.hi.get(k)
        if hi is None or hi.shape != (rows, ldp) or hi.device != p.device:
            self.hi[k] = torch.zeros(rows, ldp, dtype=torch.float32, device=p.device)       # pad columns stay zero
            self.lo[k] = torch.zeros(rows, ldp, dtype=torch.float32, device=p.device)
            self.generation += 1
        _lib.call("nasrec_planes_refresh", p.data_ptr(), cols, rows, cols, self.first(p), self.hi[k].data_ptr(),
                  self.lo[k].data_ptr(), ldp)
        self.seen[k] = (p.data_ptr(), p._version)

    def sync(self, params: List[torch.Tensor]) -> int:
        """Bring the planes of ``params`` (the 2-D ones) in step with the weights; returns how many were rebuilt."""
        everything = self.epoch != _lib.LIB.weights_epoch
        n = 0
        for p in params:
            if not self.wanted(p):
                continue
            if everything or self.seen.get(id(p)) != (p.data_ptr(), p._version):
                self._rebuild(p)
                n += 1
        self.epoch = _lib.LIB.weights_epoch
        return n

    def planes(self, p: torch.Tensor) -> Optional[Tuple[torch.Tensor, torch.Tensor, int, int]]:
        """(hi, lo, ldp, first) of a weight, or None."""
        k = id(p)
        if k not in self.hi:
            return None
        return self.hi[k], self.lo[k], self.ldp(p), self.first(p)
