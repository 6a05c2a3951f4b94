"""Whole-step CUDA-graph capture for fixed-choice training.

A fixed subnet (main_train.py --net supernet-config, or a supernet pinned with
configure_choice + "fixed-path") launches the same ~300 small kernels every step; at
B=256 the step is launch-bound.  The fused step (forward tape, BCE, backward tape, clip,
Adagrad) contains no host synchronisation, so it is captured once and replayed.
"""
from __future__ import annotations

from typing import Optional

import torch

from .train_utils import FusedTrainer


class GraphedFusedTrainer:
    """Wraps a FusedTrainer whose model always draws the same choice."""

    def __init__(self, trainer: FusedTrainer, warmup_steps: int = 0, overlap_wgrad: bool = True):
        self.trainer = trainer
        self.overlap_wgrad = overlap_wgrad      # weight-gradient GEMMs become a parallel branch of the graph
        self.warmup_steps = warmup_steps
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._static = None
        self._out = None
        self._lr = None
        self.kernels_per_replay = 0             # kernels of the library captured in the graph (each replay launches them)

    def _strategy_is_static(self) -> bool:
        m = self.trainer.model
        return m._fixed or m._macro_path_sampling_strategy == "fixed-path"

    def step(self, int_x, cat_x, y, lr: Optional[float] = None):
        lr = self.trainer.lr if lr is None else lr
        if not self._strategy_is_static():
            raise RuntimeError("CUDA-graph replay needs a fixed choice (fixed model or 'fixed-path' strategy)")
        if self.graph is None or lr != self._lr or int_x.shape != self._static[0].shape:
            self._capture(int_x, cat_x, y, lr)
        else:
            self._static[0].copy_(int_x, non_blocking=True)
            self._static[1].copy_(cat_x, non_blocking=True)
            self._static[2].copy_(y, non_blocking=True)
            # the captured step keeps the weight planes in step by itself; a weight changed from outside since the last
            # replay (load_state_dict, an in-place edit) is noticed here and its planes are rebuilt before the replay
            self.trainer._planes()
        self.graph.replay()
        from .. import _lib
        _lib.LIB.launches += self.kernels_per_replay
        return self._out

    def _capture(self, int_x, cat_x, y, lr):
        self._static = (int_x.clone(), cat_x.clone().long(), y.clone())
        self._lr = lr
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.trainer.prime(*self._static)       # lazy host state; leaves the model untouched
            for _ in range(self.warmup_steps):      # optional real steps off the capture
                self.trainer.step(*self._static, lr=lr)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if self.overlap_wgrad:
            self.trainer.side_stream = torch.cuda.Stream()
        from .. import _lib
        before = _lib.LIB.launches
        try:
            with torch.cuda.graph(self.graph):
                self._out = self.trainer.step(*self._static, lr=lr)
        finally:
            self.trainer.side_stream = None
        self.kernels_per_replay = _lib.LIB.launches - before
