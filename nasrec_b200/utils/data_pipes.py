"""Raw batch -> model inputs on the device.

Mirror of the transform half of ``nasrec/utils/data_pipes.py`` (:135-252): the reference
parses every categorical value with a Python ``int(v, 16)`` and builds one small tensor per
column; here the raw columns are packed once (numpy fixed-width bytes), copied to the GPU and
turned into ``(int_x [B,nd] f32, cat_x [B,F] i64, y [B,1] f32)`` by ONE kernel
(``nasrec_input_transform``).  Reading TSV files / torchdata pipes is out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .config import NUM_EMBEDDINGS_AVAZU, NUM_EMBEDDINGS_CRITEO, NUM_EMBEDDINGS_KDD

DEFAULT_LABEL_NAME = "label"
MAX_HEX_WIDTH = 15


def pack_hex_columns(cols: Sequence[Sequence[str]]) -> np.ndarray:
    """[F][B] hex strings -> uint8 [F, B, W] NUL-padded (W = longest field, >= 1)."""
    packed = [np.asarray(c, dtype="S") if len(c) else np.zeros(0, dtype="S1") for c in cols]
    width = max([1] + [p.dtype.itemsize for p in packed])
    if width > MAX_HEX_WIDTH:
        raise ValueError("categorical field of %d hex digits exceeds int64 (max %d)" % (width, MAX_HEX_WIDTH))
    B = len(cols[0]) if len(cols) else 0
    out = np.zeros((len(cols), B, width), dtype=np.uint8)
    for f, p in enumerate(packed):
        if len(p) != B:
            raise ValueError("categorical column %d has %d rows, expected %d" % (f, len(p), B))
        out[f, :, :p.dtype.itemsize] = p.view(np.uint8).reshape(B, p.dtype.itemsize)
    return out


class InputTransform:
    """Callable with the reference's ``VanillaTransform*`` contract: ``batch`` maps ``int_i`` to
    a 1-D numeric tensor/array, ``cat_i`` to a list of hex strings ('' = missing), ``label`` to
    a 1-D tensor; returns device tensors ``(int_x, cat_x, y)``."""

    def __init__(self, num_embeddings: Sequence[int], num_dense: int, zero_dense: bool = False,
                 device: Optional[torch.device] = None):
        self.num_embeddings = [int(n) for n in num_embeddings]
        self.num_dense = num_dense
        self.zero_dense = zero_dense
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._rows = None
        self._err = None

    def _device_state(self):
        if self._rows is None:
            self._rows = torch.tensor(self.num_embeddings, dtype=torch.int64, device=self.device)
            self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
        return self._rows, self._err

    def transform_columns(self, dense_cols, hex_cols: Sequence[Sequence[str]], label=None, check: bool = True):
        """dense_cols: [nd][B] numbers (ignored when zero_dense); hex_cols: [F][B] strings."""
        F = len(self.num_embeddings)
        if len(hex_cols) != F:
            raise ValueError("expected %d categorical columns, got %d" % (F, len(hex_cols)))
        hexb = pack_hex_columns(hex_cols)                               # [F, B, W]
        B, W = hexb.shape[1], hexb.shape[2]
        dev = self.device
        rows, err = self._device_state()
        hex_d = torch.from_numpy(hexb).to(dev, non_blocking=True)
        nd = 0 if self.zero_dense else self.num_dense
        if nd:
            raw = np.stack([np.asarray(c, dtype=np.float32) for c in dense_cols])      # [nd, B]
            if raw.shape != (nd, B):
                raise ValueError("dense columns have shape %s, expected %s" % (raw.shape, (nd, B)))
            raw_d = torch.from_numpy(raw).to(dev, non_blocking=True)
            int_x = torch.empty(B, nd, dtype=torch.float32, device=dev)
        else:
            raw_d = None
            int_x = torch.zeros(B, self.num_dense, dtype=torch.float32, device=dev)   # data_pipes.py:181
        cat_x = torch.empty(B, F, dtype=torch.int64, device=dev)
        _lib.call("nasrec_input_transform", raw_d.data_ptr() if nd else None, 1, B, nd, hex_d.data_ptr(), W, B * W, W, F,
                  rows.data_ptr(), B, int_x.data_ptr() if nd else None, cat_x.data_ptr(), err.data_ptr())
        if check and int(err.item()):
            err.zero_()
            raise ValueError("non-hexadecimal byte in a categorical field")      # int(v, 16) raises in the reference
        y = None
        if label is not None:
            y = torch.as_tensor(np.asarray(label), dtype=torch.float32).reshape(-1, 1).to(dev, non_blocking=True)
        return int_x, cat_x, y

    def __call__(self, batch: Dict[str, object]):
        dense = [batch["int_%d" % c] for c in range(self.num_dense)] if not self.zero_dense else []
        dense = [d.cpu().numpy() if isinstance(d, torch.Tensor) else d for d in dense]
        cats = [batch["cat_%d" % f] for f in range(len(self.num_embeddings))]
        label = batch[DEFAULT_LABEL_NAME]
        label = label.cpu().numpy() if isinstance(label, torch.Tensor) else label
        return self.transform_columns(dense, cats, label)


_cached: Dict[str, InputTransform] = {}


def _get(name, ne, nd, zero):
    t = _cached.get(name)
    if t is None:
        t = _cached[name] = InputTransform(ne, nd, zero_dense=zero)
    return t


def VanillaTransformCriteo(batch):               # data_pipes.py:147-175
    return _get("criteo", NUM_EMBEDDINGS_CRITEO, 13, False)(batch)


def VanillaTransformAvazu(batch):                # data_pipes.py:190-213
    return _get("avazu", NUM_EMBEDDINGS_AVAZU, 1, True)(batch)


def VanillaTransformKDD(batch):                  # data_pipes.py:228-252
    return _get("kdd", NUM_EMBEDDINGS_KDD, 3, False)(batch)
