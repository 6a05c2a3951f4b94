"""Checkpoint / results bridge: same file formats as
``nasrec/utils/io_utils.py`` so that a reference ``supernet_checkpoint.pt`` loads into the
B200 model with ``strict=True`` and checkpoints written here load back into the reference.

Format (io_utils.py:59-79): ``torch.save({"model_state_dict": ..., ["optimizer_state_dict": ...]})``;
search results are plain pickles of record lists (searcher.py / eval_subnet_from_supernet.py).
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Any, Dict, Optional

import torch
import torch.nn as nn


def _need(path: Optional[str], what: str) -> str:
    assert path is not None, "%s should not be 'None'!" % what
    return path


def load_json(json_file_name: Optional[str] = None):
    with open(_need(json_file_name, "Json file name"), "r") as fp:
        return json.load(fp)


def dump_json(json_file_name: Optional[str], data: Any):
    with open(_need(json_file_name, "Json file name"), "w") as fp:
        json.dump(data, fp)


def create_dir(dir_name: Optional[str] = None):
    os.makedirs(_need(dir_name, "Directory name"), exist_ok=True)


def dump_pickle_data(dump_path: Optional[str], data: Any):
    with open(_need(dump_path, "Dump path"), "wb") as fp:
        pickle.dump(data, fp)


def load_pickle_data(load_path: Optional[str]):
    with open(_need(load_path, "Load path"), "rb") as fp:
        return pickle.load(fp)


def load_model_checkpoint(load_path: Optional[str]) -> Dict[str, Any]:
    """Always to host memory, like the reference (io_utils.py:59-64); the caller's
    ``model.load_state_dict(ckpt["model_state_dict"], strict=True)`` moves it to the GPU."""
    with open(_need(load_path, "Load model path"), "rb") as fh:
        return torch.load(fh, map_location=torch.device("cpu"))


def save_model_checkpoint(model: nn.Module, save_path: str, optimizer: Optional[Any] = None):
    """``optimizer`` may be a torch optimizer or a ``FusedTrainer`` (its Adagrad accumulators are
    written in ``torch.optim.Adagrad.state_dict()`` layout, see ``adagrad_state_dict``)."""
    checkpoint = {"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}}
    if optimizer is not None:
        sd = optimizer.state_dict() if hasattr(optimizer, "state_dict") else adagrad_state_dict(optimizer, model)
        checkpoint["optimizer_state_dict"] = sd
    with open(save_path, "wb") as fh:
        torch.save(checkpoint, fh)


def adagrad_state_dict(trainer, model: nn.Module, step: int = 0) -> Dict[str, Any]:
    """FusedTrainer accumulators -> the dict ``torch.optim.Adagrad(model.parameters(), lr, eps)``
    would save: per-parameter {"step", "sum"} keyed by position in ``model.parameters()``;
    parameters the fused step never touched have a zero accumulator (their gradient was zero)."""
    state = {}
    params = list(model.parameters())
    for i, p in enumerate(params):
        acc = trainer.state.get(id(p))
        state[i] = {"step": torch.tensor(float(step)),
                    "sum": (acc.detach().cpu() if acc is not None else torch.zeros_like(p, device="cpu"))}
    group = {"lr": trainer.lr, "lr_decay": 0, "eps": trainer.eps, "weight_decay": 0, "initial_accumulator_value": 0,
             "foreach": None, "maximize": False, "differentiable": False, "fused": None,
             "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def load_adagrad_state(trainer, model: nn.Module, state_dict: Dict[str, Any]):
    """Inverse of ``adagrad_state_dict``: seed a FusedTrainer from a torch Adagrad state dict."""
    params = list(model.parameters())
    for i, st in state_dict["state"].items():
        p = params[int(i)]
        trainer.state[id(p)] = st["sum"].to(device=p.device, dtype=p.dtype).clone()
    trainer._emb_ptrs = None
