"""One-shot subnet scoring for the EA search.

The reference pays, per candidate, an OS process, a fresh SuperNet (with full
embedding tables), a warm-up forward and a load_state_dict before it evaluates
(searcher/searcher_utils.py:57-126, eval_subnet_from_supernet.py:71-207).  Everything
but ``_final`` is frozen during scoring (supernet.py:850-853), so candidates that
see the same evaluation batches share the weights AND the embedding gather; only the
choice tables differ.  ``SubnetEvaluator`` keeps one resident supernet per GPU, gathers
each evaluation batch once, and streams candidates through it.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine as eng
from .engine import Tape, Var
from .supernet.modules import Run
from .supernet.supernet import SuperNet


def generate_random_choice(num_blocks: int, ops_config: Dict[str, Any]) -> Dict[str, Any]:
    """RNG-order-exact restatement of Tokenizer.generate_random_choice
    (searcher/tokenizer.py:267-336): the candidate distribution of the EA's random phase."""
    np.random.choice(num_blocks)                      # tokenizer.py:272 draws and discards a block index
    choice = {"macro": [], "micro": []}
    for b in range(num_blocks):
        n_dense = 1 + np.random.choice(min(4, b + 1))
        n_sparse = 1 + np.random.choice(min(4, b + 1))
        bi = np.random.choice(b + 1, 2)
        macro = {"dense_idx": np.random.choice(b + 1, n_dense, replace=False).reshape(-1).tolist(),
                 "sparse_idx": np.random.choice(b + 1, n_sparse, replace=False).reshape(-1).tolist(),
                 "dense_left_idx": bi[:1].reshape(-1).tolist(), "dense_right_idx": bi[1:].reshape(-1).tolist()}
        cfg = ops_config[b] if isinstance(ops_config, list) else ops_config
        while True:
            micro = {"active_nodes": sorted([int(np.random.choice(cfg["dense_nodes"]))] +
                                            [int(np.random.choice(cfg["sparse_nodes"]))]),
                     "dense_in_dims": int(np.random.choice(cfg["dense_node_dims"])),
                     "sparse_in_dims": int(np.random.choice(cfg["sparse_node_dims"])),
                     "dense_sparse_interact": int(np.random.choice([0, 1])),
                     "deep_fm": int(np.random.choice([0, 1]))}
            if micro["active_nodes"] != cfg["zero_nodes"]:
                break
        choice["macro"].append(macro)
        choice["micro"].append(micro)
    return choice


def binary_metrics_device(logits: torch.Tensor, y: torch.Tensor) -> Tuple[float, float, float]:
    """(accuracy@0.5, ROC-AUC, log-loss) over concatenated predictions, as
    train_utils.py:158-178.  Log-loss comes from the fused BCE kernel; AUC is the rank
    statistic with average ranks for ties (== sklearn.metrics.roc_auc_score)."""
    z = logits.reshape(-1).contiguous()
    t = y.reshape(-1).contiguous()
    loss, _ = eng.bce_with_logits(z, t, want_grad=False)
    n = z.numel()
    order = torch.argsort(z, stable=True)
    zs = z[order]
    ts = t[order]
    # average ranks of tied groups
    new = torch.ones(n, dtype=torch.bool, device=z.device)
    new[1:] = zs[1:] != zs[:-1]
    gid = torch.cumsum(new.to(torch.int64), 0) - 1
    pos = torch.arange(1, n + 1, dtype=torch.float64, device=z.device)
    ng = int(gid[-1].item()) + 1
    gsum = torch.zeros(ng, dtype=torch.float64, device=z.device).index_add_(0, gid, pos)
    gcnt = torch.zeros(ng, dtype=torch.float64, device=z.device).index_add_(0, gid, torch.ones_like(pos))
    ranks = (gsum / gcnt)[gid]
    npos = ts.double().sum()
    nneg = n - npos
    auc = (ranks[ts > 0.5].sum() - npos * (npos + 1) / 2) / (npos * nneg)
    acc = ((z > 0).float() == t).float().mean()      # sigmoid(z) > 0.5  <=>  z > 0
    return float(acc.item()), float(auc.item()), float(loss.item())


class SubnetEvaluator:
    """Scores many candidates against shared, resident supernet weights."""

    def __init__(self, model: SuperNet):
        assert not model._fixed, "one-shot scoring needs the weight-sharing supernet"
        self.model = model
        self._emb_cache: Dict[int, torch.Tensor] = {}

    @torch.no_grad()
    def logits(self, choice, int_x: torch.Tensor, cat_x: torch.Tensor) -> torch.Tensor:
        m = self.model
        if m._needs_materialize():
            m.materialize(int_x.shape[1])
        run = Run(Tape(False), emb_cache=self._emb_cache)
        out = m._run_network(run, Var(int_x), cat_x, choice["macro"], choice["micro"])
        return out.t

    @torch.no_grad()
    def _gathered(self, cat_x: torch.Tensor) -> torch.Tensor:
        """Embedding rows of one evaluation batch, gathered once and shared by every candidate."""
        hit = self._emb_cache.get(cat_x.data_ptr())
        if hit is None:
            run = Run(Tape(False), emb_cache=self._emb_cache)
            m = self.model
            eng.embedding(run.tape, m._tables, [run.pv(e.weight) for e in m._embedding], cat_x, cache=self._emb_cache)
            hit = self._emb_cache[cat_x.data_ptr()]
        return hit

    @torch.no_grad()
    def score(self, choices: Sequence[Dict[str, Any]], batches: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
              use_cuda_graph: bool = False) -> List[Dict[str, float]]:
        """Each candidate: logits on every batch -> log-loss, AUC, accuracy.  The embedding
        gather of a batch is done once and shared by all candidates (tables are frozen).  With
        use_cuda_graph the candidate's forward is captured on the first batch and replayed on
        the others (all batches must then share one shape), removing the per-launch host cost;
        capture + instantiation cost ~70 ms per candidate, so it only pays for many small batches
        (at 8192-sample batches scoring is device-bound and eager is faster)."""
        ys = torch.cat([b[2].reshape(-1) for b in batches])
        same_shape = all(b[0].shape == batches[0][0].shape for b in batches)
        res = []
        if not (use_cuda_graph and same_shape and len(batches) > 2):
            for ch in choices:
                outs = [self.logits(ch, b[0], b[1]).reshape(-1) for b in batches]
                acc, auc, loss = binary_metrics_device(torch.cat(outs), ys)
                res.append({"test_acc": acc, "test_auroc": auc, "test_loss": loss})
            return res
        m = self.model
        if m._needs_materialize():
            m.materialize(batches[0][0].shape[1])
        rows = [self._gathered(b[1]) for b in batches]          # shared across candidates
        B = batches[0][0].shape[0]
        s_int = batches[0][0].clone()
        s_cat = batches[0][1].clone()
        s_rows = rows[0].clone()
        out_all = torch.empty(len(batches), B, dtype=torch.float32, device=s_int.device)
        self.logits(choices[0], batches[0][0], batches[0][1])    # lazy kernel attributes outside capture
        for ch in choices:
            cache = {s_cat.data_ptr(): s_rows}
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run = Run(Tape(False), emb_cache=cache)
                out = m._run_network(run, Var(s_int), s_cat, ch["macro"], ch["micro"]).t
            for bi, b in enumerate(batches):
                s_int.copy_(b[0], non_blocking=True)
                s_rows.copy_(rows[bi], non_blocking=True)
                g.replay()
                out_all[bi].copy_(out.reshape(-1), non_blocking=True)
            acc, auc, loss = binary_metrics_device(out_all.reshape(-1), ys)
            res.append({"test_acc": acc, "test_auroc": auc, "test_loss": loss})
            del g
        return res

    def release(self):
        self._emb_cache.clear()
