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

import copy
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine as eng
from .engine import Seg, Tape, Var
from .supernet.modules import Run
from .supernet.supernet import SuperNet
from .utils.lr_schedule import finetune_lr_sequence

_MACRO_KEYS = ("dense_idx", "sparse_idx", "dense_left_idx", "dense_right_idx")
_MICRO_KEYS = ("active_nodes", "dense_in_dims", "sparse_in_dims", "dense_sparse_interact", "deep_fm")


def _plain(o):
    """numpy scalars/arrays -> builtin ints/lists (choices travel through pickle/JSON)."""
    if isinstance(o, dict):
        return {str(k): _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, np.ndarray):
        return [_plain(v) for v in o.tolist()]
    if isinstance(o, np.generic):
        return o.item()
    return o


def _draw_macro(b: int) -> Dict[str, List[int]]:
    """One block's connection draw shared by generate_random_choice and mutate_spec
    (tokenizer.py:199-227, 280-308): at most 4 dense and 4 sparse sources, one left/right pair."""
    n_dense = 1 + np.random.choice(min(4, b + 1))
    n_sparse = 1 + np.random.choice(min(4, b + 1))
    bi = np.random.choice(b + 1, 2)
    return {"dense_idx": np.random.choice(b + 1, n_dense, replace=False).reshape(-1).tolist(),
            "sparse_idx": np.random.choice(b + 1, n_sparse, replace=False).reshape(-1).tolist(),
            "dense_left_idx": bi[:1].reshape(-1).tolist(), "dense_right_idx": bi[1:].reshape(-1).tolist()}


def _draw_micro(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """One block's operator draw, rejecting the all-zero pair (tokenizer.py:246-258, 310-330)."""
    while True:
        micro = {"active_nodes": sorted([int(np.random.choice(cfg["dense_nodes"]))] +
                                        [int(np.random.choice(cfg["sparse_nodes"]))]),
                 "dense_in_dims": int(np.random.choice(cfg["dense_node_dims"])),
                 "sparse_in_dims": int(np.random.choice(cfg["sparse_node_dims"])),
                 "dense_sparse_interact": int(np.random.choice([0, 1])),
                 "deep_fm": int(np.random.choice([0, 1]))}
        if micro["active_nodes"] != cfg["zero_nodes"]:
            return micro


def generate_random_choice(num_blocks: int, ops_config: Dict[str, Any]) -> Dict[str, Any]:
    """RNG-order-exact restatement of Tokenizer.generate_random_choice
    (searcher/tokenizer.py:267-336): the candidate distribution of the EA's random phase."""
    np.random.choice(num_blocks)                      # tokenizer.py:272 draws and discards a block index
    choice = {"macro": [], "micro": []}
    for b in range(num_blocks):
        choice["macro"].append(_draw_macro(b))
        choice["micro"].append(_draw_micro(ops_config[b] if isinstance(ops_config, list) else ops_config))
    return choice


class Tokenizer:
    """searcher/tokenizer.py:30-336: choice -> integer token / hash string (the EA's duplicate
    filter and the ``hash_token`` field of results.pickle) and the mutation operator."""

    def __init__(self, num_blocks: int, ops_config: Any):
        self._num_blocks = num_blocks
        self._ops_config = ops_config

    def _cfg(self, b: int) -> Dict[str, Any]:
        return self._ops_config[b] if isinstance(self._ops_config, list) else self._ops_config

    def tokenize(self, choice: Dict[str, Any]) -> np.ndarray:
        """tokenizer.py:158-186: per block 4 x num_blocks connection bits; then per block
        num_nodes activity bits, the two width indices and two one-hot pairs."""
        nb = self._num_blocks
        enc: List[int] = []
        for mac in choice["macro"]:
            for key in _MACRO_KEYS:
                picked = set(int(v) for v in np.asarray(mac[key]).reshape(-1))
                enc += [1 if i in picked else 0 for i in range(nb)]
        for b, mic in enumerate(choice["micro"]):
            cfg = self._cfg(b)
            active = set(int(v) for v in np.asarray(mic["active_nodes"]).reshape(-1))
            enc += [1 if i in active else 0 for i in range(cfg["num_nodes"])]
            enc.append(list(cfg["dense_node_dims"]).index(int(mic["dense_in_dims"])))
            enc.append(list(cfg["sparse_node_dims"]).index(int(mic["sparse_in_dims"])))
            enc += [1, 0] if mic["dense_sparse_interact"] == 0 else [0, 1]
            enc += [1, 0] if mic["deep_fm"] == 0 else [0, 1]
        return np.asarray(enc, dtype=np.int64)

    def hash_token(self, token) -> str:               # tokenizer.py:188-190
        return "".join(str(int(x)) for x in token)

    def generate_random_choice(self) -> Dict[str, Any]:
        return generate_random_choice(self._num_blocks, self._ops_config)

    def mutate_spec(self, choice: Dict[str, Any]) -> Dict[str, Any]:
        """tokenizer.py:192-265: redraw ONE field of ONE block (a whole fresh draw is made and
        one key of it kept, so the RNG consumption matches the reference)."""
        b = int(np.random.choice(self._num_blocks))
        level = "macro" if np.random.random() > 0.5 else "micro"
        out = copy.deepcopy(choice)
        if level == "macro":
            fresh = _draw_macro(b)
            key = str(np.random.choice(list(_MACRO_KEYS)))
        else:
            fresh = _draw_micro(self._cfg(b))
            key = str(np.random.choice(list(_MICRO_KEYS)))
        out[level][b][key] = copy.deepcopy(fresh[key])
        return out


_metrics_ws: Dict[Tuple[int, int], torch.Tensor] = {}


def binary_metrics_device(logits: torch.Tensor, y: torch.Tensor) -> Tuple[float, float, float]:
    """(accuracy@0.5, ROC-AUC, log-loss) over concatenated predictions, as train_utils.py:158-178,
    computed by ``nasrec_binary_metrics`` (device radix sort on the fp32 sigmoid outputs + exact
    integer Mann-Whitney count with ties at 1/2 == sklearn.metrics.roc_auc_score); only three
    doubles cross to the host."""
    from . import _lib
    z = logits.reshape(-1).contiguous().float()
    t = y.reshape(-1).contiguous().float()
    n = z.numel()
    if t.numel() != n or n == 0:
        raise ValueError("logits and labels must be non-empty and of equal length")
    need = _lib.query("nasrec_binary_metrics_ws_bytes", n)
    key = (z.device.index or 0, 0)
    ws = _metrics_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _metrics_ws[key] = torch.empty(int(need * 1.25), dtype=torch.uint8, device=z.device)
    out = torch.empty(3, dtype=torch.float64, device=z.device)
    _lib.call("nasrec_binary_metrics", z.data_ptr(), t.data_ptr(), n, ws.data_ptr(), ws.numel(), out.data_ptr())
    acc, auc, loss = out.tolist()
    return float(acc), float(auc), float(loss)


class SubnetEvaluator:
    """Scores many candidates against shared, resident supernet weights."""

    def __init__(self, model: SuperNet, use_native: bool = True, group: int = 16):
        assert not model._fixed, "one-shot scoring needs the weight-sharing supernet"
        self.model = model
        # Gathered embedding rows of the evaluation batches, shared by the candidates of ONE score() call: the cache is
        # keyed by the batch's device pointer, so it must not outlive the call (the allocator may hand the address to
        # another batch, the caller may refill a staging buffer, training may move the tables in between).
        self._emb_cache: Dict[int, torch.Tensor] = {}
        self._in_score = False
        self.group = group                      # candidates per nasrec_multi_subnet_eval call
        self.multi_stats = [0, 0]               # blocks computed / reused by the batched path so far
        self.use_native = use_native            # C++ executor (nasrec_b200/native.py) when the model allows it
        self._net = None
        self._native_checked = False

    def _native(self):
        if self.use_native and not self._native_checked:
            from .native import NativeNet
            self._native_checked = True
            if NativeNet.unsupported_reason(self.model) is None:
                self._net = NativeNet(self.model, state_of=None, pgrad_bytes=1 << 20)
        return self._net

    @torch.no_grad()
    def logits(self, choice, int_x: torch.Tensor, cat_x: torch.Tensor) -> torch.Tensor:
        m = self.model
        if m._needs_materialize():
            m.materialize(int_x.shape[1])
        net = self._native()
        try:
            if net is not None:
                from .native import NativeNet
                net.refresh()
                rows = self._gathered(cat_x) if not any(e.weight.requires_grad for e in m._embedding) else None
                return net.forward(NativeNet.encode_choice(choice["macro"], choice["micro"]), int_x.contiguous(),
                                   cat_x if rows is None else None, emb_rows=rows)
            run = Run(Tape(False), emb_cache=self._emb_cache)
            out = m._run_network(run, Var(int_x), cat_x, choice["macro"], choice["micro"])
            return out.t
        finally:
            if not self._in_score:
                self._emb_cache.clear()

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
        self._in_score = True
        try:
            return self._score(choices, batches, use_cuda_graph)
        finally:
            self._in_score = False
            self._emb_cache.clear()

    def _score(self, choices, batches, use_cuda_graph):
        ys = torch.cat([b[2].reshape(-1) for b in batches])
        same_shape = all(b[0].shape == batches[0][0].shape for b in batches)
        res = []
        m = self.model
        if m._needs_materialize():
            m.materialize(batches[0][0].shape[1])
        net = self._native()
        frozen = not any(e.weight.requires_grad for e in m._embedding)
        if net is not None and not use_cuda_graph and frozen:
            # batched multi-subnet path (nasrec_multi_subnet_eval): `group` candidates per call and per batch; blocks the
            # candidates share (same choice, same upstream) are computed once
            from .native import NativeNet
            net.refresh()
            enc = [NativeNet.encode_choice(ch["macro"], ch["micro"]) for ch in choices]
            total = ys.numel()
            for g0 in range(0, len(choices), self.group):
                grp = enc[g0:g0 + self.group]
                outs = torch.empty(len(grp), total, dtype=torch.float32, device=ys.device)
                off = 0
                for b in batches:
                    cat = b[1] if b[1].dtype == torch.int64 else b[1].long()
                    lg = net.forward_multi(grp, b[0].contiguous(), None, emb_rows=self._gathered(cat.contiguous()))
                    self.multi_stats[0] += net.last_multi_stats[0]
                    self.multi_stats[1] += net.last_multi_stats[1]
                    outs[:, off:off + lg.shape[1]] = lg
                    off += lg.shape[1]
                for k in range(len(grp)):
                    acc, auc, loss = binary_metrics_device(outs[k], ys)
                    res.append({"test_acc": acc, "test_auroc": auc, "test_loss": loss})
            return res
        if not (use_cuda_graph and same_shape and len(batches) > 2):
            for ch in choices:
                outs = [self.logits(ch, b[0], b[1]).reshape(-1) for b in batches]
                acc, auc, loss = binary_metrics_device(torch.cat(outs), ys)
                res.append({"test_acc": acc, "test_auroc": auc, "test_loss": loss})
            return res
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

    def finetune_and_score(self, choice: Dict[str, Any],
                           train_batches: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                           eval_batches: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                           lr: float = 0.04, eps: float = 1e-2, clip: Optional[float] = 5.0,
                           trunk_samples: int = 8192, restore: bool = True) -> Dict[str, Any]:
        """The reference's per-candidate recipe (eval_subnet_from_supernet.py:71-207 with
        --finetune_whole_supernet 0): Adagrad(lr, eps) + clip on ``_final`` only, one step per
        training batch under the cosine schedule with warm-up = steps // 10
        (lr_schedule.finetune_lr_sequence), then log-loss / AUC / accuracy on the evaluation batches.

        Everything below ``_final`` is frozen, so the trunk is a pure function of the batch: it is run
        once over groups of up to ``trunk_samples`` concatenated training samples, and the optimizer
        steps then walk the cached penultimate features batch by batch (same batches, same order, same
        update as the reference's loop; only the redundant trunk launches are gone).  ``restore``
        puts the supernet's ``_final`` back afterwards so candidates stay independent."""
        from .utils.train_utils import FusedTrainer
        m = self.model
        if m._needs_materialize():
            m.materialize(train_batches[0][0].shape[1] if train_batches else eval_batches[0][0].shape[1])
        W, b = m._final.weight, m._final.bias
        saved = (W.detach().clone(), b.detach().clone())
        macro, micro = choice["macro"], choice["micro"]
        n = len(train_batches)
        lrs = finetune_lr_sequence(n, lr) if n else []
        trainer = FusedTrainer(m, lr, eps, clip)
        ok = {id(W), id(b)}
        losses: List[torch.Tensor] = []
        i = 0
        with _lib_pin():
            while i < n:
                j, tot = i, 0
                while j < n and (j == i or tot + train_batches[j][0].shape[0] <= trunk_samples):
                    tot += train_batches[j][0].shape[0]
                    j += 1
                group = train_batches[i:j]
                int_x = group[0][0] if j - i == 1 else torch.cat([g[0] for g in group])
                cat_x = group[0][1] if j - i == 1 else torch.cat([g[1] for g in group])
                cat_x = cat_x if cat_x.dtype == torch.int64 else cat_x.long()
                with torch.no_grad():
                    segs = m._run_trunk(Run(Tape(False)), Var(int_x.contiguous()), cat_x.contiguous(), macro, micro)
                row = 0
                for k in range(i, j):
                    Bk = train_batches[k][0].shape[0]
                    tape = Tape(True)
                    run = Run(tape, grad_ok=ok)
                    sl = [Seg(sg.v, sg.off + row * sg.ld, sg.ld, sg.width, sg.w_off) for sg in segs]
                    logits = m._run_head(run, sl, Bk)
                    loss, dl = eng.bce_with_logits(logits.t, train_batches[k][2])
                    logits.g = dl
                    tape.backward()
                    trainer.apply(run, None, lrs[k])
                    losses.append(loss)
                    row += Bk
                i = j
        res = self.score([choice], eval_batches)[0]
        res["train_loss"] = [float(v) for v in torch.cat([l.reshape(1) for l in losses]).tolist()] if losses else []
        res["choice"] = choice
        if restore:
            with torch.no_grad():
                W.copy_(saved[0])
                b.copy_(saved[1])
        else:
            res["final_weight"], res["final_bias"] = W.detach().clone(), b.detach().clone()
        return res

    def release(self):
        self._emb_cache.clear()


def _lib_pin():
    from . import _lib
    return _lib.pin_stream()


def draw_fixed_path_candidate(model: SuperNet) -> Dict[str, Any]:
    """How the reference's random phase obtains a candidate: a fresh supernet switched to
    'fixed-path' draws its subnet on the first forward (searcher_utils.py:57-70 with choice=None,
    eval_subnet_from_supernet.py:103; samplers supernet.py:772-812, 1305-1313)."""
    model.macro_last_choice = None
    model._fixed_path_called = False
    for blk in model._blocks:
        blk.micro_last_choice = None
        blk._fixed_path_called = False
    model.configure_path_sampling_strategy("fixed-path")
    model._sample()
    return _plain(model.choice)


class Searcher:
    """In-process, GPU-resident restatement of searcher/searcher.py:26-295.

    The reference evaluates each candidate in a fresh OS process (one per GPU id) that rebuilds the
    supernet, warms it up and reloads the checkpoint; here one resident supernet per rank scores a
    contiguous shard of every batch of candidates (candidates are independent units, no
    data-path collective) and the per-candidate records are gathered to every rank, so all ranks
    take identical EA decisions from identical numpy RNG streams.  Records keep the results.pickle
    schema: {"choice", "test_acc", "test_auroc", "test_loss", "hash_token"}."""

    CRITERIA = ("test_loss", "test_acc", "test_auroc")

    def __init__(self, evaluator: SubnetEvaluator, tokenizer: Tokenizer, train_batches, eval_batches,
                 lr: float = 0.04, finetune: bool = True, group=None):
        self.evaluator = evaluator
        self._tokenizer = tokenizer
        self.train_batches = train_batches
        self.eval_batches = eval_batches
        self.lr = lr
        self.finetune = finetune
        self.group = group
        self.all_results: List[Dict[str, Any]] = []

    # -- evaluation of a batch of candidates, sharded over ranks
    def evaluate(self, choices: Sequence[Dict[str, Any]]) -> List[Dict[str, Any]]:
        from . import parallel
        rank, world = parallel.rank_world(self.group)
        lo, hi = parallel.shard_range(len(choices), world, rank)
        mine = []
        for ch in choices[lo:hi]:
            if self.finetune and self.train_batches:
                r = self.evaluator.finetune_and_score(ch, self.train_batches, self.eval_batches, lr=self.lr)
                r.pop("train_loss", None)
            else:
                r = self.evaluator.score([ch], self.eval_batches)[0]
            r["choice"] = _plain(ch)
            r["hash_token"] = self._tokenizer.hash_token(self._tokenizer.tokenize(ch))
            mine.append(r)
        return parallel.allgather_records(mine, self.group) if world > 1 else mine

    @classmethod
    def _sort_results_with_criterion(cls, results: Sequence[Dict[str, Any]], criterion: str = "test_loss"):
        """searcher.py:56-80: ascending loss, descending accuracy / AUROC (stable argsort)."""
        order = np.argsort(np.asarray([r[criterion] for r in results]).flatten())
        if criterion in ("test_acc", "test_auroc"):
            order = order[::-1]
        return [results[int(i)] for i in order]

    def random_search_from_supernet(self, budget: int = 200, criterion: str = "test_loss", top_k: int = 5,
                                    sorted: bool = True) -> List[Dict[str, Any]]:   # searcher.py:88-165
        assert top_k <= budget, "Should have 'top_k' smaller than 'budget'."
        assert criterion in self.CRITERIA, NotImplementedError("Criterion {} is not supported!".format(criterion))
        cands = [draw_fixed_path_candidate(self.evaluator.model) for _ in range(budget)]
        self.all_results = self.evaluate(cands)
        if sorted:
            return self._sort_results_with_criterion(self.all_results, criterion)[:top_k]
        return self.all_results[:top_k]

    def regularized_evolution_from_supernet(self, n_generations: int = 50, n_childs: int = 16,
                                            init_population: int = 100, sample_size: int = 5,
                                            criterion: str = "test_loss", top_k: int = 2) -> List[Dict[str, Any]]:
        """searcher.py:167-295: aging evolution.  Each generation samples ``sample_size`` members,
        mutates the best of them ``num_mutations`` times per child (more early, fewer late), rejects
        already-visited hashes, scores the ``n_childs`` children as ONE sharded batch, appends them and
        retires the ``n_childs`` oldest members.  Returns the per-generation top_k history."""
        assert criterion in self.CRITERIA, NotImplementedError("Criterion {} is not supported!".format(criterion))
        assert top_k <= sample_size, ValueError(
            "You must maintain more than 'top_k' children to append 'top_k' archs to history.")
        assert sample_size < init_population, ValueError(
            "Sample size must be no greater than the number of population ('init_population')!")
        population = list(self.random_search_from_supernet(budget=init_population, criterion=criterion,
                                                           top_k=init_population, sorted=False))
        history: List[Dict[str, Any]] = []
        visited = set()
        for gen in range(n_generations):
            picked = np.random.choice(len(population), sample_size, replace=False)     # searcher.py:82-86
            parent = self._sort_results_with_criterion([population[int(i)] for i in picked], criterion)[0]
            num_mutations = (n_generations - gen) // (max(20, n_generations // 5)) + 1
            children = []
            for _ in range(n_childs):
                child = copy.deepcopy(parent["choice"])
                while True:
                    for _ in range(num_mutations):
                        child = self._tokenizer.mutate_spec(child)
                    h = self._tokenizer.hash_token(self._tokenizer.tokenize(child))
                    if h not in visited:
                        visited.add(h)
                        break
                children.append(child)
            scored = self.evaluate(children)
            population += scored
            history += self._sort_results_with_criterion(scored, criterion)[:top_k]
            population = population[n_childs:]
        self.all_results = population
        return history
