"""B200-native mirror of ``nasrec/supernet/supernet.py``: ``SuperNet`` /
``SuperNetBlock`` with the reference's constructor kwargs, attributes, samplers
(call-for-call identical use of numpy's global legacy RNG), lazy life cycle and
state-dict layout -- executed as fused sm_100a kernels over *compact* tensors.

Where the reference builds four zero-padded ``torch.cat``s per block and runs
full-width masked modules (supernet.py:530-588), this implementation keeps
every block output compact ([B,d] / [B,s(+8),16]) and hands each node a segment
list, so only the sampled subnet's support is ever read or multiplied, and
blocks that cannot reach the logit are skipped (bit-identical logits/grads:
the skipped work is exactly zero or unreachable in the reference's autograd
graph too).

Citations: NasRec repo root.
"""
from __future__ import annotations

import copy
from itertools import combinations
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import engine as eng
from ..engine import Seg, Tape, Var
from ..utils.config import NUM_EMBEDDINGS_CRITEO
from .modules import (CleverMaskGenerator, CleverZeroTensorGenerator, DotProduct, ElasticLinear, ElasticLinear3D,
                      FactorizationMachine3D, Run, SigmoidGating, Sum, Transformer, Zeros2D, Zeros3D, _materialize,
                      run_with_autograd)
from .utils import (anypath_choice_fn, assert_valid_ops_config, pick, pick_index, pick_with_replacement,
                    pick_without_replacement)

EMB = 16

# supernet.py:53-113 -- node factories, same keys and call signatures
_node_choices = {
    "linear-2d": lambda use_layernorm, max_dims_or_dims, activation, fixed: ElasticLinear(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, activation=activation, fixed=fixed),
    "zeros-2d": lambda use_layernorm, max_dims_or_dims, activation, fixed: Zeros2D(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, activation=activation, fixed=fixed),
    "sigmoid-gating": lambda use_layernorm, max_dims_or_dims, activation, fixed: SigmoidGating(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, activation=activation, fixed=fixed),
    "sum": lambda use_layernorm, max_dims_or_dims, activation, fixed: Sum(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, activation=activation, fixed=fixed),
    "dot-product": lambda use_layernorm, max_dims_or_dims, embedding_dim, fixed: DotProduct(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, embedding_dim=embedding_dim, fixed=fixed),
    "zeros-3d": lambda use_layernorm, max_dims_or_dims, activation, embedding_dim, fixed: Zeros3D(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, fixed=fixed, embedding_dim=embedding_dim,
        activation=activation),
    "transformer": lambda use_layernorm, max_dims_or_dims, activation, embedding_dim, fixed: Transformer(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, fixed=fixed, embedding_dim=embedding_dim,
        activation=activation),
    "linear-3d": lambda use_layernorm, max_dims_or_dims, activation, embedding_dim, fixed: ElasticLinear3D(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, activation=activation,
        embedding_dim=embedding_dim, fixed=fixed),
    "fm": lambda use_layernorm, max_dims_or_dims, fixed: FactorizationMachine3D(
        use_layernorm=use_layernorm, max_dims_or_dims=max_dims_or_dims, fixed=fixed),
}

_dense_unary_nodes = ["linear-2d", "zeros-2d"]          # supernet.py:116
_dense_binary_nodes = ["sum", "sigmoid-gating"]          # supernet.py:118
_dense_sparse_nodes = ["dot-product"]                    # supernet.py:120
_sparse_nodes = ["zeros-3d", "transformer", "linear-3d"]  # supernet.py:122

# supernet.py:134-178
ops_config_lib = {
    "xlarge": {
        "num_nodes": 6,
        "node_names": ["linear-2d", "dot-product", "sigmoid-gating", "sum", "transformer", "linear-3d"],
        "dense_node_dims": [16, 32, 64, 128, 256, 512, 768, 1024],
        "sparse_node_dims": [16, 32, 48, 64],
        "dense_nodes": [0, 1, 2, 3],
        "sparse_nodes": [4, 5],
        "zero_nodes": [],
    },
    "xlarge-zeros": {
        "num_nodes": 8,
        "node_names": ["linear-2d", "dot-product", "sigmoid-gating", "sum", "zeros-2d", "transformer", "zeros-3d",
                       "linear-3d"],
        "dense_node_dims": [16, 32, 64, 128, 256, 512, 768, 1024],
        "sparse_node_dims": [16, 32, 48, 64],
        "dense_nodes": [0, 1, 2, 3, 4],
        "sparse_nodes": [5, 6, 7],
        "zero_nodes": [4, 6],
    },
    "autoctr": {
        "num_nodes": 3,
        "node_names": ["linear-2d", "dot-product", "linear-3d"],
        "dense_node_dims": [16, 32, 64, 128, 256, 512, 768, 1024],
        "sparse_node_dims": [16, 32, 48, 64],
        "dense_nodes": [0, 1],
        "sparse_nodes": [2],
        "zero_nodes": [],
    },
}
assert_valid_ops_config(ops_config_lib)

_zeros_generator = CleverZeroTensorGenerator()
_mask_generator = CleverMaskGenerator()

# supernet.py:188-207
path_sampling_strategy_lib = {
    "default": {"macro": "any-path", "micro": "single-path"},
    "single-path": {"macro": "single-path", "micro": "single-path"},
    "any-path": {"macro": "any-path", "micro": "any-path"},
    "full-path": {"macro": "full-path", "micro": "full-path"},
    "fixed-path": {"macro": "fixed-path", "micro": "fixed-path"},
    "evo-2shot-path": {"macro": "evo-2shot-path", "micro": "evo-2shot-path"},
}

DS_INTERACT_NUM_SPLITS = 8       # supernet.py:882
_MACRO_KEYS = ("dense_idx", "sparse_idx", "dense_left_idx", "dense_right_idx")


def _ints(v) -> List[int]:
    return [int(x) for x in np.asarray(v).reshape(-1).tolist()]


def _warmup_thresh(counter: int, steps: int) -> float:
    """supernet.py:446-453 / 1014-1020: linear decay of the full-path probability."""
    if counter < steps and counter > 0:
        return 1.0 - counter / (steps + 1e-10)
    return 0


class _DSrc:
    __slots__ = ("v", "w")

    def __init__(self, v: Var, w: int):
        self.v, self.w = v, w


class _SSrc:
    __slots__ = ("v", "s", "g")

    def __init__(self, v: Var, s: int, g: int):
        self.v, self.s, self.g = v, s, g


class SuperNet(nn.Module):
    """Top-level supernet (supernet.py:210-880)."""

    def __init__(
        self,
        num_blocks: int,
        ops_config: Any,
        use_layernorm: bool,
        activation: str = "relu",
        num_embeddings: List[int] = NUM_EMBEDDINGS_CRITEO,
        sparse_input_size: int = 26,
        embedding_dim: int = 16,
        last_n_blocks_out: int = 1,
        path_sampling_strategy: str = "default",
        fixed: bool = False,
        fixed_choice: Any = None,
        place_embedding_on_cpu: bool = False,
        anypath_choice: str = "uniform",
        supernet_training_steps: int = 0,
        candidate_choices: Optional[List] = None,
        use_final_sigmoid: bool = False,
    ):
        super().__init__()
        assert num_blocks >= 1, ValueError(
            "Supernet must contain a minimum of 1 block, but found {}!".format(num_blocks))
        if embedding_dim != EMB:
            raise NotImplementedError("the sm_100a kernels are specialised for embedding_dim=16 (64-byte rows)")
        if last_n_blocks_out != 1:
            raise NotImplementedError("last_n_blocks_out != 1 is never used by the reference entry points")
        if place_embedding_on_cpu:
            raise NotImplementedError("place_embedding_on_cpu is out of scope: tables live in HBM (SURVEY 8a/a1)")
        self._num_blocks = num_blocks
        self._ops_config = ops_config
        self._use_layernorm = use_layernorm
        self._activation = activation
        self._last_n_blocks_out = last_n_blocks_out
        self._sparse_input_size = sparse_input_size
        self._num_embeddings = num_embeddings
        self._embedding_dim = embedding_dim
        self._path_sampling_strategy = path_sampling_strategy
        self._macro_path_sampling_strategy = path_sampling_strategy_lib[path_sampling_strategy]["macro"]
        self._candidate_choices = candidate_choices
        self._fixed = fixed
        self._embedding = self._embedding_layers(sparse_input_size, num_embeddings, embedding_dim)
        self._final = nn.LazyLinear(1)
        self._final_sigmoid = nn.Sigmoid() if use_final_sigmoid else None
        self._place_embedding_on_cpu = place_embedding_on_cpu
        self._supernet_training_steps = supernet_training_steps
        self._anypath_choice_fn = anypath_choice_fn[anypath_choice]
        self._supernet_train_steps_counter = -1
        self._device_args = None
        self._fixed_path_called = False
        self._tables = eng.EmbeddingTables()
        self._num_dense_features: Optional[int] = None

        if self._fixed and fixed_choice is not None:
            self.choice = fixed_choice
            self.macro_last_choice = fixed_choice["macro"]
        else:
            self.choice = []
            self.macro_last_choice = None
        if self._fixed:
            assert self._macro_path_sampling_strategy == "fixed-path", ValueError(
                "'fixed_path_strategy' should be explicitly specified when 'fixed' option is True.")

        blocks = []
        for idx in range(num_blocks):
            cfg = ops_config[idx] if isinstance(ops_config, list) else ops_config
            blocks.append(SuperNetBlock(
                cfg, use_layernorm, int(max(cfg["dense_node_dims"])), int(max(cfg["sparse_node_dims"])),
                embedding_dim, activation,
                path_sampling_strategy=path_sampling_strategy_lib[path_sampling_strategy]["micro"],
                fixed=fixed,
                fixed_micro_choice=None if (fixed_choice is None) or (not fixed) else fixed_choice["micro"][idx],
                anypath_choice=anypath_choice, supernet_training_steps=supernet_training_steps,
                sparse_input_size=sparse_input_size))
        self._blocks = nn.ModuleList(blocks)

    # ------------------------------------------------------------------ reference API
    def get_dense_parameters(self):                      # supernet.py:357-361
        return list(self._blocks.parameters()) + list(self._final.parameters())

    def get_sparse_parameters(self):                     # supernet.py:364-366
        return list(self._embedding.parameters())

    def load_embeddings_from_dlrm(self, dlrm_ckpt_path=None):    # supernet.py:368-383
        if dlrm_ckpt_path is not None:
            checkpoint = torch.load(dlrm_ckpt_path, map_location=torch.device("cpu"))
            assert "model_state_dict" in checkpoint.keys(), "Please use the DLRM checkpoint to load!"
            checkpoint = checkpoint["model_state_dict"]
            for idx in range(len(self._embedding)):
                device = self._embedding[idx].weight.data.device
                self._embedding[idx].weight.data = checkpoint["embedding_layers.{}.weight".format(idx)].to(device)

    def configure_path_sampling_strategy(self, strategy):       # supernet.py:385-402
        assert strategy in path_sampling_strategy_lib, "Strategy {} is not found!".format(strategy)
        self._path_sampling_strategy = strategy
        self._macro_path_sampling_strategy = path_sampling_strategy_lib[strategy]["macro"]
        for block in self._blocks:
            block._micro_path_sampling_strategy = path_sampling_strategy_lib[strategy]["micro"]

    def _embedding_layers(self, sparse_input_size, num_embeddings, embedding_dim):   # supernet.py:404-410
        return nn.ModuleList([nn.Embedding(num_embeddings[i], embedding_dim) for i in range(sparse_input_size)])

    def configure_choice(self, choice: Any):             # supernet.py:842-848
        self.choice = copy.deepcopy(choice)
        self.macro_last_choice = copy.deepcopy(choice["macro"])
        for idx in range(self._num_blocks):
            self._blocks[idx].configure_choice(choice["micro"][idx])

    def set_mode_to_finelune_last_only(self):            # supernet.py:850-853 (sic)
        self._embedding.requires_grad_(False)
        self._blocks.requires_grad_(False)
        self._final.requires_grad_(True)

    def set_mode_to_normal_mode(self):                   # supernet.py:855-858
        self._embedding.requires_grad_(True)
        self._blocks.requires_grad_(True)
        self._final.requires_grad_(True)

    def set_mode_to_layernorm_calibrate(self):           # supernet.py:860-868
        self._embedding.requires_grad_(False)
        self._blocks.requires_grad_(False)
        self._final.requires_grad_(False)
        for _, m in self._blocks.named_modules():
            if isinstance(m, nn.LayerNorm):
                m.requires_grad_(True)

    def set_mode_to_finetune_no_embedding(self):         # supernet.py:870-873
        self._embedding.requires_grad_(False)
        self._blocks.requires_grad_(True)
        self._final.requires_grad_(True)

    def get_all_subnet_macro_choices(self, block_idx: int):      # supernet.py:670-712
        n = 1 + block_idx
        out = {"dense_left_idx": [], "dense_right_idx": [], "dense_idx": [], "sparse_idx": []}
        for k in range(1, n + 1):
            out["dense_idx"] += list(combinations(list(range(n)), k))
            out["sparse_idx"] += list(combinations(list(range(n)), k))
        for k in range(1, min(2, n + 1)):
            out["dense_left_idx"] += list(combinations(list(range(n)), k))
            out["dense_right_idx"] += list(combinations(list(range(n)), k))
        return out

    def get_all_subnet_choices(self):                    # supernet.py:714-721
        all_choices = {"macro": [], "micro": []}
        for b in range(self._num_blocks):
            all_choices["macro"].append(self.get_all_subnet_macro_choices(b))
            all_choices["micro"].append(self._blocks[b].get_all_subnet_micro_choices())
        return all_choices

    # ------------------------------------------------------------------ macro samplers (RNG-order exact)
    def _get_single_path_choice(self, n: int):           # supernet.py:723-736
        bi = pick_with_replacement(n, 1 * 2)
        return {"dense_idx": [pick_index(n)], "sparse_idx": [pick_index(n)],
                "dense_left_idx": [int(bi[0])], "dense_right_idx": [int(bi[1])]}

    def _draw_multi(self, n: int, count_fn):
        nd_ = count_fn(n)
        ns_ = count_fn(n)
        bi = pick_with_replacement(n, 1 * 2)
        dense = pick_without_replacement(n, nd_)
        sparse = pick_without_replacement(n, ns_)
        return {"dense_idx": dense, "sparse_idx": sparse, "dense_left_idx": [int(bi[0])],
                "dense_right_idx": [int(bi[1])]}

    def _get_any_path_choice(self, n: int):              # supernet.py:738-770
        return self._draw_multi(n, self._anypath_choice_fn)

    def _get_fixed_path_choice(self, n: int):            # supernet.py:772-812 (always 'uniform')
        return self._draw_multi(n, anypath_choice_fn["uniform"])

    def _get_full_path_choice(self, n: int):             # supernet.py:814-824
        return {k: np.arange(n) for k in _MACRO_KEYS}

    def _get_choice(self):                               # supernet.py:432-511
        thresh = _warmup_thresh(self._supernet_train_steps_counter, self._supernet_training_steps)
        strat = self._macro_path_sampling_strategy
        nb = self._num_blocks
        if strat == "single-path":
            full = np.random.random() < thresh
            choice = [self._get_full_path_choice(1 + i) if full else self._get_single_path_choice(1 + i)
                      for i in range(nb)]
        elif strat == "full-path":
            choice = [self._get_full_path_choice(1 + i) for i in range(nb)]
        elif strat == "any-path":
            full = np.random.random() < thresh
            choice = [self._get_full_path_choice(1 + i) if full else self._get_any_path_choice(1 + i)
                      for i in range(nb)]
        elif strat == "fixed-path" and self.macro_last_choice is None:
            if self._fixed_path_called:
                raise ValueError("Error! fixed-path choice should be generated only once for each supernet!")
            self._fixed_path_called = True
            choice = [self._get_fixed_path_choice(1 + i) for i in range(nb)]
        elif strat == "fixed-path":
            choice = self.macro_last_choice
        elif strat == "evo-2shot-path":
            assert self._candidate_choices is not None, \
                "You must specify self._candidate_choices before using 'evo-2shot-path'!"
            cand = self._candidate_choices[np.random.randint(len(self._candidate_choices))]["choice"]
            for i in range(nb):
                self._blocks[i].configure_choice(cand["micro"][i])
            choice = cand["macro"]
        else:
            raise NotImplementedError("Path strategy {} is not supported!".format(strat))
        if strat != "full-path":
            self.__dict__["macro_last_choice"] = choice
        return choice

    # ------------------------------------------------------------------ lazy life cycle
    def materialize(self, num_dense_features: int):
        """Host-only equivalent of the reference's warm-up forward
        (train_utils.py:413-433): size every lazy layer for the zero-padded widths
        D_i = nd + 1024 i, S_i = F + 72 i (weight sharing) or for the chosen widths
        (fixed), deleting the layers the reference deletes."""
        nd_, F = int(num_dense_features), self._sparse_input_size
        dev = self._embedding[0].weight.device
        w_hist, r_hist = [nd_], [F]
        for i, blk in enumerate(self._blocks):
            if self._fixed:
                mac, mic = self.macro_last_choice[i], blk.micro_last_choice
                if mac is None or mic is None:
                    raise RuntimeError("fixed model without a choice cannot be materialised before its first forward")
                sel = {k: sorted(set(_ints(mac[k]))) for k in _MACRO_KEYS}
                Kd = sum(w_hist[j] for j in sel["dense_idx"])
                Ks = sum(r_hist[j] for j in sel["sparse_idx"])
                Kl = sum(w_hist[j] for j in sel["dense_left_idx"])
                Kr = sum(w_hist[j] for j in sel["dense_right_idx"])
                d, s = int(mic["dense_in_dims"]), int(mic["sparse_in_dims"])
                dsi = int(mic["dense_sparse_interact"])
            else:
                Kd = Kl = Kr = nd_ + blk._max_dims_or_dims_dense * i if i else nd_
                Ks = F + (blk._max_dims_or_dims_sparse + DS_INTERACT_NUM_SPLITS) * i if i else F
                d, s, dsi = blk._max_dims_or_dims_dense, blk._max_dims_or_dims_sparse, 1
            blk._prepare_nodes(Kd, Ks, Kl, Kr, d, dev)
            w_hist.append(d)
            r_hist.append(s + (DS_INTERACT_NUM_SPLITS if (dsi == 1 or not self._fixed) else 0))
        _materialize(self._final, w_hist[-1] + r_hist[-1] * EMB)
        self._num_dense_features = nd_
        return self

    def _needs_materialize(self) -> bool:
        return isinstance(self._final, nn.LazyLinear) and self._final.has_uninitialized_params()

    # ------------------------------------------------------------------ forward
    def forward(self, int_feats: torch.Tensor, cat_feats: torch.Tensor, choices=None):
        if not int_feats.is_cuda:
            raise RuntimeError("nasrec_b200.SuperNet runs on CUDA only (no CPU fallback); move the model and "
                               "inputs to a B200 with .to('cuda')")
        macro, micro = self._sample(choices)
        if self._needs_materialize():
            self.materialize(int_feats.shape[1])
        cat = cat_feats if cat_feats.dtype == torch.int64 else cat_feats.long()
        cat = cat.contiguous()

        def body(run: Run, iv: Sequence[Var]):
            return [self._run_network(run, iv[0], cat, macro, micro)]

        out = run_with_autograd(self, [int_feats, cat], body)[0]
        return self._final_sigmoid(out) if self._final_sigmoid is not None else out

    def fixed_forward(self, int_feats, cat_feats, choices: Any):     # supernet.py:605-668
        return self.forward(int_feats, cat_feats, choices)

    def _sample(self, choices=None) -> Tuple[List[Dict], List[Dict]]:
        """Host side of forward (supernet.py:513-529, 574-585): counters, macro then
        per-block micro draws in block order, self.choice bookkeeping."""
        d = self.__dict__
        if not self._fixed:
            d["_supernet_train_steps_counter"] = self._supernet_train_steps_counter + 1
        d["choice"] = {"micro": [], "macro": []}
        macro = self._get_choice() if choices is None else choices["macro"]
        self.choice["macro"] = macro
        micro = []
        for i, blk in enumerate(self._blocks):
            # the reference passes choices["micro"] (the whole list) here, which raises
            # TypeError; the per-block entry is what was meant.
            micro.append(blk._sample(None if choices is None else choices["micro"][i]))
            self.choice["micro"].append(blk.choice)
        return macro, micro

    def _liveness(self, macro, micro) -> List[bool]:
        """Blocks that can reach the logit (SURVEY A.12 iv).  A block is live iff it is
        the last one or a later live block has an active node consuming the kind of
        input it was selected for."""
        nb = self._num_blocks
        need = [False] * (nb + 1)          # need[j]: source j (0 = stem, i+1 = block i) is consumed
        need[nb] = True
        for i in reversed(range(nb)):
            if not need[i + 1]:
                continue
            names = [self._blocks[i]._node_names[n] for n in _ints(micro[i]["active_nodes"])]
            mac = macro[i]
            if any(n in ("linear-2d", "dot-product") for n in names):
                for j in _ints(mac["dense_idx"]):
                    need[j] = True
            if any(n in _dense_binary_nodes for n in names):
                for j in _ints(mac["dense_left_idx"]) + _ints(mac["dense_right_idx"]):
                    need[j] = True
            if any(n in ("transformer", "linear-3d", "dot-product") for n in names):
                for j in _ints(mac["sparse_idx"]):
                    need[j] = True
        return need[1:]

    def _run_network(self, run: Run, int_x: Var, cat_x: torch.Tensor, macro, micro) -> Var:
        segs = self._run_trunk(run, int_x, cat_x, macro, micro)
        return self._run_head(run, segs, int_x.t.shape[0])

    def _run_head(self, run: Run, segs: Sequence[Seg], B: int) -> Var:
        """The final projection to the logit (supernet.py:592-598) on the trunk's segment list."""
        return eng.linear_ln(run.tape, segs, B, run.pv(self._final.weight), run.pv(self._final.bias), None, False, 1,
                             w_full_support=self._fixed)

    def _run_trunk(self, run: Run, int_x: Var, cat_x: torch.Tensor, macro, micro) -> List[Seg]:
        """Stem + choice blocks; returns the last block's outputs as the segment list the
        final layer consumes (what the reference concatenates at supernet.py:592-597)."""
        B, nd_ = int_x.t.shape
        F = self._sparse_input_size
        if cat_x.shape != (B, F):
            raise ValueError("cat_feats must be [%d, %d], got %s" % (B, F, tuple(cat_x.shape)))
        sp0 = eng.embedding(run.tape, self._tables, [run.pv(m.weight) for m in self._embedding], cat_x,
                            sparse_sink=run.sparse_sink, cache=run.emb_cache)
        dsrc: List[Optional[_DSrc]] = [_DSrc(int_x, nd_)]
        ssrc: List[Optional[_SSrc]] = [_SSrc(sp0, F, 0)]
        live = self._liveness(macro, micro)
        for i, blk in enumerate(self._blocks):
            if not live[i]:
                dsrc.append(None)
                ssrc.append(None)
                continue
            bundle = blk._segments(dsrc, ssrc, macro[i], micro[i], nd_, F)
            d_out, s_out = blk._run(run, micro[i], bundle, B, int_x.t.device)
            dsrc.append(d_out)
            ssrc.append(s_out)
        dl, sl = dsrc[-1], ssrc[-1]
        blk = self._blocks[-1]
        if self._fixed:
            segs = [Seg(dl.v, 0, dl.w, dl.w, 0), Seg(sl.v, 0, (sl.s + sl.g) * EMB, (sl.s + sl.g) * EMB, dl.w)]
        else:                                            # supernet.py:592-597 on the zero-padded layout
            maxd, maxs = blk._max_dims_or_dims_dense, blk._max_dims_or_dims_sparse
            bs = (sl.s + sl.g) * EMB
            segs = [Seg(dl.v, 0, dl.w, dl.w, 0), Seg(sl.v, 0, bs, sl.s * EMB, maxd)]
            if sl.g:
                segs.append(Seg(sl.v, sl.s * EMB, bs, sl.g * EMB, maxd + maxs * EMB))
        return segs

    def to(self, *args, **kwargs):                       # supernet.py:826-840 (CPU-embedding branch dropped)
        self._device_args = args
        return super().to(*args, **kwargs)

    def discretize_config_each_block(self, probs):       # supernet.py:875-880
        return [self._blocks[i].discretize_config_each_block(probs[i]) for i in range(self._num_blocks)]


class SuperNetBlock(nn.Module):
    """One choice block (supernet.py:884-1381)."""

    def __init__(self, ops_config: Any, use_layernorm: bool, max_dims_or_dims_dense: int,
                 max_dims_or_dims_sparse: int, embedding_dim: int, activation: str = "relu",
                 path_sampling_strategy: str = "single-path", fixed: bool = False, fixed_micro_choice=None,
                 anypath_choice: str = "uniform", supernet_training_steps: int = 0, sparse_input_size: int = 26):
        super().__init__()
        self._num_nodes = ops_config["num_nodes"]
        self._dense_nodes = ops_config["dense_nodes"]
        self._sparse_nodes = ops_config["sparse_nodes"]
        self._node_names = ops_config["node_names"]
        self._dense_node_dims = ops_config["dense_node_dims"]
        self._sparse_node_dims = ops_config["sparse_node_dims"]
        self._zero_nodes = ops_config["zero_nodes"]
        self._sparse_input_size = sparse_input_size
        self._use_layernorm = use_layernorm
        self._max_dims_or_dims_dense = max_dims_or_dims_dense
        self._max_dims_or_dims_sparse = max_dims_or_dims_sparse
        self._embedding_dim = embedding_dim
        self._activation = activation
        self._micro_path_sampling_strategy = path_sampling_strategy
        self._fixed = fixed
        self._fixed_micro_choice = fixed_micro_choice
        self._anypath_choice_fn = anypath_choice_fn[anypath_choice]
        self._supernet_training_steps = supernet_training_steps
        self._device_args = None
        self._supernet_train_steps_counter = -1
        self._fixed_path_called = False
        self._nodes = nn.ModuleList()
        self.micro_last_choice = self._fixed_micro_choice if self._fixed else None
        if self._fixed:
            choice = self._get_choice()
            choice_nodes = _ints(choice["active_nodes"])
        else:
            choice = None
            choice_nodes = list(range(self._num_nodes))
        for i in range(self._num_nodes):                 # supernet.py:950-982
            if i not in choice_nodes:
                self._nodes.append(nn.ModuleList([]))
                continue
            name = self._node_names[i]
            if name in _dense_binary_nodes + _dense_unary_nodes:
                node = _node_choices[name](use_layernorm,
                                           int(choice["dense_in_dims"]) if fixed else max_dims_or_dims_dense,
                                           activation, fixed=fixed)
            elif name in _sparse_nodes:
                node = _node_choices[name](use_layernorm,
                                           int(choice["sparse_in_dims"]) if fixed else max_dims_or_dims_sparse,
                                           embedding_dim=embedding_dim, fixed=fixed, activation=activation)
            elif name in _dense_sparse_nodes:
                node = _node_choices[name](use_layernorm,
                                           int(choice["dense_in_dims"]) if fixed else max_dims_or_dims_dense,
                                           embedding_dim, fixed=fixed)
            else:
                raise NotImplementedError("Block name {} is not supported in supernet!".format(name))
            self._nodes.append(node)
        # dense -> sparse merger (supernet.py:985-995)
        if self._fixed and int(choice["dense_sparse_interact"]) == 0:
            self.project_emb_dim, self.project_emb_dim_layernorm = None, None
        else:
            self.ds_interact_expanded_dim = DS_INTERACT_NUM_SPLITS * embedding_dim
            self.project_emb_dim = nn.LazyLinear(self.ds_interact_expanded_dim, bias=not use_layernorm)
            self.project_emb_dim_layernorm = nn.LayerNorm(self.ds_interact_expanded_dim, eps=1e-5) \
                if use_layernorm else None
        # sparse -> dense merger (supernet.py:997-1003)
        if self._fixed and int(choice["deep_fm"]) == 0:
            self.deep_fm, self.deep_fm_output_ln = None, None
        else:
            self.deep_fm_dims = max(self._dense_node_dims) if not self._fixed else int(choice["dense_in_dims"])
            self.deep_fm = FactorizationMachine3D(fixed=self._fixed, use_layernorm=use_layernorm,
                                                  max_dims_or_dims=self.deep_fm_dims)
        self.choice = []

    # ------------------------------------------------------------------ micro samplers (RNG-order exact)
    def _draw_tail(self, active):
        return {"active_nodes": active,
                "dense_in_dims": pick(self._dense_node_dims),
                "sparse_in_dims": pick(self._sparse_node_dims),
                "dense_sparse_interact": pick_index(2),
                "deep_fm": pick_index(2)}

    def _get_single_path_choice(self):                   # supernet.py:1244-1263
        while True:
            active = sorted([pick(self._dense_nodes)] + [pick(self._sparse_nodes)])
            choice = self._draw_tail(active)
            if choice["active_nodes"] != self._zero_nodes:
                return choice

    def _get_full_path_choice(self):                     # supernet.py:1265-1276
        return {"active_nodes": np.arange(self._num_nodes), "dense_in_dims": np.max(self._dense_node_dims),
                "sparse_in_dims": np.max(self._sparse_node_dims), "dense_sparse_interact": 1, "deep_fm": 1}

    def _get_any_path_choice(self):                      # supernet.py:1278-1303
        while True:
            n_d = self._anypath_choice_fn(len(self._dense_nodes))
            n_s = self._anypath_choice_fn(len(self._sparse_nodes))
            dense = pick_without_replacement(self._dense_nodes, n_d)
            sparse = pick_without_replacement(self._sparse_nodes, n_s)
            choice = self._draw_tail(sorted(dense + sparse))
            if choice["active_nodes"] != self._zero_nodes:
                return choice

    def _get_fixed_path_choice(self):                    # supernet.py:1305-1313
        return self._get_single_path_choice()

    def _get_choice(self):                               # supernet.py:1009-1061
        thresh = _warmup_thresh(self._supernet_train_steps_counter, self._supernet_training_steps)
        strat = self._micro_path_sampling_strategy
        if strat == "single-path":
            choice = self._get_full_path_choice() if np.random.random() < thresh else self._get_single_path_choice()
        elif strat == "full-path":
            choice = self._get_full_path_choice()
        elif strat == "any-path":
            choice = self._get_full_path_choice() if np.random.random() < thresh else self._get_any_path_choice()
        elif strat == "fixed-path" and self.micro_last_choice is None:
            if self._fixed_path_called:
                raise ValueError("Error! fixed-path choice should be generated only once for each supernet!")
            self._fixed_path_called = True
            choice = self._get_fixed_path_choice()
        elif strat in ("fixed-path", "evo-2shot-path"):
            choice = self.micro_last_choice
        else:
            raise NotImplementedError("Path strategy {} is not supported!".format(strat))
        if strat != "full-path":
            self.__dict__["micro_last_choice"] = choice      # (plain attribute: skip nn.Module.__setattr__, hot path)
        return choice

    def _sample(self, choices=None):
        """Host side of SuperNetBlock.forward (supernet.py:1067-1076)."""
        choice = self._get_choice() if choices is None else choices
        d = self.__dict__
        d["choice"] = choice
        if not self._fixed:
            d["_supernet_train_steps_counter"] = self._supernet_train_steps_counter + 1
        return choice

    def configure_choice(self, choice):                  # supernet.py:1316-1318
        self.choice = copy.deepcopy(choice)
        self.micro_last_choice = copy.deepcopy(choice)

    def get_all_subnet_micro_choices(self):              # supernet.py:1164-1183
        out = {"active_nodes": [], "dense_in_dims": [], "sparse_in_dims": [], "dense_sparse_interact": [0, 1]}
        for sn in self._sparse_nodes:
            for dn in self._dense_nodes:
                out["active_nodes"].append((dn, sn))
        out["dense_in_dims"] = [(x,) for x in self._dense_node_dims]
        out["sparse_in_dims"] = [(x,) for x in self._sparse_node_dims]
        return out

    def discretize_config_each_block(self, probs, dense_nodes_topk=2, sparse_nodes_topk=1, in_dims_topk=2,
                                     include_zeros_3d=False):      # supernet.py:1320-1377
        cfg = {"num_nodes": 0, "node_names": [], "dense_node_dims": [], "dense_nodes": [], "sparse_nodes": [],
               "zero_nodes": []}
        cnt = 0
        for key, pool, topk, dst in (("dense_probs", self._dense_nodes, dense_nodes_topk, "dense_nodes"),
                                     ("sparse_probs", self._sparse_nodes, sparse_nodes_topk, "sparse_nodes")):
            order = np.argsort(probs[key])[::-1][:topk]
            if key == "sparse_probs" and include_zeros_3d:
                z = [k for k, n in enumerate(pool) if self._node_names[n] == "zeros-3d"]
                assert z, "'zeros-3d' is not part of this search space"
                if z[0] not in order:
                    order = np.append(order, z[0])
            for k in order:
                n = pool[k]
                cfg["node_names"].append(self._node_names[n])
                cfg[dst].append(cnt)
                if n in self._zero_nodes:
                    cfg["zero_nodes"].append(cnt)
                cnt += 1
        cfg["num_nodes"] = cnt
        order = np.argsort(probs["in_dims_probs"])[::-1][:in_dims_topk]
        cfg["dense_node_dims"] = sorted(self._dense_node_dims[k] for k in order)
        return cfg

    # ------------------------------------------------------------------ lazy life cycle
    def _prepare_nodes(self, Kd: int, Ks: int, Kl: int, Kr: int, d: int, device):
        for i, node in enumerate(self._nodes):
            name = self._node_names[i]
            if isinstance(node, nn.ModuleList):
                continue
            if name == "linear-2d":
                node._prepare(Kd)
            elif name == "dot-product":
                node._prepare(Kd, Ks)
            elif name == "sum":
                node._prepare(max(Kl, Kr))
            elif name == "sigmoid-gating":
                node._prepare(max(Kl, Kr), device)
            elif name in ("transformer", "linear-3d"):
                node._prepare(Ks)
        out_width = d if self._fixed else self._max_dims_or_dims_dense
        if self.project_emb_dim is not None:
            if out_width != self._embedding_dim * DS_INTERACT_NUM_SPLITS:
                _materialize(self.project_emb_dim, out_width)
            else:                                        # supernet.py:1143-1145, 1224-1226
                self.project_emb_dim, self.project_emb_dim_layernorm = None, None
        if self.deep_fm is not None:
            self.deep_fm._prepare()

    # ------------------------------------------------------------------ execution
    @staticmethod
    def _input_needs(names):
        """(dense_idx, sparse_idx, left/right) consumed by the active node kinds."""
        return (any(n in ("linear-2d", "dot-product") for n in names),
                any(n in ("transformer", "linear-3d", "dot-product") for n in names),
                any(n in _dense_binary_nodes for n in names))

    def _run(self, run: Run, micro, bundle, B: int, dev):
        """SuperNetBlock.forward / fixed_forward (supernet.py:1067-1162, 1185-1242) on compact tensors.
        `bundle` = ((dense segs, K), (sparse segs, S), (left segs, K), (right segs, K))."""
        tape = run.tape
        active = _ints(micro["active_nodes"])
        d, s = int(micro["dense_in_dims"]), int(micro["sparse_in_dims"])
        dsi, dfm = int(micro["dense_sparse_interact"]), int(micro["deep_fm"])
        if dsi not in (0, 1):
            raise NotImplementedError("Bug reported for dense/sparse interact.")
        if not self._fixed:
            assert d <= self._max_dims_or_dims_dense and s <= self._max_dims_or_dims_sparse, ValueError(
                "'dims_in_use' should always be smaller than 'max_dims_or_dims'")
        g = DS_INTERACT_NUM_SPLITS if dsi == 1 else 0
        (dsegs, Kd), (ssegs, Ks), (lsegs, Kl), (rsegs, Kr) = bundle

        dense_out = Var(torch.empty(B, d, dtype=torch.float32, device=dev))
        rows = s + g
        sparse_out = Var(torch.empty(B, rows, EMB, dtype=torch.float32, device=dev))
        nd_w = ns_w = 0
        for i in active:
            name, node = self._node_names[i], self._nodes[i]
            if name == "linear-2d":
                node._run(run, dsegs, Kd, B, d, out=dense_out, ldy=d, accumulate=int(nd_w > 0))
                nd_w += 1
            elif name == "dot-product":
                node._run(run, dsegs, Kd, ssegs, Ks, B, d, out=dense_out, ldy=d, accumulate=int(nd_w > 0))
                nd_w += 1
            elif name in ("sum", "sigmoid-gating"):
                node._run(run, lsegs, Kl, rsegs, Kr, B, d, out=dense_out, ldy=d, accumulate=int(nd_w > 0))
                nd_w += 1
            elif name in ("transformer", "linear-3d"):
                node._run(run, ssegs, Ks, B, s, out=sparse_out, out_bstride=rows * EMB, accumulate=int(ns_w > 0))
                ns_w += 1
            elif name in ("zeros-2d", "zeros-3d"):
                pass                                     # contributes exact zeros to the node sum
            else:
                raise NotImplementedError("Block name {} is not supported!".format(name))
        if nd_w == 0:
            dense_out.t.zero_()
        if ns_w == 0:
            sparse_out.t.zero_()

        # dense -> sparse merger reads the node sum *before* the FM term is added (supernet.py:1137-1157)
        dense_sum = dense_out
        out_width = d if self._fixed else self._max_dims_or_dims_dense
        project = dsi == 1 and out_width != self._embedding_dim * DS_INTERACT_NUM_SPLITS
        if project:
            eng.linear_ln(tape, [Seg(dense_sum, 0, d, d, 0)], B, run.pv(self.project_emb_dim.weight),
                          run.pv(self.project_emb_dim.bias), run.ln(self.project_emb_dim_layernorm), False,
                          DS_INTERACT_NUM_SPLITS * EMB, out=sparse_out, out_off=s * EMB, ldy=rows * EMB,
                          w_full_support=self._fixed)
        if dfm == 1:
            if project and tape.enabled:
                # keep the pre-FM sum alive for the merger's weight gradient
                dense_out = Var(torch.empty(B, d, dtype=torch.float32, device=dev))
                eng.copy2d(tape, dense_sum, 0, d, B, d, dense_out, 0, d, 0)
            self.deep_fm._run(run, sparse_out, s, rows * EMB, B, d, out=dense_out, ldy=d, accumulate=1)
        if dsi == 1 and not project:
            # fixed model with d == 128: the merger is a view of dense_out that also sees the
            # in-place FM addition (supernet.py:1224-1236)
            eng.copy2d(tape, dense_out, 0, d, B, d, sparse_out, s * EMB, rows * EMB, 0)
        return _DSrc(dense_out, d), _SSrc(sparse_out, s, g)

    def _segments(self, dsrc, ssrc, macro, micro, nd_: int, F: int):
        """Segment lists standing in for the four torch.cat inputs of the block
        (supernet.py:532-573 weight sharing, :621-638 fixed: ascending source order).
        Only the lists an active node will read are built, so a source skipped as dead
        is never dereferenced."""
        need_dense, need_sparse, need_lr = self._input_needs(
            [self._node_names[i] for i in _ints(micro["active_nodes"])])
        maxd, maxs = self._max_dims_or_dims_dense, self._max_dims_or_dims_sparse
        n_src = len(dsrc)
        sel = {k: sorted(set(_ints(macro[k]))) for k in _MACRO_KEYS}

        def dense(key, want):
            if not want:
                return [], 0
            segs, off = [], 0
            for j in sel[key]:
                src = dsrc[j]
                if src is None:
                    raise RuntimeError("liveness analysis skipped a block that is still consumed")
                w_off = off if self._fixed else (0 if j == 0 else nd_ + maxd * (j - 1))
                segs.append(Seg(src.v, 0, src.w, src.w, w_off))
                off += src.w
            total = off if self._fixed else (nd_ + maxd * (n_src - 1) if n_src > 1 else nd_)
            return segs, total

        def sparse(want):
            if not want:
                return [], 0
            segs, off = [], 0
            for j in sel["sparse_idx"]:
                src = ssrc[j]
                if src is None:
                    raise RuntimeError("liveness analysis skipped a block that is still consumed")
                rows = src.s + src.g
                bs = rows * EMB
                if self._fixed:
                    segs.append(Seg(src.v, 0, bs, rows, off))
                    off += rows
                else:
                    base = 0 if j == 0 else F + (maxs + DS_INTERACT_NUM_SPLITS) * (j - 1)
                    segs.append(Seg(src.v, 0, bs, src.s, base))
                    if src.g:
                        segs.append(Seg(src.v, src.s * EMB, bs, src.g, base + maxs))
            total = off if self._fixed else (F + (maxs + DS_INTERACT_NUM_SPLITS) * (n_src - 1) if n_src > 1 else F)
            return segs, total

        return dense("dense_idx", need_dense), sparse(need_sparse), dense("dense_left_idx", need_lr), \
            dense("dense_right_idx", need_lr)

    def forward(self, tensors, choices=None):
        """Reference-shaped entry point on zero-padded tensors (supernet.py:1067-1162):
        (dense [B,D], sparse [B,S,16], left [B,D], right [B,D]) -> (dense_out, sparse_out)."""
        choice = self._sample(choices)
        dense_t, sparse_t, left_t, right_t = tensors
        B = dense_t.shape[0]
        dev = dense_t.device
        Kd, Ks, Kl, Kr = dense_t.shape[1], sparse_t.shape[1], left_t.shape[1], right_t.shape[1]
        d = int(choice["dense_in_dims"])
        self._prepare_nodes(Kd, Ks, Kl, Kr, d, dev)
        maxd, maxs = self._max_dims_or_dims_dense, self._max_dims_or_dims_sparse

        def body(run: Run, iv):
            bundle = (([Seg(iv[0], 0, Kd, Kd, 0)], Kd), ([Seg(iv[1], 0, Ks * EMB, Ks, 0)], Ks),
                      ([Seg(iv[2], 0, Kl, Kl, 0)], Kl), ([Seg(iv[3], 0, Kr, Kr, 0)], Kr))
            do, so = self._run(run, choice, bundle, B, dev)
            if self._fixed:
                return [do.v, so.v]
            # re-inflate to the reference's zero-padded layout
            dp = Var(torch.zeros(B, maxd, dtype=torch.float32, device=dev))
            eng.copy2d(run.tape, do.v, 0, do.w, B, do.w, dp, 0, maxd, 0)
            rows = maxs + DS_INTERACT_NUM_SPLITS
            spv = Var(torch.zeros(B, rows, EMB, dtype=torch.float32, device=dev))
            eng.copy2d(run.tape, so.v, 0, (so.s + so.g) * EMB, B, so.s * EMB, spv, 0, rows * EMB, 0)
            if so.g:
                eng.copy2d(run.tape, so.v, so.s * EMB, (so.s + so.g) * EMB, B, so.g * EMB, spv, maxs * EMB,
                           rows * EMB, 0)
            return [dp, spv]

        outs = run_with_autograd(self, [dense_t, sparse_t, left_t, right_t], body)
        return outs[0], outs[1]

    def fixed_forward(self, tensors, choices):           # supernet.py:1185-1242
        return self.forward(tensors, choices)
