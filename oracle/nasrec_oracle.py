"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU restatement (torch fp32 on CPU + numpy) of the NASRec supernet hot path:
stem -> 7 choice blocks -> final logit, the BCE/clip/Adagrad step body, and the
integer side of the embedding path (row gather, unique-row sets).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product (``nasrec_b200``)
never does; it fails loudly when its CUDA library is missing.

Parity pin: this restatement is checked against the *unmodified reference run
in the build container* through the fixtures in ``tests/golden/`` (generated
by ``tests/golden/make_golden.py``, which imports /root/reference).  The
reference itself ships no tests or golden vectors, so the
fixtures are outputs of the reference itself.

The arithmetic lives in a third-party dependency of the reference (PyTorch:
``F.linear``, ``F.layer_norm``, softmax, ``torch.optim.Adagrad``,
``clip_grad_norm_``; reference pin pytorch=1.12.0, environment.yml:179; this
image has 2.11.0).  This file restates the reference's *composition* of those
primitives, deliberately keeping what makes the reference slow on purpose
(zero-padded concats, full-width masked modules, dense embedding gradients and
a dense Adagrad over every table row) so that timing it is an honest CPU
baseline of the reference algorithm.

Every function cites the reference file:line it follows
(paths relative to /root/reference/).
"""
from __future__ import annotations

import math
import zlib
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EMB_DIM = 16            # nasrec/supernet/supernet.py:224
DS_SPLITS = 8           # nasrec/supernet/supernet.py:882  DS_INTERACT_NUM_SPLITS
MHA_HEADS = 8           # nasrec/supernet/modules.py:26
LN_EPS = 1e-5           # nn.LayerNorm default, modules.py:158,625,630
LN_INIT = 0.17          # nasrec/supernet/modules.py:598

# nasrec/supernet/supernet.py:134-178 (search-space definitions; data, restated)
OPS_CONFIG = {
    "xlarge": dict(
        num_nodes=6,
        node_names=["linear-2d", "dot-product", "sigmoid-gating", "sum", "transformer", "linear-3d"],
        dense_node_dims=[16, 32, 64, 128, 256, 512, 768, 1024],
        sparse_node_dims=[16, 32, 48, 64],
        dense_nodes=[0, 1, 2, 3], sparse_nodes=[4, 5], zero_nodes=[]),
    "xlarge-zeros": dict(
        num_nodes=8,
        node_names=["linear-2d", "dot-product", "sigmoid-gating", "sum", "zeros-2d",
                    "transformer", "zeros-3d", "linear-3d"],
        dense_node_dims=[16, 32, 64, 128, 256, 512, 768, 1024],
        sparse_node_dims=[16, 32, 48, 64],
        dense_nodes=[0, 1, 2, 3, 4], sparse_nodes=[5, 6, 7], zero_nodes=[4, 6]),
    "autoctr": dict(
        num_nodes=3,
        node_names=["linear-2d", "dot-product", "linear-3d"],
        dense_node_dims=[16, 32, 64, 128, 256, 512, 768, 1024],
        sparse_node_dims=[16, 32, 48, 64],
        dense_nodes=[0, 1], sparse_nodes=[2], zero_nodes=[]),
}

DENSE_UNARY = ("linear-2d", "zeros-2d")       # supernet.py:116
DENSE_BINARY = ("sum", "sigmoid-gating")      # supernet.py:118
DENSE_SPARSE = ("dot-product",)               # supernet.py:120
SPARSE_NODES = ("zeros-3d", "transformer", "linear-3d")  # supernet.py:122


# --------------------------------------------------------------------------
# deterministic state filling shared by the golden generator and the tests
# --------------------------------------------------------------------------
def fill_state_dict(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic, order-independent values for every tensor of a state dict.

    One numpy RandomState per tensor, keyed by crc32(name) ^ seed, so both the
    golden generator (which copies the result into the reference model) and the
    tests (which feed the oracle / the CUDA path) build bit-identical weights.
    Scales are chosen to keep activations O(1) through 7 blocks.
    """
    out = {}
    for name in shapes:
        shape = tuple(int(s) for s in shapes[name])
        rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        n = int(np.prod(shape)) if len(shape) else 1
        leaf = name.rsplit(".", 1)[-1]
        is_ln = ("layernorm" in name) or name.split(".")[-2].endswith("_ln") or ("_ln." in name)
        if name.startswith("_embedding."):
            v = rs.standard_normal(n).astype(np.float32) * 0.25
        elif is_ln and leaf == "weight":
            v = (1.0 + 0.1 * rs.standard_normal(n)).astype(np.float32)
            if "_attn_ln" in name or "_attn_fc_ln" in name:
                v = (LN_INIT * (1.0 + 0.1 * rs.standard_normal(n))).astype(np.float32)
        elif leaf in ("bias", "in_proj_bias"):
            v = (0.05 * rs.standard_normal(n)).astype(np.float32)
        else:  # linear-like weight [out, in]
            fan_out, fan_in = shape[0], shape[-1]
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = rs.uniform(-a, a, n).astype(np.float32)
        out[name] = torch.from_numpy(v.reshape(shape).copy())
    return out


def synth_batch(batch: int, num_dense: int, num_embeddings: Sequence[int], seed: int,
                zipf: bool = False, all_zero_dense: bool = False):
    """Synthetic batch in the shapes of data_pipes.py:135-175.

    dense: log(max(0,x)+1) of Poisson counts (data_pipes.py:137); ids in [0, N_f)
    with id 0 = "missing" with prob 0.02 (data_pipes.py:141,164); y ~ Bernoulli(.25).
    """
    rs = np.random.RandomState(seed)
    if all_zero_dense:   # Avazu pseudo-dense feature, data_pipes.py:181
        int_x = np.zeros((batch, num_dense), np.float32)
    else:
        int_x = np.log1p(rs.poisson(3.0, (batch, num_dense))).astype(np.float32)
    cols = []
    for n in num_embeddings:
        if n <= 1:
            c = np.zeros(batch, np.int64)
        elif zipf:
            r = np.minimum(rs.zipf(1.05, batch), n - 1).astype(np.int64)
            c = (r * 2654435761 % (n - 1)) + 1
        else:
            c = rs.randint(1, n, batch).astype(np.int64)
        c[rs.rand(batch) < 0.02] = 0
        cols.append(c)
    cat_x = np.stack(cols, 1)
    y = (rs.rand(batch, 1) < 0.25).astype(np.float32)
    return torch.from_numpy(int_x), torch.from_numpy(cat_x), torch.from_numpy(y)


# --------------------------------------------------------------------------
# integer side of the embedding path (numpy; bit-exact comparisons)
# --------------------------------------------------------------------------
def embedding_gather(tables: Sequence[np.ndarray], cat_x: np.ndarray) -> np.ndarray:
    """sparse[b,f,:] = W_f[cat[b,f],:]   (supernet.py:420-426)."""
    B, Fn = cat_x.shape
    out = np.empty((B, Fn, tables[0].shape[1]), np.float32)
    for f in range(Fn):
        out[:, f, :] = tables[f][cat_x[:, f]]
    return out


def embedding_row_sets(cat_x: np.ndarray) -> List[np.ndarray]:
    """Rows of each table that receive a gradient = sorted unique ids per column
    (dense nn.Embedding grad is nonzero exactly there; supernet.py:407)."""
    return [np.unique(cat_x[:, f]) for f in range(cat_x.shape[1])]


def embedding_grad_rows(cat_x: np.ndarray, gout: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Per table: (sorted unique rows, summed grad rows); duplicates summed in
    ascending sample order -- the order the CUDA sort-reduce uses."""
    res = []
    for f in range(cat_x.shape[1]):
        rows, inv = np.unique(cat_x[:, f], return_inverse=True)
        acc = np.zeros((rows.shape[0], gout.shape[2]), np.float32)
        for b in range(cat_x.shape[0]):          # ascending b => deterministic
            acc[inv[b]] += gout[b, f]
        res.append((rows, acc))
    return res


# --------------------------------------------------------------------------
# building blocks (torch fp32, differentiable)
# --------------------------------------------------------------------------
def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def _ln(sd, key, x):
    w = sd[key + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[key + ".bias"], LN_EPS)


def _prefix_mask(n: int, d: int) -> torch.Tensor:
    """[1]*d + [0]*(n-d)   (modules.py:90-93)."""
    m = torch.zeros(n)
    m[:d] = 1.0
    return m


def _pad_equal(left, right):
    """modules.py:403-430."""
    dl, dr = left.shape[-1], right.shape[-1]
    if dl == dr:
        return left, right
    z = torch.zeros(left.shape[0], abs(dl - dr))
    if dl < dr:
        return torch.cat([left, z], 1), right
    return left, torch.cat([right, z], 1)


def fc(sd, pfx, x, d, maxd, ln, fixed):
    """ElasticLinear, modules.py:162-181."""
    out = _lin(sd, pfx + "._linear", x)
    if ln:
        out = _ln(sd, pfx + "._layernorm", out)
    out = torch.relu(out)
    return out if fixed else out * _prefix_mask(maxd, d)


def efc(sd, pfx, sp, s, maxs, ln, fixed):
    """ElasticLinear3D, modules.py:212-235."""
    out = _lin(sd, pfx + "._linear", sp.transpose(1, 2))
    if ln:
        out = _ln(sd, pfx + "._layernorm", out)
    out = torch.relu(out)
    if not fixed:
        out = out * _prefix_mask(maxs, s)
    return out.transpose(1, 2)


def dot_product(sd, pfx, dense, sp, d, maxd, ln, fixed):
    """DotProduct, modules.py:321-401."""
    if dense.shape[-1] != EMB_DIM:
        x = _lin(sd, pfx + "._dense_proj", dense)
        if ln:
            x = _ln(sd, pfx + "._dense_layernorm", x)
    else:
        x = dense
    y = sp                                   # last dim is always EMB_DIM (modules.py:348-354)
    P = round(math.sqrt(2 * maxd))           # modules.py:298
    if y.shape[1] != P:
        y = _lin(sd, pfx + "._sparse_inp_proj", y.transpose(1, 2))
        if ln:
            y = _ln(sd, pfx + "._sparse_inp_proj_layernorm", y)
        y = y.transpose(1, 2)
    T = torch.cat([x.unsqueeze(1), y], 1)
    Z = torch.bmm(T, T.transpose(1, 2))
    n = Z.shape[1]
    li, lj = torch.tril_indices(n, n, offset=-1)      # modules.py:375-379
    R = Z[:, li, lj]
    out = _lin(sd, pfx + "._linear_proj", R) if R.shape[-1] != maxd else R
    if ln:
        out = _ln(sd, pfx + "._linear_layernorm", out)
    return out if fixed else out * _prefix_mask(maxd, d)


def sum_node(sd, pfx, left, right, d, maxd, ln, fixed):
    """Sum, modules.py:458-501."""
    left, right = _pad_equal(left, right)
    out = left + right
    if out.shape[-1] != maxd:
        out = _lin(sd, pfx + "._linear_proj", out)
    if ln:
        out = _ln(sd, pfx + "._layernorm", out)
    return out if fixed else out * _prefix_mask(maxd, d)


def sigmoid_gating(sd, pfx, left, right, d, maxd, ln, fixed):
    """SigmoidGating + LazySelfLinear, modules.py:504-595."""
    left, right = _pad_equal(left, right)
    g = torch.sigmoid(_lin(sd, pfx + "._left_self_linear._linear", left))
    out = g * right
    if out.shape[-1] != maxd:
        out = _lin(sd, pfx + "._linear_proj", out)
    if ln:
        out = _ln(sd, pfx + "._layernorm", out)
    return out if fixed else out * _prefix_mask(maxd, d)


def mha_self(sd, pfx, x):
    """nn.MultiheadAttention(16, 8, batch_first=True) self-attention, no masks,
    need_weights=False (modules.py:624,664).  Published algorithm of
    torch.nn.functional.multi_head_attention_forward: packed in-proj, per-head
    softmax(q k^T / sqrt(hd)) v, out-proj."""
    B, L, E = x.shape
    hd = E // MHA_HEADS
    qkv = F.linear(x, sd[pfx + ".in_proj_weight"], sd[pfx + ".in_proj_bias"])
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(B, L, MHA_HEADS, hd).transpose(1, 2)
    k = k.reshape(B, L, MHA_HEADS, hd).transpose(1, 2)
    v = v.reshape(B, L, MHA_HEADS, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, L, E)
    return F.linear(o, sd[pfx + ".out_proj.weight"], sd[pfx + ".out_proj.bias"])


def transformer(sd, pfx, sp, s, maxs, ln, fixed):
    """Transformer, modules.py:642-688."""
    p = _lin(sd, pfx + "._linear_proj", sp.transpose(1, 2))
    if ln:
        p = _ln(sd, pfx + "._proj_ln", p)
    if not fixed:
        p = p * _prefix_mask(maxs, s)
    p = p.transpose(1, 2)                                    # [B, tokens, 16]
    a = _ln(sd, pfx + "._attn_ln", mha_self(sd, pfx + "._mha", p) + p)
    f = _lin(sd, pfx + ".attn_fc2", torch.relu(_lin(sd, pfx + ".attn_fc1", a)))
    o = _ln(sd, pfx + "._attn_fc_ln", a + f)
    if not fixed:
        o = (o.transpose(1, 2) * _prefix_mask(maxs, s)).transpose(1, 2)
    return o


def fm3d(sd, pfx, sp, d, maxd, ln, fixed):
    """FactorizationMachine3D, modules.py:733-750."""
    ix = sp.sum(1) ** 2 - (sp ** 2).sum(1)
    if ix.shape[-1] != maxd:
        ix = _lin(sd, pfx + "._linear_proj", ix)
        if ln:
            ix = _ln(sd, pfx + "._linear_layernorm", ix)
    return ix if fixed else ix * _prefix_mask(maxd, d)


# --------------------------------------------------------------------------
# choice block and supernet
# --------------------------------------------------------------------------
def _as_list(v):
    return [int(x) for x in np.asarray(v).reshape(-1).tolist()]


def block_forward(sd, bi, ops, ln, fixed, micro, dense, sp, left, right):
    """SuperNetBlock.forward / fixed_forward, supernet.py:1067-1162, 1185-1242."""
    names = ops["node_names"]
    active = _as_list(micro["active_nodes"])
    d, s = int(micro["dense_in_dims"]), int(micro["sparse_in_dims"])
    maxd = d if fixed else max(ops["dense_node_dims"])
    maxs = s if fixed else max(ops["sparse_node_dims"])
    B = dense.shape[0]
    out2d, out3d = [], []
    for i, name in enumerate(names):
        pfx = "_blocks.%d._nodes.%d" % (bi, i)
        if i not in active:
            if fixed:
                continue
            if name in SPARSE_NODES:                         # supernet.py:1096-1111
                out3d.append(torch.zeros(B, maxs, sp.shape[2]))
            else:                                            # supernet.py:1084-1094
                out2d.append(torch.zeros(B, maxd))
            continue
        if name == "linear-2d":
            out2d.append(fc(sd, pfx, dense, d, maxd, ln, fixed))
        elif name == "zeros-2d":                             # modules.py:252-270
            out2d.append(torch.zeros(B, maxd if not fixed else d))
        elif name == "dot-product":
            out2d.append(dot_product(sd, pfx, dense, sp, d, maxd, ln, fixed))
        elif name == "sum":
            out2d.append(sum_node(sd, pfx, left, right, d, maxd, ln, fixed))
        elif name == "sigmoid-gating":
            out2d.append(sigmoid_gating(sd, pfx, left, right, d, maxd, ln, fixed))
        elif name == "transformer":
            out3d.append(transformer(sd, pfx, sp, s, maxs, ln, fixed))
        elif name == "linear-3d":
            out3d.append(efc(sd, pfx, sp, s, maxs, ln, fixed))
        elif name == "zeros-3d":                             # modules.py:703-718
            out3d.append(torch.zeros(B, maxs, sp.shape[2]))
        else:
            raise NotImplementedError(name)
    dense_out = torch.stack(out2d, -1).sum(-1)               # supernet.py:1133
    sparse_out = torch.stack(out3d, -1).sum(-1)              # supernet.py:1134
    dsi, dfm = int(micro["dense_sparse_interact"]), int(micro["deep_fm"])
    pb = "_blocks.%d" % bi
    aliased = False
    if dsi == 1:
        if dense_out.shape[-1] != EMB_DIM * DS_SPLITS:       # supernet.py:1138-1142
            proj = _lin(sd, pb + ".project_emb_dim", dense_out)
            if ln:
                proj = _ln(sd, pb + ".project_emb_dim_layernorm", proj)
        else:                                                # supernet.py:1143-1145, 1224-1226
            proj = dense_out                                 # same storage; sees the += below
            aliased = True
    if dfm == 1:                                             # supernet.py:1154-1157
        dense_out = dense_out + fm3d(sd, pb + ".deep_fm", sparse_out, d, maxd, ln, fixed)
        if aliased:
            proj = dense_out
    if dsi == 1:
        proj = proj.reshape(-1, DS_SPLITS, EMB_DIM)
        sparse_out = torch.cat([sparse_out, proj], 1)
    elif not fixed:                                          # supernet.py:1147-1150,1161
        sparse_out = torch.cat([sparse_out, torch.zeros(B, DS_SPLITS, EMB_DIM)], 1)
    return dense_out, sparse_out                             # fixed & dsi==0: no concat (:1241-1242)


def supernet_forward(sd: Dict[str, torch.Tensor], cfg: Dict[str, Any], choice: Dict[str, Any],
                     int_x: torch.Tensor, cat_x: torch.Tensor) -> torch.Tensor:
    """SuperNet.forward (weight sharing, supernet.py:513-602) and
    SuperNet.fixed_forward (supernet.py:605-668).  cfg keys: ops ("xlarge"|...),
    use_layernorm, fixed, num_blocks."""
    ops = OPS_CONFIG[cfg["ops"]]
    ln, fixed, nb = bool(cfg["use_layernorm"]), bool(cfg["fixed"]), int(cfg["num_blocks"])
    Fn = cat_x.shape[1]
    sp0 = torch.stack([F.embedding(cat_x[:, f], sd["_embedding.%d.weight" % f]) for f in range(Fn)], 1)
    dense_list, sparse_list = [int_x], [sp0]
    for i in range(nb):
        mac = choice["macro"][i]
        sel = {k: set(_as_list(mac[k])) for k in ("dense_idx", "sparse_idx", "dense_left_idx", "dense_right_idx")}
        d_in, s_in, l_in, r_in = [], [], [], []
        for j in range(len(dense_list)):     # ascending j in both modes (supernet.py:536-568, 625-633)
            for key, lst, src in (("dense_idx", d_in, dense_list), ("sparse_idx", s_in, sparse_list),
                                  ("dense_left_idx", l_in, dense_list), ("dense_right_idx", r_in, dense_list)):
                if j in sel[key]:
                    lst.append(src[j])
                elif not fixed:
                    lst.append(torch.zeros_like(src[j]))
        dense = torch.cat(d_in, -1)
        sp = torch.cat(s_in, 1)
        left = torch.cat(l_in, -1)
        right = torch.cat(r_in, -1)
        do, so = block_forward(sd, i, ops, ln, fixed, choice["micro"][i], dense, sp, left, right)
        dense_list.append(do)
        sparse_list.append(so)
    feats = torch.cat([dense_list[-1], sparse_list[-1].flatten(1)], -1)     # supernet.py:592-597
    return _lin(sd, "_final", feats)


def full_path_choice(cfg) -> Dict[str, Any]:
    """supernet.py:814-824 and 1265-1276."""
    ops = OPS_CONFIG[cfg["ops"]]
    nb = int(cfg["num_blocks"])
    macro = [{k: list(range(i + 1)) for k in ("dense_idx", "sparse_idx", "dense_left_idx", "dense_right_idx")}
             for i in range(nb)]
    micro = [dict(active_nodes=list(range(ops["num_nodes"])), dense_in_dims=max(ops["dense_node_dims"]),
                  sparse_in_dims=max(ops["sparse_node_dims"]), dense_sparse_interact=1, deep_fm=1)
             for _ in range(nb)]
    return {"macro": macro, "micro": micro}


# --------------------------------------------------------------------------
# step body (train_utils.py:262-286) with the reference's dense optimizer
# --------------------------------------------------------------------------
class OracleTrainer:
    """zero_grad -> forward -> BCEWithLogits -> backward -> clip_grad_norm_(5.0)
    -> torch.optim.Adagrad(eps=1e-2).step(), all params dense (every embedding
    row is touched by the optimizer, as in train_supernet.py:121-123)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg, lr: float, clip: float = 5.0, eps: float = 1e-2):
        self.cfg = cfg
        self.params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
        self.opt = torch.optim.Adagrad(list(self.params.values()), lr=lr, eps=eps)
        self.clip = clip

    def step(self, choice, int_x, cat_x, y):
        self.opt.zero_grad()
        logits = supernet_forward(self.params, self.cfg, choice, int_x, cat_x)
        loss = F.binary_cross_entropy_with_logits(logits, y)
        loss.backward()
        total = torch.nn.utils.clip_grad_norm_(list(self.params.values()), self.clip)
        self.opt.step()
        return logits.detach(), float(loss.detach()), float(total)

    @torch.no_grad()
    def forward(self, choice, int_x, cat_x):
        return supernet_forward(self.params, self.cfg, choice, int_x, cat_x)


def finetune_last_only(sd, cfg, choice, train_batches, eval_batches, lr: float, clip: float = 5.0,
                       eps: float = 1e-2, min_lr: float = 1e-8):
    """The EA's per-candidate recipe (eval_subnet_from_supernet.py:118-200 with
    --finetune_whole_supernet 0; loop train_utils.py:262-300,386): only ``_final`` trains
    (supernet.py:850-853), Adagrad(eps=1e-2) + clip 5.0, cosine schedule with warm-up = steps // 10
    positioned by step(epoch=-1) (lr_schedule.py:98-159), one scheduler step after every batch but
    the last.  Returns (per-step losses, per-step lrs, tuned final weight/bias, eval logits)."""
    import math
    steps = len(train_batches)
    warm = steps // 10
    params = {k: v.clone() for k, v in sd.items()}
    fw = torch.nn.Parameter(params["_final.weight"])
    fb = torch.nn.Parameter(params["_final.bias"])
    params["_final.weight"], params["_final.bias"] = fw, fb
    opt = torch.optim.Adagrad([fw, fb], lr=lr, eps=eps)

    def lr_at(pos):                                  # lr_schedule.py:98-120 with base_lr == min_lr
        if pos == -1:
            return min_lr
        if pos < warm:
            return (lr - min_lr) * pos / warm + min_lr
        return min_lr + (lr - min_lr) * (1 + math.cos(math.pi * (pos - warm) / (steps - warm))) / 2

    losses, lrs = [], []
    for b, (int_x, cat_x, y) in enumerate(train_batches):
        cur = lr_at(b - 1)
        for g in opt.param_groups:
            g["lr"] = cur
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(supernet_forward(params, cfg, choice, int_x, cat_x), y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([fw, fb], clip)
        opt.step()
        losses.append(float(loss.detach()))
        lrs.append(cur)
    with torch.no_grad():
        outs = [supernet_forward(params, cfg, choice, bx[0], bx[1]) for bx in eval_batches]
    return losses, lrs, fw.detach(), fb.detach(), (torch.cat(outs) if outs else None)


def loss_and_grads(sd, cfg, choice, int_x, cat_x, y):
    """Logits, BCE loss and dense grads of every tensor that takes part."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits = supernet_forward(params, cfg, choice, int_x, cat_x)
    loss = F.binary_cross_entropy_with_logits(logits, y)
    names = list(params)
    grads = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    return logits.detach(), loss.detach(), {n: g for n, g in zip(names, grads) if g is not None}


def input_transform(ints: np.ndarray, hex_cols: Sequence[Sequence[str]], num_embeddings: Sequence[int],
                    zero_dense: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """VanillaTransform{Criteo,KDD,Avazu} (data_pipes.py:135-252) on raw columns:
    dense log(max(0,x)+1) in fp32 (Avazu: zeros); id = int(v,16) (or -1 when empty) .fmod(N_f-1) + 1."""
    x = np.asarray(ints, dtype=np.int64)
    if zero_dense:
        int_x = np.zeros(x.shape, dtype=np.float32)
    else:
        int_x = torch.log(torch.from_numpy(np.maximum(x, 0) + 1)).numpy().astype(np.float32)
    B = len(hex_cols[0]) if len(hex_cols) else x.shape[0]
    cat_x = np.zeros((B, len(hex_cols)), dtype=np.int64)
    for f, col in enumerate(hex_cols):
        raw = np.asarray([int(v, 16) if v else -1 for v in col], dtype=np.int64)
        cat_x[:, f] = np.fmod(raw, num_embeddings[f] - 1) + 1
    return int_x, cat_x


def binary_metrics(logits: np.ndarray, y: np.ndarray) -> Tuple[float, float, float]:
    """accuracy@0.5 on sigmoid, ROC-AUC (rank statistic with average ranks for
    ties == sklearn.metrics.roc_auc_score), BCE log-loss.  train_utils.py:158-178."""
    z = logits.astype(np.float64).reshape(-1)
    t = y.astype(np.float64).reshape(-1)
    p = 1.0 / (1.0 + np.exp(-z))
    acc = float(((p > 0.5) == (t > 0.5)).mean())
    loss = float(np.mean(np.maximum(z, 0) - z * t + np.log1p(np.exp(-np.abs(z)))))
    order = np.argsort(p, kind="mergesort")
    ps = p[order]
    ranks = np.empty_like(ps)
    i = 0
    n = len(ps)
    while i < n:
        j = i
        while j + 1 < n and ps[j + 1] == ps[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty_like(ranks)
    r[order] = ranks
    npos = t.sum()
    nneg = n - npos
    auc = float((r[t > 0.5].sum() - npos * (npos + 1) / 2.0) / (npos * nneg)) if npos > 0 and nneg > 0 else float("nan")
    return acc, auc, loss
