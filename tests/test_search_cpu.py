"""CPU: the EA host logic around the hot path against
sequences recorded from the unmodified reference -- LR schedules, Tokenizer
(token / hash / mutation RNG order), the oracle's last-layer fine-tune recipe,
and the in-process Searcher (single rank and sharded over 2 gloo ranks)."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nasrec_b200.search import Searcher, Tokenizer, draw_fixed_path_candidate
from nasrec_b200.supernet.supernet import ops_config_lib
from nasrec_b200.utils.lr_schedule import (ConstantWithWarmup, CosineAnnealingWarmupRestarts, CosineCursor,
                                           finetune_lr_sequence)
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden


def _opt(lr):
    return torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)


@pytest.mark.parametrize("tag", ["cosine_20_w2", "cosine_mult2_gamma"])
def test_cosine_schedule_matches_reference_sequence(tag):
    G = load_golden("lr_schedules")[0][tag]
    o = _opt(G["opt_lr"])
    s = CosineAnnealingWarmupRestarts(o, **G["kwargs"])
    seq = [o.param_groups[0]["lr"]]
    s.step(epoch=-1)
    seq.append(o.param_groups[0]["lr"])
    for _ in range(len(G["seq"]) - 2):
        s.step()
        seq.append(o.param_groups[0]["lr"])
    assert seq == G["seq"]                                  # host scalars: exact
    for e, lr, cyc, pos in G["seeks"]:
        s.step(epoch=e)
        assert (o.param_groups[0]["lr"], s.cycle, s.step_in_cycle) == (lr, cyc, pos)
    # the optimizer-free cursor walks the same positions
    c = CosineCursor(**G["kwargs"])
    c.advance()
    assert c.seek(-1) == G["seq"][1]
    assert [c.advance() for _ in range(len(G["seq"]) - 2)] == G["seq"][2:]


def test_constant_warmup_matches_reference_sequence():
    G = load_golden("lr_schedules")[0]["constant_w5"]
    o = _opt(G["opt_lr"])
    s = ConstantWithWarmup(o, num_warmup_steps=5)
    seq = [o.param_groups[0]["lr"]]
    for _ in range(12):
        o.step()
        s.step()
        seq.append(o.param_groups[0]["lr"])
    assert seq == G["seq"]


def test_finetune_lr_sequence_is_what_the_reference_loop_applies():
    G = load_golden("ea_finetune")[0]
    assert finetune_lr_sequence(G["steps"], G["lr"]) == G["cands"][0]["lrs"]
    assert finetune_lr_sequence(G["steps"], G["lr"])[0] == 1e-8        # first step at min_lr


@pytest.mark.parametrize("ops", ["xlarge", "autoctr"])
def test_tokenizer_matches_reference(ops):
    G = load_golden("tokenizer")[0][ops]
    tok = Tokenizer(7, ops_config_lib[ops])
    np.random.seed(4321)
    cands = [tok.generate_random_choice() for _ in range(3)]
    assert cands == G["cands"]
    assert [tok.tokenize(c).tolist() for c in cands] == G["tokens"]
    assert [tok.hash_token(tok.tokenize(c)) for c in cands] == G["hashes"]
    np.random.seed(99)
    cur = cands[0]
    for step in G["mutations_seed99"]:
        nxt = tok.mutate_spec(cur)
        assert nxt == step["choice"]
        assert tok.hash_token(tok.tokenize(nxt)) == step["hash"]
        assert sum(a != b for a, b in zip(tok.tokenize(cur), tok.tokenize(nxt))) <= 8   # one field of one block
        cur = nxt
    # numpy-array valued (full-path) choices tokenise like lists
    arr = copy.deepcopy(cands[1])
    for mac in arr["macro"]:
        for k in mac:
            mac[k] = np.asarray(mac[k])
    assert tok.hash_token(tok.tokenize(arr)) == G["hashes"][1]


def test_oracle_finetune_last_only_matches_reference_golden():
    G, A = load_golden("ea_finetune")
    sd = orc.fill_state_dict({k: tuple(v) for k, v in G["shapes"].items()}, G["state_seed"])
    ne = G["num_embeddings"]
    tr = [orc.synth_batch(G["train_seeds"][1], 13, ne, seed=G["train_seeds"][0] + b) for b in range(G["steps"])]
    ev = [orc.synth_batch(G["eval_seeds"][1], 13, ne, seed=G["eval_seeds"][0] + b) for b in range(G["eval_seeds"][2])]
    for ci, c in enumerate(G["cands"]):
        losses, lrs, fw, fb, z = orc.finetune_last_only(sd, G["cfg"], c["choice"], tr, ev, G["lr"])
        assert lrs == c["lrs"]
        assert np.abs(np.asarray(losses) - np.asarray(c["losses"])).max() < 1e-5
        assert np.abs(fw.numpy() - A["cand%d/final_weight" % ci]).max() < 1e-5
        assert np.abs(fb.numpy() - A["cand%d/final_bias" % ci]).max() < 1e-5
        assert np.abs(z.numpy() - A["cand%d/eval_logits" % ci]).max() < 2e-5
        ys = torch.cat([b[2] for b in ev])
        assert abs(orc.binary_metrics(z.numpy(), ys.numpy())[2] - c["test_loss"]) < 1e-5


class _FakeModel:
    """Sampler-only stand-in for a resident supernet (no CUDA): what draw_fixed_path_candidate touches."""

    def __init__(self):
        from nasrec_b200 import SuperNet
        self.m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True,
                          num_embeddings=[10] * 26, sparse_input_size=26)


class _FakeEvaluator:
    """Deterministic 'loss' from the candidate's hash, so EA decisions can be replayed."""

    def __init__(self, model):
        self.model = model
        self.calls = 0

    def score(self, choices, batches, use_cuda_graph=False):
        out = []
        tok = Tokenizer(7, ops_config_lib["xlarge"])
        for ch in choices:
            self.calls += 1
            h = tok.hash_token(tok.tokenize(ch))
            v = (int(h, 2) % 100003) / 100003.0 if set(h) <= {"0", "1"} else (sum(map(int, h)) % 997) / 997.0
            out.append({"test_acc": 1 - v, "test_auroc": 0.5 + v / 4, "test_loss": 0.4 + v})
        return out


def _run_search(group=None):
    ev = _FakeEvaluator(_FakeModel().m)
    s = Searcher(ev, Tokenizer(7, ops_config_lib["xlarge"]), [], [], finetune=False, group=group)
    np.random.seed(5)
    top = s.random_search_from_supernet(budget=9, criterion="test_loss", top_k=3)
    np.random.seed(6)
    hist = s.regularized_evolution_from_supernet(n_generations=3, n_childs=4, init_population=8, sample_size=3,
                                                 criterion="test_auroc", top_k=2)
    return top, hist, s, ev


def test_searcher_random_and_evolution_single_rank():
    top, hist, s, ev = _run_search()
    assert len(top) == 3 and top[0]["test_loss"] <= top[1]["test_loss"] <= top[2]["test_loss"]
    assert set(top[0]) == {"choice", "test_acc", "test_auroc", "test_loss", "hash_token"}     # results.pickle schema
    assert len(hist) == 3 * 2 and len(s.all_results) == 8           # population size is conserved (aging)
    for a, b in zip(hist[0::2], hist[1::2]):
        assert a["test_auroc"] >= b["test_auroc"]                    # descending for accuracy-like criteria
    assert ev.calls == 9 + 8 + 3 * 4
    hashes = [r["hash_token"] for r in hist]
    assert len(set(hashes)) == len(hashes)                           # visited-hash rejection
    top2, hist2, _, _ = _run_search()
    assert [r["hash_token"] for r in top2] == [r["hash_token"] for r in top]        # same seeds, same search
    assert hashes == [r["hash_token"] for r in hist2]


def test_random_candidates_come_from_the_fixed_path_sampler():
    """The random phase draws through SuperNet's own fixed-path sampler (golden: samplers.json)."""
    meta, _ = load_golden("samplers")
    m = _FakeModel().m
    np.random.seed(9)
    got = draw_fixed_path_candidate(m)
    assert got == meta["fixed_path_seed9"]
    again = draw_fixed_path_candidate(m)                             # a second draw is a NEW candidate
    assert again != got


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        top, hist, s, ev = _run_search()
        q.put((rank, [r["hash_token"] for r in top], [r["hash_token"] for r in hist], ev.calls))
    finally:
        dist.destroy_process_group()


def test_searcher_sharded_over_two_ranks_equals_single_rank():
    top, hist, _, ev = _run_search()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
    for rank, t, h, calls in got:
        assert t == [r["hash_token"] for r in top]
        assert h == [r["hash_token"] for r in hist]
        assert calls < ev.calls                                       # each rank scored only its shard
    assert sum(g[3] for g in got) == ev.calls


# ------------------------------------------------------------------ input transform
@pytest.mark.parametrize("ds", ["criteo", "avazu", "kdd"])
def test_oracle_input_transform_matches_reference_golden(ds):
    G = load_golden("input_transform")[0][ds]
    int_x, cat_x = orc.input_transform(np.asarray(G["ints"]), G["hex"], G["num_embeddings"], zero_dense=(ds == "avazu"))
    assert np.array_equal(cat_x, np.asarray(G["cat_x"]))                     # bit-exact indices
    assert np.array_equal(int_x, np.asarray(G["int_x"], dtype=np.float32))
    missing = np.asarray([[v == "" for v in col] for col in G["hex"]]).T
    assert (cat_x[missing] == 0).all() and (cat_x[~missing] >= 1).all()     # missing -> row 0, ids in [1, N-1]
    assert (cat_x < np.asarray(G["num_embeddings"])[None, :]).all()


def test_pack_hex_columns_layout_and_limits():
    from nasrec_b200.utils.data_pipes import pack_hex_columns
    p = pack_hex_columns([["a", "", "0fF3"], ["12345678", "9", ""]])
    assert p.shape == (2, 3, 8) and p.dtype == np.uint8
    assert bytes(p[0, 0]) == b"a" + b"\0" * 7 and bytes(p[0, 1]) == b"\0" * 8 and bytes(p[1, 0]) == b"12345678"
    assert pack_hex_columns([["", ""]]).shape == (1, 2, 1)
    with pytest.raises(ValueError):
        pack_hex_columns([["0123456789abcdef"]])                              # 16 digits do not fit int64
    with pytest.raises(ValueError):
        pack_hex_columns([["a"], ["b", "c"]])
