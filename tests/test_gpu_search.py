"""GPU: the EA's per-candidate recipe (last-layer fine-tune + scoring) through the CUDA path
against the unmodified reference's recorded run (tests/golden/ea_finetune.*), and the
in-process Searcher end to end on a resident supernet."""
import numpy as np
import pytest
import torch

from nasrec_b200 import SuperNet, ops_config_lib
from nasrec_b200.search import Searcher, SubnetEvaluator, Tokenizer
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _resident(G):
    cfg, ne = G["cfg"], G["num_embeddings"]
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib[cfg["ops"]], use_layernorm=True, num_embeddings=ne,
                 sparse_input_size=len(ne), path_sampling_strategy="full-path").to("cuda")
    m.materialize(13)
    sd = orc.fill_state_dict({k: tuple(v) for k, v in G["shapes"].items()}, G["state_seed"])
    m.load_state_dict(sd, strict=True)
    return m, sd


def _batches(G):
    ne = G["num_embeddings"]
    tr = [orc.synth_batch(G["train_seeds"][1], 13, ne, seed=G["train_seeds"][0] + b) for b in range(G["steps"])]
    ev = [orc.synth_batch(G["eval_seeds"][1], 13, ne, seed=G["eval_seeds"][0] + b) for b in range(G["eval_seeds"][2])]
    cu = lambda bs: [tuple(t.cuda() for t in b) for b in bs]
    return cu(tr), cu(ev)


@pytest.mark.parametrize("trunk_samples", [8192, 40, 1])
def test_finetune_and_score_matches_reference_golden(trunk_samples):
    """trunk_samples only regroups the frozen trunk's launches (all 12 batches at once, 2 at a
    time, one by one): the optimizer walk and its result must not change."""
    G, A = load_golden("ea_finetune")
    m, sd = _resident(G)
    tr, ev = _batches(G)
    ev_obj = SubnetEvaluator(m)
    w0 = m._final.weight.detach().clone()
    for ci, c in enumerate(G["cands"]):
        r = ev_obj.finetune_and_score(c["choice"], tr, ev, lr=G["lr"], trunk_samples=trunk_samples, restore=False)
        assert np.abs(np.asarray(r["train_loss"]) - np.asarray(c["losses"])).max() < 2e-5
        assert np.abs(r["final_weight"].cpu().numpy() - A["cand%d/final_weight" % ci]).max() < 1e-5
        assert np.abs(r["final_bias"].cpu().numpy() - A["cand%d/final_bias" % ci]).max() < 1e-5
        assert abs(r["test_loss"] - c["test_loss"]) < 1e-5
        z = torch.cat([ev_obj.logits(c["choice"], b[0], b[1]).reshape(-1) for b in ev]).cpu().numpy()
        assert np.abs(z - A["cand%d/eval_logits" % ci].reshape(-1)).max() < 3e-5
        with torch.no_grad():                      # next candidate starts from the checkpoint again
            m._final.weight.copy_(sd["_final.weight"].cuda())
            m._final.bias.copy_(sd["_final.bias"].cuda())
    # restore=True leaves the resident supernet untouched
    ev_obj.finetune_and_score(G["cands"][0]["choice"], tr, ev, lr=G["lr"])
    assert torch.equal(m._final.weight, w0)
    # nothing but _final may have moved
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_searcher_end_to_end_on_gpu():
    G, _ = load_golden("ea_finetune")
    m, sd = _resident(G)
    tr, ev = _batches(G)
    tok = Tokenizer(7, ops_config_lib["xlarge"])
    s = Searcher(SubnetEvaluator(m), tok, tr[:4], ev, lr=G["lr"])
    np.random.seed(11)
    hist = s.regularized_evolution_from_supernet(n_generations=2, n_childs=3, init_population=5, sample_size=3,
                                                 criterion="test_loss", top_k=1)
    assert len(hist) == 2 and len(s.all_results) == 5
    for r in hist + list(s.all_results):
        assert set(r) == {"choice", "test_acc", "test_auroc", "test_loss", "hash_token"}
        assert np.isfinite(r["test_loss"]) and 0.0 <= r["test_auroc"] <= 1.0
        # the record is reproducible from its choice alone: oracle recipe on the same batches
    r = hist[0]
    host = lambda bs: [tuple(t.cpu() for t in b) for b in bs]
    _, _, _, _, z = orc.finetune_last_only(sd, G["cfg"], r["choice"], host(tr[:4]), host(ev), G["lr"])
    ys = torch.cat([b[2] for b in host(ev)])
    acc, auc, loss = orc.binary_metrics(z.numpy(), ys.numpy())
    assert abs(loss - r["test_loss"]) < 1e-5 and abs(auc - r["test_auroc"]) < 1e-4
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_reference_training_loop_drop_in_matches_golden():
    """The reference's own per-candidate flow (eval_subnet_from_supernet.py:118-200) written against
    the mirrored API -- stock torch Adagrad, the mirrored cosine scheduler and
    train_and_test_one_epoch -- reproduces the run recorded from the unmodified reference."""
    from nasrec_b200.utils import train_utils as tu
    from nasrec_b200.utils.lr_schedule import CosineAnnealingWarmupRestarts
    G, A = load_golden("ea_finetune")
    m, sd = _resident(G)
    ne, steps, lr = G["num_embeddings"], G["steps"], G["lr"]
    tr = [orc.synth_batch(G["train_seeds"][1], 13, ne, seed=G["train_seeds"][0] + b) for b in range(steps)]
    ev = [orc.synth_batch(G["eval_seeds"][1], 13, ne, seed=G["eval_seeds"][0] + b) for b in range(G["eval_seeds"][2])]
    c = G["cands"][1]
    m.configure_choice(c["choice"])
    m.configure_path_sampling_strategy("fixed-path")
    m.set_mode_to_finelune_last_only()
    opt = torch.optim.Adagrad(m.parameters(), lr=lr, eps=1e-2)
    sch = CosineAnnealingWarmupRestarts(opt, first_cycle_steps=steps, warmup_steps=steps // 10, max_lr=lr, min_lr=1e-8)
    sch.step(epoch=-1)
    logs = tu.train_and_test_one_epoch(m, 0, opt, sch, tr, ev, torch.nn.BCEWithLogitsLoss(),
                                       lambda mod: tu.get_l2_loss(mod, 0, None, gpu="cuda"), G["train_seeds"][1], "cuda",
                                       display_interval=1, max_train_steps=steps, max_eval_steps=len(ev),
                                       test_interval=max(2, steps), test_only_at_last_step=True, grad_clip_value=5.0)
    assert np.abs(np.asarray(logs["train_loss"]) - np.asarray(c["losses"])).max() < 2e-5
    assert logs["iters"] == list(range(steps)) and len(logs["test_loss"]) == 1
    assert abs(logs["test_loss"][0] - c["test_loss"]) < 1e-5
    assert np.abs(m._final.weight.detach().cpu().numpy() - A["cand1/final_weight"]).max() < 1e-5
    assert 0.0 <= logs["test_AUROC"][0] <= 1.0 and 0.0 <= logs["test_Accuracy"][0] <= 1.0


def _xlarge_resident(seed=31):
    meta, _ = load_golden("supernet_xlarge_criteo")
    cfg, ne, nd = meta["cfg"], meta["num_embeddings"], meta["nd"]
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=ne,
                 sparse_input_size=len(ne), path_sampling_strategy="full-path").to("cuda")
    m.materialize(nd)
    sd = orc.fill_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, seed)
    m.load_state_dict(sd, strict=True)
    m.requires_grad_(False)
    return m, sd, cfg, ne, nd


def test_multi_subnet_eval_matches_oracle_per_candidate():
    """nasrec_multi_subnet_eval (searcher_utils.py:57-104, eval_subnet_from_supernet.py:182-198): 16 NASRec-Full candidates
    x 4 evaluation batches in batched calls == the oracle's logits candidate by candidate (1e-5), bit-identical to the
    one-candidate-at-a-time executor path, and the same log-loss / AUC / accuracy records."""
    from nasrec_b200.native import NativeNet
    from nasrec_b200.search import generate_random_choice
    m, sd, cfg, ne, nd = _xlarge_resident()
    np.random.seed(5)
    cands = [generate_random_choice(7, ops_config_lib["xlarge"]) for _ in range(16)]
    host = [orc.synth_batch(96, nd, ne, seed=700 + b, zipf=True) for b in range(4)]
    dev = [tuple(t.cuda() for t in b) for b in host]
    ev = SubnetEvaluator(m, group=16)
    recs = ev.score(cands, dev)
    assert ev.multi_stats[0] > 0, "the batched path was not taken"
    net = ev._native()
    ys = torch.cat([b[2].reshape(-1) for b in host]).numpy()
    for ci, ch in enumerate(cands):
        z_ref = torch.cat([orc.supernet_forward(sd, cfg, ch, b[0], b[1]).reshape(-1) for b in host]).detach().numpy()
        enc = NativeNet.encode_choice(ch["macro"], ch["micro"])
        z_one = torch.cat([net.forward(enc, b[0].contiguous(), b[1].contiguous()).reshape(-1) for b in dev])
        z_multi = torch.cat([net.forward_multi([enc], b[0].contiguous(), b[1].contiguous()).reshape(-1) for b in dev])
        assert torch.equal(z_one, z_multi), ci
        z = z_one.cpu().numpy()
        assert np.abs(z - z_ref).max() <= 1e-5 * max(1.0, np.abs(z_ref).max()), ci
        acc, auc, loss = orc.binary_metrics(z_ref, ys)
        assert abs(recs[ci]["test_loss"] - loss) < 1e-5 and abs(recs[ci]["test_auroc"] - auc) < 1e-6
        assert abs(recs[ci]["test_acc"] - acc) < 1e-6


def test_multi_subnet_eval_shares_blocks_within_an_ea_generation():
    """The children of one generation differ from their parent in one field of one block (tokenizer.py:192-265): the
    batched path computes the blocks they share once, and every child's logits stay bit-identical to its own forward."""
    from nasrec_b200.native import NativeNet
    m, sd, cfg, ne, nd = _xlarge_resident(seed=32)
    tok = Tokenizer(7, ops_config_lib["xlarge"])
    np.random.seed(9)
    parent = tok.generate_random_choice()
    children = [tok.mutate_spec(parent) for _ in range(8)]
    int_x, cat_x, _ = (t.cuda() for t in orc.synth_batch(130, nd, ne, seed=801, zipf=True))
    net = NativeNet(m, state_of=None, pgrad_bytes=1 << 20)
    encs = [NativeNet.encode_choice(c["macro"], c["micro"]) for c in [parent] + children]
    z = net.forward_multi(encs, int_x.contiguous(), cat_x.contiguous())
    computed, reused = net.last_multi_stats
    assert reused > 0 and computed < 9 * 7, (computed, reused)
    for k, e in enumerate(encs):
        assert torch.equal(z[k], net.forward(e, int_x.contiguous(), cat_x.contiguous()).reshape(-1)), k
