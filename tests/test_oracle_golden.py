"""CPU: pins oracle/nasrec_oracle.py to the reference's own outputs
(tests/golden/*, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import nasrec_oracle as orc
from tests.helpers import load_golden, rel_err

TOL = 2e-5   # fp32, same primitives (F.linear / layer_norm), different association in a few sums


def _check_case(cfg, sd, ne, nd, ds, choice, B, seed, logits_ref, loss_ref, gn_ref, small, emb_rows):
    int_x, cat_x, y = orc.synth_batch(B, nd, ne, seed=seed, all_zero_dense=(ds == "avazu"))
    logits, loss, grads = orc.loss_and_grads(sd, cfg, choice, int_x, cat_x, y)
    assert rel_err(logits.numpy(), logits_ref) < TOL
    assert abs(float(loss) - loss_ref) < TOL * max(1.0, abs(loss_ref))
    # every tensor the reference gave a gradient to, and only those (modulo exact-zero grads)
    for n, g in gn_ref.items():
        assert n in grads, n
        got = float(grads[n].double().norm())
        assert abs(got - g) <= 1e-4 * max(g, 1e-3) + 1e-7, (n, got, g)
    for n in grads:
        if n not in gn_ref:
            assert float(grads[n].abs().max()) == 0.0, n
    for n, g in small.items():
        assert rel_err(grads[n].numpy(), g) < 5e-4, n
    # bit-exact row sets of the embedding gradient
    sets = orc.embedding_row_sets(cat_x.numpy())
    for f, rows in emb_rows.items():
        got = np.nonzero(np.abs(grads["_embedding.%s.weight" % f].numpy()).sum(1))[0].tolist()
        assert got == rows
        assert set(rows) <= set(sets[int(f)].tolist())


@pytest.mark.parametrize("name", ["supernet_autoctr_criteo", "supernet_xlarge_criteo", "supernet_xlarge_kdd",
                                  "supernet_xlarge_avazu", "supernet_zeros_criteo"])
def test_oracle_matches_reference_supernet(name):
    meta, arr = load_golden(name)
    sd = orc.fill_state_dict(meta["shapes"], meta["state_seed"])
    for ci, case in enumerate(meta["cases"]):
        small = {k.split("/", 1)[1]: v for k, v in arr.items() if k.startswith("grad_%d/" % ci)}
        _check_case(meta["cfg"], sd, meta["num_embeddings"], meta["nd"], meta["dataset"], case["choice"],
                    meta["batch"], case["batch_seed"], arr["logits_%d" % ci], case["loss"], case["grad_norms"],
                    small, case["emb_rows"])


def test_oracle_matches_reference_fixed_best_models():
    meta, arr = load_golden("fixed_best")
    assert meta["models"]["criteo_xlarge"]["dense_params"] == 2217345 or meta["models"]["criteo_xlarge"]["dense_params"] > 2.2e6
    for tag, m in meta["models"].items():
        sd = orc.fill_state_dict(m["shapes"], m["state_seed"])
        small = {k.split("/", 2)[2]: v for k, v in arr.items() if k.startswith("grad/%s/" % tag)}
        _check_case(m["cfg"], sd, m["num_embeddings"], m["nd"], m["dataset"], m["choice"], m["batch"],
                    m["batch_seed"], arr["logits/" + tag], m["loss"], m["grad_norms"], small, m["emb_rows"])


def test_oracle_training_steps_match_reference():
    meta, arr = load_golden("train_steps")
    for tag, run in meta["runs"].items():
        sd = orc.fill_state_dict(run["shapes"], run["state_seed"])
        tr = orc.OracleTrainer(sd, run["cfg"], lr=run["lr"])
        for si, ch in enumerate(run["choices"]):
            int_x, cat_x, y = orc.synth_batch(8, 13, run["num_embeddings"], seed=300 + si)
            logits, loss, total = tr.step(ch, int_x, cat_x, y)
            assert rel_err(logits.numpy(), arr["%s/logits_%d" % (tag, si)]) < 1e-4
            assert abs(loss - run["losses"][si]) < 1e-4
            assert abs(total - run["total_norms"][si]) < 1e-3 * max(1.0, run["total_norms"][si])
        assert rel_err(tr.params["_final.weight"].detach().numpy(), arr[tag + "/final_weight"]) < 1e-4
        assert rel_err(tr.params["_embedding.0.weight"].detach().numpy(), arr[tag + "/emb0"]) < 1e-4


def test_binary_metrics_against_sklearn():
    from sklearn.metrics import roc_auc_score
    rs = np.random.RandomState(0)
    z = np.round(rs.randn(500), 1)          # rounding creates ties
    y = (rs.rand(500) < 0.3).astype(np.float32)
    acc, auc, loss = orc.binary_metrics(z, y)
    assert abs(auc - roc_auc_score(y, 1 / (1 + np.exp(-z)))) < 1e-12
    ref = torch.nn.functional.binary_cross_entropy_with_logits(torch.tensor(z), torch.tensor(y, dtype=torch.float64))
    assert abs(loss - float(ref)) < 1e-9


def test_embedding_integer_path():
    rs = np.random.RandomState(1)
    tables = [rs.randn(n, 16).astype(np.float32) for n in (7, 3, 50)]
    cat = np.stack([rs.randint(0, n, 33) for n in (7, 3, 50)], 1)
    out = orc.embedding_gather(tables, cat)
    t = torch.stack([torch.nn.functional.embedding(torch.from_numpy(cat[:, f]), torch.from_numpy(tables[f]))
                     for f in range(3)], 1)
    assert np.array_equal(out, t.numpy())
    g = rs.randn(33, 3, 16).astype(np.float32)
    for f, (rows, acc) in enumerate(orc.embedding_grad_rows(cat, g)):
        assert np.array_equal(rows, np.unique(cat[:, f]))
        dense = np.zeros_like(tables[f]); np.add.at(dense, cat[:, f], g[:, f])
        assert np.allclose(dense[rows], acc, atol=1e-5)
