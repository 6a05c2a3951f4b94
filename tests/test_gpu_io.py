"""GPU: device-side evaluation metrics and raw-batch transform
through the C ABI, against the reference's recorded outputs and the oracle."""
import numpy as np
import pytest
import torch

from nasrec_b200.search import binary_metrics_device
from nasrec_b200.utils.data_pipes import InputTransform
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ds", ["criteo", "avazu", "kdd"])
def test_input_transform_matches_reference_golden(ds):
    G = load_golden("input_transform")[0][ds]
    tf = InputTransform(G["num_embeddings"], G["nd"], zero_dense=(ds == "avazu"))
    batch = {"label": torch.tensor(G["label"])}
    ints = np.asarray(G["ints"])
    for c in range(G["nd"]):
        batch["int_%d" % c] = torch.tensor(ints[:, c])
    for f in range(G["F"]):
        batch["cat_%d" % f] = G["hex"][f]
    int_x, cat_x, y = tf(batch)
    assert cat_x.dtype == torch.int64 and int_x.dtype == torch.float32 and y.shape == (len(G["label"]), 1)
    assert np.array_equal(cat_x.cpu().numpy(), np.asarray(G["cat_x"]))                   # bit-exact indices
    ref = np.asarray(G["int_x"], dtype=np.float32)
    assert np.abs(int_x.cpu().numpy() - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max())   # logf: 1 ulp
    assert np.array_equal(y.cpu().numpy(), np.asarray(G["y"], dtype=np.float32))


def test_input_transform_large_random_and_errors():
    rng = np.random.RandomState(5)
    B, F, nd = 4096, 26, 13
    ne = [int(v) for v in rng.randint(2, 10_000_000, size=F)]
    ne[3] = 3                                                                            # smallest useful table
    hexs = [["" if rng.rand() < 0.05 else "%x" % int(rng.randint(0, 1 << 32, dtype=np.int64)) for _ in range(B)]
            for _ in range(F)]
    hexs[0][0] = "ffffffffffffffe"                                                       # 15 digits: int64 edge
    ints = rng.randint(-5, 1 << 20, size=(B, nd))
    tf = InputTransform(ne, nd)
    int_x, cat_x, _ = tf.transform_columns([ints[:, c] for c in range(nd)], hexs)
    ri, rc = orc.input_transform(ints, hexs, ne)
    assert np.array_equal(cat_x.cpu().numpy(), rc)
    assert np.abs(int_x.cpu().numpy() - ri).max() <= 2e-6
    empty = tf.transform_columns([np.zeros(0)] * nd, [[] for _ in range(F)])
    assert empty[0].shape == (0, nd) and empty[1].shape == (0, F)
    bad = [list(c) for c in hexs]
    bad[2][7] = "12g4"
    with pytest.raises(ValueError):
        tf.transform_columns([ints[:, c] for c in range(nd)], bad)
    again = tf.transform_columns([ints[:, c] for c in range(nd)], hexs)                  # flag was reset
    assert np.array_equal(again[1].cpu().numpy(), rc)


@pytest.mark.parametrize("n,ties", [(1, False), (2, False), (257, False), (10_000, True), (300_001, True)])
def test_binary_metrics_matches_oracle(n, ties):
    g = torch.Generator().manual_seed(n)
    z = torch.randn(n, generator=g) * 2
    if ties:
        z = (z * 4).round() / 4                                      # heavy ties, incl. across labels
        z[: n // 10] = 40.0                                          # saturated sigmoid: all tie at p == 1
    y = (torch.rand(n, generator=g) < torch.sigmoid(z * 0.5)).float()
    if n > 1:
        y[0], y[1] = 0.0, 1.0
    acc, auc, loss = binary_metrics_device(z.cuda(), y.cuda())
    p32 = torch.sigmoid(z).numpy()                                   # the reference ranks fp32 sigmoid outputs
    racc, rauc, rloss = orc.binary_metrics(z.numpy(), y.numpy())
    assert abs(loss - rloss) < 1e-6 * max(1.0, abs(rloss))
    assert acc == pytest.approx(float(((p32 > 0.5) == (y.numpy() > 0.5)).mean()), abs=1e-12)
    if n == 1:
        assert np.isnan(auc)
        return
    # exact Mann-Whitney count.  Grouping is done on z: for these inputs sigmoid is injective on the
    # distinct z values (and all the saturated ones are equal), so p and z have the same tie sets --
    # whereas torch's CPU sigmoid may round the same z differently in its SIMD body and scalar tail.
    zz = z.numpy().astype(np.float64)
    order = np.argsort(zz, kind="mergesort")
    ps, ys = zz[order], y.numpy().astype(np.float64)[order]
    uniq, start, cnt = np.unique(ps, return_index=True, return_counts=True)
    negpre = np.concatenate([[0], np.cumsum(1 - ys)])
    twice = 0
    for s, c in zip(start, cnt):
        negs_in = negpre[s + c] - negpre[s]
        twice += int((c - negs_in) * (2 * negpre[s] + negs_in))
    P = float(ys.sum())
    want = twice / (2.0 * P * (n - P))
    assert auc == pytest.approx(want, abs=1e-12)
    assert abs(auc - rauc) < 1e-4                                    # oracle ranks fp64 sigmoids (tie sets differ)


def test_binary_metrics_equals_sklearn_definition_small():
    """Hand-checkable: 3 positives, 3 negatives, one cross-label tie."""
    z = torch.tensor([2.0, 1.0, 1.0, 0.0, -1.0, 3.0])
    y = torch.tensor([1.0, 1.0, 0.0, 0.0, 0.0, 1.0])
    acc, auc, loss = binary_metrics_device(z.cuda(), y.cuda())
    # pairs (pos, neg): 3*3 = 9; wins: 2.0>{1,0,-1}=3, 1.0>{0,-1}=2 + tie .5, 3.0>all=3 -> 8.5/9
    assert auc == pytest.approx(8.5 / 9.0, abs=1e-15)
    assert acc == pytest.approx(5.0 / 6.0, abs=1e-15)                # predictions p>0.5 <=> z>0: [1,1,1,0,0,1]
    ref = torch.nn.functional.binary_cross_entropy_with_logits(z, y).item()
    assert loss == pytest.approx(ref, rel=1e-6)
