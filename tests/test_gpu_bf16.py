"""GPU: bf16 compute mode (BASELINE config #4; the reference's --use_amp hooks, nasrec/utils/train_utils.py:146,247-286)
against the unmodified reference run under torch.autocast("cpu", dtype=bfloat16) (tests/golden/bf16_autocast.*).

STATED TOLERANCE.  bf16 keeps 8 significant bits; the reference itself moves by ~1e-2 (RMS, relative to the logit RMS)
when it is switched from fp32 to autocast (recorded in the fixture as fp32_vs_bf16_logit_rms).  The CUDA path rounds the
same GEMM operands to bf16 and accumulates in fp32, but does not round GEMM *outputs* to bf16 as autocast does, so it sits
between the two: logits within 2e-2 RMS-relative of the autocast reference, log-loss within 2e-2 relative, per-tensor
gradient norms within 10 %."""
import numpy as np
import pytest
import torch

import nasrec_b200
from nasrec_b200 import _lib
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden
from tests.test_gpu_supernet import _build, _run_case

pytestmark = pytest.mark.gpu

LOGIT_RMS_TOL = 2e-2
LOSS_TOL = 2e-2
GRAD_NORM_TOL = 1e-1


@pytest.mark.parametrize("tag", ["avazu_xlarge", "criteo_xlarge"])
def test_bf16_mode_matches_reference_autocast(tag):
    meta, arr = load_golden("bf16_autocast")
    mm = meta["models"][tag]
    m, _sd = _build(mm["cfg"], mm["num_embeddings"], mm["nd"], mm["shapes"], mm["state_seed"], mm["choice"])
    int_x, cat_x, y = orc.synth_batch(mm["batch"], mm["nd"], mm["num_embeddings"], seed=mm["batch_seed"],
                                      all_zero_dense=(mm["dataset"] == "avazu"))
    with nasrec_b200.precision("bf16"):
        assert _lib.LIB.gemm_mode() == 2
        logits, loss, grads = _run_case(m, mm["cfg"], mm["choice"], int_x, cat_x, y)
    assert _lib.LIB.gemm_mode() == 3
    ref = arr["logits/" + tag].reshape(-1)
    ref32 = arr["logits_fp32/" + tag].reshape(-1)
    rms = float(np.sqrt(np.mean(ref32.astype(np.float64) ** 2)))
    d = float(np.sqrt(np.mean((logits.numpy().reshape(-1).astype(np.float64) - ref) ** 2))) / rms
    assert d < LOGIT_RMS_TOL, d
    # and it really computed in reduced precision: it must differ from the fp32 reference by more than fp32 noise
    d32 = float(np.sqrt(np.mean((logits.numpy().reshape(-1).astype(np.float64) - ref32) ** 2))) / rms
    assert d32 > 1e-5, d32
    assert abs(loss - mm["loss"]) < LOSS_TOL * max(1.0, abs(mm["loss"]))
    for n, g in mm["grad_norms"].items():
        if g < 1e-6 or n.startswith("_embedding"):
            continue
        got = float(grads[n].double().norm())
        assert abs(got - g) <= GRAD_NORM_TOL * g + 1e-6, (n, got, g)


def test_bf16_native_training_steps_track_fp32():
    """The C++ executor in bf16 mode (weight planes rebuilt as rn_bf16(W) and kept in step by the optimizer): three training
    steps stay close to the same steps in fp32-parity mode and switching back restores bit-exact fp32 behaviour."""
    from nasrec_b200 import SuperNet, ops_config_lib
    from nasrec_b200.native import NativeTrainer
    meta, _ = load_golden("supernet_autoctr_criteo")
    cfg, ne, nd = meta["cfg"], meta["num_embeddings"], meta["nd"]

    def fresh():
        m = SuperNet(num_blocks=7, ops_config=ops_config_lib[cfg["ops"]], use_layernorm=True, num_embeddings=ne,
                     sparse_input_size=len(ne), path_sampling_strategy="full-path").to("cuda")
        m.materialize(nd)
        m.load_state_dict(orc.fill_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, 9), strict=True)
        m.configure_choice(meta["cases"][1]["choice"])
        m.configure_path_sampling_strategy("fixed-path")
        return m, NativeTrainer(m, lr=0.12)

    batches = [tuple(t.cuda() for t in orc.synth_batch(64, nd, ne, seed=60 + i)) for i in range(3)]
    a, ta = fresh()
    b, tb = fresh()
    c, tc = fresh()
    ref = [ta.step(*bt)[0].clone() for bt in batches]
    with nasrec_b200.precision("bf16"):
        got = [tb.step(*bt)[0].clone() for bt in batches]
    assert tb.net is not None
    for r, g in zip(ref, got):
        rms = float(r.pow(2).mean().sqrt())
        d = float((r - g).pow(2).mean().sqrt()) / rms
        assert 1e-6 < d < 5e-2, d
    # a trainer used only in fp32 mode after the switch back is bit-identical to one that never saw bf16
    again = [tc.step(*bt)[0].clone() for bt in batches]
    for r, g in zip(ref, again):
        assert torch.equal(r, g)
