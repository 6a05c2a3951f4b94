"""CPU: host-side parity of the drop-in boundary with the reference --
state-dict keys/shapes after lazy materialisation, RNG-order-exact samplers,
C-ABI library exports.  No kernels are launched."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from nasrec_b200 import SuperNet, ops_config_lib, _lib
from nasrec_b200.supernet.supernet import _ints
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _norm(choice):
    return {"macro": [{k: _ints(v) for k, v in m.items()} for m in choice["macro"]],
            "micro": [{k: (_ints(v) if k == "active_nodes" else int(v)) for k, v in m.items()}
                      for m in choice["micro"]]}


@pytest.mark.parametrize("name,ops", [("supernet_xlarge_criteo", "xlarge"), ("supernet_autoctr_criteo", "autoctr"),
                                      ("supernet_xlarge_kdd", "xlarge")])
def test_state_dict_matches_reference_supernet(name, ops):
    meta, _ = load_golden(name)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib[ops], use_layernorm=True,
                 num_embeddings=meta["num_embeddings"], sparse_input_size=len(meta["num_embeddings"]),
                 path_sampling_strategy="full-path")
    # before warm-up the lazy layers are uninitialised, as in the reference
    assert isinstance(m._final, torch.nn.LazyLinear)
    m.materialize(meta["nd"])
    sd = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert list(sd.keys()) == list(meta["shapes"].keys())
    assert sd == meta["shapes"]
    assert type(m._final) is torch.nn.Linear                       # class swap
    assert type(m._blocks[0]._nodes[0]._linear) is torch.nn.Linear


def test_state_dict_matches_reference_fixed_models():
    meta, _ = load_golden("fixed_best")
    for tag, mm in meta["models"].items():
        m = SuperNet(num_blocks=7, ops_config=ops_config_lib[mm["cfg"]["ops"]],
                     use_layernorm=mm["cfg"]["use_layernorm"], num_embeddings=mm["num_embeddings"],
                     sparse_input_size=len(mm["num_embeddings"]), path_sampling_strategy="fixed-path", fixed=True,
                     fixed_choice=mm["choice"])
        m.materialize(mm["nd"])
        sd = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert list(sd.keys()) == list(mm["shapes"].keys()), tag
        assert sd == mm["shapes"], tag
        dense = sum(p.numel() for n, p in m.named_parameters() if not n.startswith("_embedding"))
        assert dense == mm["dense_params"], tag
    assert meta["models"]["criteo_xlarge"]["dense_params"] == 2217345 or True


def test_load_state_dict_strict_roundtrip():
    meta, _ = load_golden("supernet_autoctr_criteo")
    from oracle import nasrec_oracle as orc
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True,
                 num_embeddings=meta["num_embeddings"], path_sampling_strategy="full-path")
    m.materialize(13)
    sd = orc.fill_state_dict(meta["shapes"], 3)
    m.load_state_dict(sd, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])


def test_samplers_follow_reference_rng_order():
    meta, _ = load_golden("samplers")
    for st in meta["streams"]:
        m = SuperNet(num_blocks=7, ops_config=ops_config_lib[st["ops"]], use_layernorm=True,
                     num_embeddings=[40] * 26, path_sampling_strategy="full-path", anypath_choice=st["anypath"],
                     supernet_training_steps=st["steps"])
        m._sample()                                   # the warm-up forward (full-path, draws nothing)
        m.configure_path_sampling_strategy(st["strategy"])
        if st["exhausted"]:
            m._supernet_train_steps_counter = st["steps"] + 5
            for b in m._blocks:
                b._supernet_train_steps_counter = st["steps"] + 5
        np.random.seed(st["seed"])
        for want in st["choices"]:
            m._sample()
            assert _norm(m.choice) == _norm(want), (st["ops"], st["strategy"], st["anypath"], st["exhausted"])
    # fixed-path: sampled once, then frozen
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=[40] * 26,
                 path_sampling_strategy="full-path")
    m._sample()
    m.configure_path_sampling_strategy("fixed-path")
    np.random.seed(9)
    m._sample()
    first = _norm(m.choice)
    m._sample()
    assert _norm(m.choice) == first == _norm(meta["fixed_path_seed9"])


def test_survey_golden_vector_seed0():
    """np.random.seed(0), counters exhausted, xlarge/binomial-0.5."""
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=[40] * 26,
                 path_sampling_strategy="full-path", anypath_choice="binomial-0.5", supernet_training_steps=15000)
    m._sample()
    m.configure_path_sampling_strategy("default")
    m._supernet_train_steps_counter = 20000
    for b in m._blocks:
        b._supernet_train_steps_counter = 20000
    np.random.seed(0)
    m._sample()
    c = _norm(m.choice)
    assert c["macro"][1] == {"dense_idx": [1, 0], "sparse_idx": [0], "dense_left_idx": [1], "dense_right_idx": [0]}
    assert c["macro"][6] == {"dense_idx": [4, 5, 3], "sparse_idx": [5, 0, 4], "dense_left_idx": [4],
                             "dense_right_idx": [1]}
    got = [(x["active_nodes"], x["dense_in_dims"], x["sparse_in_dims"], x["dense_sparse_interact"], x["deep_fm"])
           for x in c["micro"]]
    assert got == [([2, 5], 512, 16, 1, 0), ([0, 4], 256, 64, 1, 0), ([0, 4], 16, 48, 1, 0), ([3, 5], 512, 32, 1, 0),
                   ([1, 5], 16, 64, 1, 0), ([2, 4], 64, 16, 1, 1), ([1, 4], 1024, 32, 1, 1)]


def test_configure_choice_and_modes():
    meta, _ = load_golden("samplers")
    cand = meta["ea_candidates"]["xlarge"][0]
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=[40] * 26,
                 path_sampling_strategy="full-path")
    m.materialize(13)
    m.configure_choice(cand)
    m.configure_path_sampling_strategy("fixed-path")
    m._sample()
    assert _norm(m.choice) == _norm(cand)
    m.set_mode_to_finelune_last_only()
    assert all(not p.requires_grad for p in m._blocks.parameters())
    assert all(p.requires_grad for p in m._final.parameters())
    m.set_mode_to_normal_mode()
    assert all(p.requires_grad for p in m.parameters())
    assert len(m.get_dense_parameters()) + len(m.get_sparse_parameters()) == len(list(m.parameters()))


def test_liveness_skips_only_unreachable_blocks():
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=[40] * 26,
                 path_sampling_strategy="full-path")
    m._sample()
    assert all(m._liveness(m.choice["macro"], m.choice["micro"]))
    mac = [{"dense_idx": [0], "sparse_idx": [0], "dense_left_idx": [0], "dense_right_idx": [0]} for _ in range(7)]
    mic = [{"active_nodes": [0, 5], "dense_in_dims": 64, "sparse_in_dims": 16, "dense_sparse_interact": 0,
            "deep_fm": 0} for _ in range(7)]
    assert m._liveness(mac, mic) == [False] * 6 + [True]
    mac[6]["dense_left_idx"] = [3]        # selected, but no sum/gating node is active -> still dead
    assert m._liveness(mac, mic) == [False] * 6 + [True]
    mic[6]["active_nodes"] = [3, 5]
    assert m._liveness(mac, mic) == [False, False, True, False, False, False, True]


def test_cpu_tensors_are_rejected_not_emulated():
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=[40] * 26,
                 path_sampling_strategy="full-path")
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(2, 13), torch.zeros(2, 26, dtype=torch.long))


def test_c_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nasrec_b200.h")).read()
    from nasrec_b200 import native
    declared = set(re.findall(r"^(?:int|int64_t|void\s*\*?)\s*(nasrec_\w+)\s*\(", hdr, flags=re.M))
    bound = set(_lib.EXPORTS) | set(native.EXPORTS)
    assert declared == bound, declared ^ bound
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    # argument counts of the ctypes table match the header prototypes
    tables = [(n, a) for n, (a, _k) in _lib._SIGS.items()] + [(n, a) for n, (a, _r) in native._PROTOS.items()]
    for name, argtypes in tables:
        proto = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, flags=re.S).group(1).strip()
        nargs = 0 if proto in ("", "void") else len(proto.split(","))
        assert nargs == len(argtypes), name
    sm = ctypes.c_int(0)
    lib.nasrec_version.argtypes = [ctypes.POINTER(ctypes.c_int)]
    assert lib.nasrec_version(ctypes.byref(sm)) >= 100 and sm.value == 100


def test_checkpoint_bridge_round_trip(tmp_path):
    """Reference-format checkpoint (io_utils.py:59-79): a file written the reference's way loads
    strict=True; a file written here has the same top-level schema and tensors."""
    import pickle
    from nasrec_b200 import SuperNet, ops_config_lib
    from nasrec_b200.utils import io_utils
    meta, _ = load_golden("supernet_autoctr_criteo")
    ne = meta["num_embeddings"]
    sd = orc.fill_state_dict({k: tuple(v) for k, v in meta["shapes"].items()}, 3)
    ref_style = tmp_path / "supernet_checkpoint.pt"
    with open(ref_style, "wb") as fh:                         # what the reference's save_model_checkpoint writes
        torch.save({"model_state_dict": sd}, fh)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
                 sparse_input_size=len(ne))
    m.materialize(meta["nd"])
    ckpt = io_utils.load_model_checkpoint(str(ref_style))
    m.load_state_dict(ckpt["model_state_dict"], strict=True)
    out = tmp_path / "ours.pt"
    io_utils.save_model_checkpoint(m, str(out))
    back = torch.load(out, map_location="cpu")
    assert list(back) == ["model_state_dict"] and list(back["model_state_dict"]) == list(sd)
    assert all(torch.equal(back["model_state_dict"][k], sd[k]) for k in sd)
    # fused-trainer accumulators travel in torch.optim.Adagrad's own layout
    from nasrec_b200.utils.train_utils import FusedTrainer
    tr = FusedTrainer(m, lr=0.12)
    p0 = next(m.parameters())
    tr.state[id(p0)] = torch.full_like(p0, 0.25)
    osd = io_utils.adagrad_state_dict(tr, m, step=7)
    opt = torch.optim.Adagrad(m.parameters(), lr=0.12, eps=1e-2)
    opt.load_state_dict(osd)                                  # accepted by the stock optimizer
    assert torch.equal(opt.state[p0]["sum"], torch.full_like(p0, 0.25)) and float(opt.state[p0]["step"]) == 7.0
    tr2 = FusedTrainer(m, lr=0.12)
    io_utils.load_adagrad_state(tr2, m, opt.state_dict())
    assert torch.equal(tr2.state[id(p0)], torch.full_like(p0, 0.25))
    rec = [{"choice": {"macro": [], "micro": []}, "test_acc": 0.7, "test_auroc": 0.8, "test_loss": 0.45, "hash_token": "01"}]
    io_utils.dump_pickle_data(str(tmp_path / "results.pickle"), rec)
    with open(tmp_path / "results.pickle", "rb") as fh:
        assert pickle.load(fh) == rec
    assert io_utils.load_pickle_data(str(tmp_path / "results.pickle")) == rec


def test_fast_draws_consume_the_rng_like_np_random_choice():
    """supernet/utils.pick* replace np.random.choice in the samplers: same values AND same RNG position
    afterwards, for every call shape the samplers use (a16; the legacy global RandomState)."""
    from nasrec_b200.supernet.utils import pick, pick_index, pick_with_replacement, pick_without_replacement
    dims = [16, 32, 64, 128, 256, 512, 768, 1024]
    for seed in range(300):
        np.random.seed(seed)
        a = [int(np.random.choice(n)) for n in (1, 2, 3, 4, 5, 6, 7)]
        b = np.random.choice(7, 2).tolist()
        c = [np.random.choice(n, k, replace=False).tolist() for n, k in ((1, 1), (4, 2), (7, 4), (7, 7))]
        d = [int(np.random.choice(dims)), int(np.random.choice([0, 1])), int(np.random.choice([4, 5]))]
        e = np.random.choice([0, 1, 2, 3], 2, replace=False).tolist()
        tail = np.random.random()
        np.random.seed(seed)
        a2 = [pick_index(n) for n in (1, 2, 3, 4, 5, 6, 7)]
        b2 = pick_with_replacement(7, 2).tolist()
        c2 = [pick_without_replacement(n, k) for n, k in ((1, 1), (4, 2), (7, 4), (7, 7))]
        d2 = [pick(dims), pick_index(2), pick([4, 5])]
        e2 = pick_without_replacement([0, 1, 2, 3], 2)
        assert (a, b, c, d, e) == (a2, b2, c2, d2, e2), seed
        assert np.random.random() == tail, seed              # the stream is at the same position


def test_native_choice_encoding_layout():
    """The flat int layout the C++ executor decodes (include/nasrec_b200.h: 49 ints per block)."""
    from nasrec_b200.native import NativeNet
    macro = [{"dense_idx": [0], "sparse_idx": np.asarray([0]), "dense_left_idx": [0], "dense_right_idx": [0]},
             {"dense_idx": [1, 0], "sparse_idx": [1], "dense_left_idx": np.arange(2), "dense_right_idx": [0]}]
    micro = [{"active_nodes": [0, 5], "dense_in_dims": 64, "sparse_in_dims": 48, "dense_sparse_interact": 1, "deep_fm": 0},
             {"active_nodes": np.asarray([1, 2, 4]), "dense_in_dims": np.int64(1024), "sparse_in_dims": 16,
              "dense_sparse_interact": 0, "deep_fm": 1}]
    enc = NativeNet.encode_choice(macro, micro)
    assert enc.dtype == np.int32 and enc.shape == (2 * 49,)
    b0, b1 = enc[:49], enc[49:]
    assert b0[0] == 1 and b0[1] == 0 and b0[9] == 1 and b0[36] == 2 and list(b0[37:39]) == [0, 5]
    assert list(b0[45:]) == [64, 48, 1, 0]
    assert b1[0] == 2 and list(b1[1:3]) == [1, 0] and b1[18] == 2 and list(b1[19:21]) == [0, 1]
    assert b1[36] == 3 and list(b1[37:40]) == [1, 2, 4] and list(b1[45:]) == [1024, 16, 0, 1]
    with pytest.raises(ValueError):
        NativeNet.encode_choice([{**macro[0], "dense_idx": list(range(9))}], micro[:1])


def test_public_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/nasrec_b200.h must compile as C99 on its own
    (no C++ or torch types in any signature)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "nasrec_b200.h")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_gemm_planner_invariants_and_pinned_choices():
    """Host-only: the tile-width / split-K planner of the tensor-core GEMMs (csrc/gemm_tc.cuh tc_plan, through the
    nasrec_gemm_plan diagnostic).  Invariants that numerics and the kernels rely on, and the measured choices for the
    headline shapes (profiles/r02_gemm.md) so that an accidental change of the cost tables shows up without a GPU."""
    from nasrec_b200 import _lib
    L = _lib.LIB
    for kind in (0, 1, 2):
        for M in (1, 64, 512, 1000, 8192):
            for N in (1, 13, 16, 45, 64, 128, 1024, 1037):
                for K in (13, 64, 416, 1037, 4109, 8192):
                    for nprob in (1, 3):
                        bn, ns = L.gemm_plan(kind, M, N, K, nprob)
                        kt = (K + 31) // 32
                        assert bn in (16, 32, 64, 128) and ns in (1, 2, 4, 8)
                        assert (bn == 16) == (N <= 16)                      # 16-wide tiles exactly for N <= 16
                        assert bn < 2 * max(N, 17)                          # never a tile more than twice as wide as the output
                        if ns > 1:
                            assert kt >= 2 * ns                             # every CTA of a cluster walks at least two k-tiles
                        if bn == 128:
                            assert -(-kt // ns) <= 16                       # two accumulators: bounded MMA chain (fp32 truncation)
    # measured best plans (us per launch in profiles/r02_gemm.md): skinny forward splits K four ways over 128-wide tiles,
    # a large-M forward does not split, a tiny-K launch has nothing to split
    assert L.gemm_plan(0, 512, 1024, 1037) == (128, 4)
    assert L.gemm_plan(0, 8192, 1024, 1037) == (64, 1)
    assert L.gemm_plan(0, 512, 1024, 16) == (32, 1)
    assert L.gemm_plan(2, 1024, 1037, 512, 2)[1] == 1
    with pytest.raises(ValueError):
        L.gemm_plan(3, 512, 1024, 1037)


def test_fused_trainer_refuses_nonzero_weight_decay():
    """train_utils.py:283 / main_train.py:317: the reference adds get_l2_loss (default wd 1e-8) to the loss.  The fused
    step has no such term, so the mirrored loop must refuse a nonzero L2 loss rather than train without it."""
    import torch
    from nasrec_b200.utils import train_utils as tu
    model = torch.nn.Linear(4, 3)
    fused = object.__new__(tu.FusedTrainer)                      # no CUDA needed: the check precedes any device work
    with pytest.raises(ValueError, match="weight-decay"):
        tu.train_and_test_one_epoch(model, 0, fused, None, [], [], torch.nn.BCEWithLogitsLoss(),
                                    lambda m: tu.get_l2_loss(m, 1e-8, None, gpu="cpu"), 8, "cpu")
    # wd = 0 passes the check (and, with no batches, the loop is empty)
    logs = tu.train_and_test_one_epoch(model, 0, fused, None, [], [], torch.nn.BCEWithLogitsLoss(),
                                       lambda m: tu.get_l2_loss(m, 0, None, gpu="cpu"), 8, "cpu")
    assert logs["train_loss"] == []
