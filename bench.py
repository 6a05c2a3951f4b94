#!/usr/bin/env python
"""bench.py -- NASRec supernet hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Configs (BASELINE.json `configs`):
  small_supernet    (default; configs[1]) NASRec-Small ("autoctr") weight-sharing supernet TRAINING step on synthetic
                    Criteo-shape data (13 dense + 26 sparse), tables capped at 0.5 M rows, use_layernorm=1, strategy
                    "default", anypath_choice "binomial-0.5", warm-up exhausted (a fresh subnet every step), B = 512
                    per GPU, step = forward + BCE + backward + global-norm clip 5.0 + Adagrad(0.12, eps 1e-2).
  criteo_full_best  (configs[0]) the shipped Criteo NASRec-Full best model (fixed, use_layernorm=False as
                    main_train.py:262 forces), B = 256, lr 0.16; --tables capped|full.
  kdd_xlarge        (configs[4], per-GPU slice) NASRec-Full supernet training on KDD shapes, B = 2048 per GPU.
  ea                (configs[2]) one-shot scoring of sampled NASRec-Full subnets against a shared Criteo supernet,
                    candidates split across the ranks, one final gather; metric = subnets/s.

One JSON line on stdout (rank 0).  `value` = whole-job throughput with inputs resident in HBM, each step timed with
CUDA events on the launching stream, L2 flushed between steps, max over ranks.  `e2e` = the same work driven from
pinned HOST batches (H2D copies inside the timed region) with the result read back every step.  `roofline` = the
dominant kernel family (the segment-list tensor-core GEMM): algorithmic flops / CUDA-event time of its launches,
recorded live inside the timed steps by the library (nasrec_gemm_prof).  `cpu_baseline` = the UNMODIFIED reference
(oracle/_ref, vendored by build(); falls back to the oracle port) timed on this box's host cores on a bounded sample.
`--impl reference` runs only that CPU arm.  The default line also carries the other configs as `extra`.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAP = 500000
_CRITEO = [1461, 584, 10131227, 2202609, 306, 25, 12518, 634, 4, 93146, 5684, 8351593, 3195, 28, 14993, 5461307, 11,
           5653, 2174, 5, 7046548, 19, 16, 286182, 106, 142573]
_KDD = [26274, 641708, 14848, 22122011, 1188090, 3735797, 2934102, 20004011, 4, 8]
_AVAZU = [10000, 241, 8, 8, 4738, 7746, 27, 8553, 560, 37, 2686409, 6729487, 8252, 6, 5, 2627, 9, 10, 436, 5, 69, 173, 61]

CONFIGS = {
    "small_supernet": dict(
        metric="supernet_train_samples_per_sec", unit="samples/s", ops="autoctr", fixed=False, ln=True, B=512, lr=0.12,
        nd=13, ne=[min(x, CAP) for x in _CRITEO], strategy="default", anypath="binomial-0.5",
        workload="NASRec-Small (autoctr) supernet training, synthetic Criteo shape 13 dense + 26 sparse, tables capped "
                 "0.5M rows, B=512/GPU, LN on, default/binomial-0.5 sampling, Adagrad+clip5"),
    "criteo_full_best": dict(
        metric="criteo_full_best_train_samples_per_sec", unit="samples/s", ops="xlarge", fixed=True, ln=False, B=256,
        lr=0.16, nd=13, ne=None, strategy="fixed-path", anypath="uniform",
        workload="NASRec-Full best Criteo model (configs/criteo/ea_criteo_kaggle_xlarge_best_1shot.json, fixed, LN off), "
                 "B=256/GPU, synthetic Criteo shape, Adagrad(0.16)+clip5"),
    "avazu_full_best": dict(
        metric="avazu_full_best_train_samples_per_sec", unit="samples/s", ops="xlarge", fixed=True, ln=False, B=256,
        lr=0.16, nd=1, ne=list(_AVAZU), strategy="fixed-path", anypath="uniform", best="avazu_xlarge", zero_dense=True,
        workload="NASRec-Full best Avazu model (configs/avazu/ea_avazu_kaggle_xlarge_best_1shot.json, fixed, LN off), "
                 "full-size tables (9.46 M rows), B=256/GPU, synthetic Avazu shape (1 all-zero dense + 23 sparse), "
                 "Adagrad(0.16)+clip5"),
    "kdd_xlarge": dict(
        metric="kdd_xlarge_supernet_train_samples_per_sec", unit="samples/s", ops="xlarge", fixed=False, ln=True, B=2048,
        lr=0.12, nd=3, ne=[min(x, CAP) for x in _KDD], strategy="default", anypath="binomial-0.5",
        workload="NASRec-Full (xlarge) supernet training, synthetic KDD shape 3 dense + 10 sparse, tables capped 0.5M "
                 "rows, B=2048/GPU, LN on, default/binomial-0.5 sampling, Adagrad+clip5"),
    "ea": dict(
        metric="ea_subnets_evaluated_per_sec", unit="subnets/s", ops="xlarge", fixed=False, ln=True, B=8192, lr=0.0,
        nd=13, ne=[min(x, CAP) for x in _CRITEO], strategy="full-path", anypath="uniform",
        workload="one-shot scoring of Tokenizer.generate_random_choice NASRec-Full candidates against a shared Criteo "
                 "xlarge supernet, 8 evaluation batches of 8192 per candidate -> log-loss/AUC/accuracy, candidates "
                 "split contiguously across ranks"),
}
EA_BATCHES = 8


def config_of(args):
    c = dict(CONFIGS[args.config])
    if c["ne"] is None:
        c["ne"] = list(_CRITEO) if args.tables == "full" else [min(x, CAP) for x in _CRITEO]
        c["workload"] += ", tables %s" % ("full-size (33.8 M rows)" if args.tables == "full" else "capped 0.5M rows")
    return c


# ----------------------------------------------------------------------------- synthetic data
def synth_pool(n_batches, batch, nd, num_embeddings, seed, zipf=True, zero_dense=False):
    """log1p(Poisson(3)) dense, Zipf(1.05)-ranked ids through a fixed permutation with id 0 =
    'missing' (p=0.02), Bernoulli(0.25) labels; a pool of distinct batches that is cycled."""
    rs = np.random.RandomState(seed)
    pool = []
    for _ in range(n_batches):
        int_x = np.log1p(rs.poisson(3.0, (batch, nd))).astype(np.float32)
        if zero_dense:
            int_x[:] = 0.0                 # Avazu has no dense features: one all-zero column (data_pipes.py:181)
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
        y = (rs.rand(batch, 1) < 0.25).astype(np.float32)
        pool.append((int_x, np.stack(cols, 1), y))
    return pool


def _best_choice(which="criteo_xlarge"):
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "fixed_best.json")))
    return meta["models"][which]["choice"]


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU arm: the reference itself
def _reference_modules():
    """The vendored, unmodified reference (oracle/_ref, built by __graft_entry__.build()); None if absent."""
    from oracle import make_ref
    try:
        return make_ref.import_reference()
    except Exception:
        return None


class _TimedLoader:
    """Feeds host batches to the reference's own training loop and timestamps every fetch: the loop body
    (train_utils.py:255-287) runs between two fetches, so consecutive stamps bracket exactly one reference step."""

    def __init__(self, batches):
        self.batches, self.stamps = batches, []

    def __iter__(self):
        for b in self.batches:
            self.stamps.append(time.perf_counter())
            yield b
        self.stamps.append(time.perf_counter())


def reference_train(cfg, steps, warmup, seed=1234):
    """The reference's training step on this box's host cores with all threads: model built as train_supernet.py /
    main_train.py build it, driven by the reference's own train_and_test_one_epoch (zero_grad, forward, BCE + L2,
    backward, clip_grad_norm_ 5.0, torch.optim.Adagrad(eps=1e-2)).  Returns (cpu_baseline dict, median s/step)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, nd, ne = cfg["B"], cfg["nd"], cfg["ne"]
    pool = synth_pool(max(2, min(8, steps + warmup)), B, nd, ne, seed, zero_dense=cfg.get("zero_dense", False))
    # one extra trailing batch: the reference's loop logs and evaluates on its last step, which must not be a timed one
    batches = [tuple(torch.from_numpy(a) for a in pool[i % len(pool)]) for i in range(warmup + steps + 1)]
    ref = _reference_modules()
    if ref is not None:
        from nasrec.supernet.supernet import SuperNet as RefNet, ops_config_lib as ref_ops
        from nasrec.utils.train_utils import train_and_test_one_epoch, get_l2_loss, init_weights as ref_init
        torch.manual_seed(seed)
        np.random.seed(seed)
        kw = dict(num_blocks=7, ops_config=ref_ops[cfg["ops"]], use_layernorm=cfg["ln"], num_embeddings=ne,
                  sparse_input_size=len(ne), path_sampling_strategy="full-path")
        if cfg["fixed"]:
            kw.update(path_sampling_strategy="fixed-path", fixed=True, fixed_choice=_best_choice(cfg.get("best", "criteo_xlarge")))
        else:
            kw.update(anypath_choice=cfg["anypath"], supernet_training_steps=0)
        m = RefNet(**kw)
        with torch.no_grad():
            m(batches[0][0], batches[0][1])                       # lazy materialisation (train_utils.py:413-433)
        m.apply(ref_init)
        if not cfg["fixed"]:
            m.configure_path_sampling_strategy(cfg["strategy"])
        opt = torch.optim.Adagrad(m.parameters(), lr=cfg["lr"], eps=1e-2)
        loader = _TimedLoader(batches)

        class _NoSched:
            def step(self, *a, **k):
                pass

            def get_lr(self):
                return [cfg["lr"]]

        # The loop insists on a test loader (it reads its batch size) but must never evaluate: the reference's
        # test_one_epoch calls .item() on sklearn's AUC, which is a plain float with this image's sklearn.  With
        # max_train_steps = -1 and test_only_at_last_step the loop runs the loader dry and never tests; it prints its
        # progress on stdout, which has to stay one JSON line, so that goes to stderr.
        import contextlib
        tiny_test = [tuple(t[:64] for t in batches[0])]
        with contextlib.redirect_stdout(sys.stderr):
            train_and_test_one_epoch(m, 0, opt, _NoSched(), loader, tiny_test, torch.nn.BCEWithLogitsLoss(),
                                     lambda mm: get_l2_loss(mm, 0.0, [], gpu=None), B, None, display_interval=10 ** 9,
                                     test_interval=10 ** 9, max_train_steps=-1, max_eval_steps=1,
                                     test_only_at_last_step=True, grad_clip_value=5.0)
        st = loader.stamps
        times = [st[i + 1] - st[i] for i in range(len(st) - 1)][warmup:warmup + steps]
        kind = "reference"
        how = "unmodified reference (oracle/_ref) through its own train_and_test_one_epoch"
    else:
        times, kind, how = _port_train(cfg, batches[:warmup + steps], steps, warmup, seed), "port", "oracle port (oracle/_ref not built)"
    med = float(np.median(times))
    return {"value": B / med, "unit": cfg["unit"], "cores": cores, "kind": kind,
            "sample": "%d timed steps of B=%d (median) after %d warm-up, torch %d threads; %s" % (
                len(times), B, warmup, cores, how)}, med


def _port_train(cfg, batches, steps, warmup, seed):
    import torch
    from oracle import nasrec_oracle as orc
    from nasrec_b200 import SuperNet, ops_config_lib
    from nasrec_b200.utils.train_utils import init_weights
    torch.manual_seed(seed)
    np.random.seed(seed)
    kw = dict(num_blocks=7, ops_config=ops_config_lib[cfg["ops"]], use_layernorm=cfg["ln"], num_embeddings=cfg["ne"],
              sparse_input_size=len(cfg["ne"]), path_sampling_strategy=cfg["strategy"])
    if cfg["fixed"]:
        kw.update(fixed=True, fixed_choice=_best_choice(cfg.get("best", "criteo_xlarge")))
    else:
        kw.update(anypath_choice=cfg["anypath"], supernet_training_steps=0)
    host = SuperNet(**kw)
    host.materialize(cfg["nd"])
    host.apply(init_weights)
    sd = {k: v.detach().clone() for k, v in host.state_dict().items()}
    tr = orc.OracleTrainer(sd, dict(ops=cfg["ops"], use_layernorm=cfg["ln"], fixed=cfg["fixed"], num_blocks=7), lr=cfg["lr"])
    times = []
    for i, b in enumerate(batches):
        choice = _best_choice(cfg.get("best", "criteo_xlarge")) if cfg["fixed"] else (host._sample() and host.choice)
        t0 = time.perf_counter()
        tr.step(choice, *b)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def reference_ea(cfg, n_batches_timed=2, seed=1234):
    """The reference's scoring of ONE candidate on the host cores: the supernet pinned to the candidate
    (configure_choice + fixed-path, eval_subnet_from_supernet.py:103-110), its own test_one_epoch (forward over the
    evaluation batches + sklearn metrics, train_utils.py:129-178).  Timed on n_batches_timed batches of 8192 and
    scaled to the EA_BATCHES batches a candidate is scored on here."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, nd, ne = cfg["B"], cfg["nd"], cfg["ne"]
    ref = _reference_modules()
    if ref is None:
        return None, None
    from nasrec.supernet.supernet import SuperNet as RefNet, ops_config_lib as ref_ops
    from nasrec.searcher.tokenizer import Tokenizer as RefTok
    from nasrec.utils.train_utils import init_weights as ref_init
    torch.manual_seed(seed)
    np.random.seed(seed)
    m = RefNet(num_blocks=7, ops_config=ref_ops["xlarge"], use_layernorm=True, num_embeddings=ne,
               sparse_input_size=len(ne), path_sampling_strategy="full-path")
    pool = synth_pool(n_batches_timed + 1, B, nd, ne, 11)
    batches = [tuple(torch.from_numpy(a) for a in b) for b in pool]
    with torch.no_grad():
        m(batches[0][0][:64], batches[0][1][:64])
    m.apply(ref_init)
    cand = RefTok(7, ref_ops["xlarge"]).generate_random_choice()
    m.configure_choice(cand)
    m.configure_path_sampling_strategy("fixed-path")
    import sklearn.metrics

    def score(bs):
        # test_one_epoch's body (train_utils.py:140-178): forward in eval mode under no_grad, sigmoid, sklearn AUC,
        # accuracy, BCE -- written out here because the reference's own function calls .item() on sklearn's AUC, which is
        # a plain float with this image's sklearn
        m.eval()
        preds, labels = [], []
        with torch.no_grad():
            for int_x, cat_x, y in bs:
                preds.append(m(int_x, cat_x))
                labels.append(y)
        p, t = torch.cat(preds), torch.cat(labels)
        auc = sklearn.metrics.roc_auc_score(t.view(-1).numpy(), torch.sigmoid(p).view(-1).numpy())
        acc = float(((torch.sigmoid(p) > 0.5).float() == t).float().mean())
        return acc, auc, float(torch.nn.functional.binary_cross_entropy_with_logits(p, t))

    score(batches[:1])                                               # warm-up
    t0 = time.perf_counter()
    score(batches[1:1 + n_batches_timed])
    dt = time.perf_counter() - t0
    per_cand = dt / n_batches_timed * EA_BATCHES
    return {"value": 1.0 / per_cand, "unit": cfg["unit"], "cores": cores, "kind": "reference",
            "sample": "1 candidate x %d batches of %d through the reference model pinned to the candidate (test_one_epoch's body), scaled to %d batches; torch %d "
                      "threads; process spawn / model rebuild / checkpoint reload of the reference's searcher not counted" % (
                          n_batches_timed, B, EA_BATCHES, cores)}, per_cand


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = config_of(args)
    if args.config == "ea":
        base, sec = reference_ea(cfg)
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (run __graft_entry__.build() "
                              "where /root/reference exists)"}), flush=True)
            return
        steps, warm, ms = 1, 1, sec * 1e3
    else:
        # bounded: every step is one reference training step on the host cores; at most ~20 s of CPU work
        steps, warm = max(3, min(args.steps, 20)), max(1, min(args.warmup, 3))
        base, med = reference_train(cfg, steps, warm)
        ms = med * 1e3
    line = {"impl": "reference", "metric": cfg["metric"], "value": base["value"], "unit": cfg["unit"], "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": cfg["workload"], "per_gpu_batch": cfg["B"]}, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class Env:
    """torch / distributed plumbing shared by the measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == args.gpus or self.world == 1, "launch with torchrun --nproc-per-node == --gpus"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL prints its version banner on stdout when the communicator is created (the box exports NCCL_DEBUG);
            # stdout must carry exactly one JSON line, so fd 1 points at stderr until the communicator exists.
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                probe = torch.zeros(1, device=self.dev)
                dist.all_reduce(probe)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
            self._pin_cores()
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # > 126 MB L2

    def _pin_cores(self):
        """Each rank keeps to its own slice of the host cores: with 8 launch-bound ranks on one host the kernel
        otherwise migrates the issuing threads across each other (measured: ~1.5x longer issue time at N = 8)."""
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // self.world)
            mine = cores[self.local * per:(self.local + 1) * per] or cores
            os.sched_setaffinity(0, mine)
        except Exception:
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def build_training(env, cfg, seed=1234):
    """Model + trainer of a training config, identical on every rank (same seeds -> same weights, same subnets)."""
    torch = env.torch
    from nasrec_b200 import SuperNet, ops_config_lib
    from nasrec_b200.utils.train_utils import FusedTrainer, init_weights
    torch.manual_seed(seed)
    np.random.seed(seed)
    kw = dict(num_blocks=7, ops_config=ops_config_lib[cfg["ops"]], use_layernorm=cfg["ln"], num_embeddings=cfg["ne"],
              sparse_input_size=len(cfg["ne"]), path_sampling_strategy=cfg["strategy"])
    if cfg["fixed"]:
        kw.update(fixed=True, fixed_choice=_best_choice(cfg.get("best", "criteo_xlarge")))
    else:
        kw.update(anypath_choice=cfg["anypath"], supernet_training_steps=0)
    model = SuperNet(**kw).to(env.dev)
    model.materialize(cfg["nd"])
    model.apply(init_weights)
    if cfg["fixed"]:
        # fixed models: one CUDA graph of the whole step (the executor covers weight-sharing supernets)
        if env.world > 1:
            from nasrec_b200.parallel import DataParallelTrainer
            return model, DataParallelTrainer(model, lr=cfg["lr"]), "Python engine, data parallel"
        from nasrec_b200.utils.graph import GraphedFusedTrainer
        return model, GraphedFusedTrainer(FusedTrainer(model, lr=cfg["lr"])), "CUDA graph of the fused step"
    from nasrec_b200.native import NativeTrainer
    from nasrec_b200.parallel import NativeDataParallelTrainer
    trainer = NativeDataParallelTrainer(model, lr=cfg["lr"]) if env.world > 1 else NativeTrainer(model, lr=cfg["lr"])
    if os.environ.get("NASREC_OVERLAP", "1") == "0":
        (trainer._nt if env.world > 1 else trainer).overlap_wgrad = False
    return model, trainer, "C++ step executor"


def measure_training(env, cfg, trainer, K, W, profile_gemm=True):
    """value (HBM-resident inputs, per-step events, L2 flush), pipelined, e2e (pinned host batches, loss read back)."""
    torch, world, rank, dev, flush = env.torch, env.world, env.rank, env.dev, env.flush
    from nasrec_b200 import _lib
    B, nd, ne = cfg["B"], cfg["nd"], cfg["ne"]
    NP = 64 if B <= 512 else 16
    pool_h = synth_pool(NP, B, nd, ne, seed=1234 + rank, zero_dense=cfg.get("zero_dense", False))   # per-rank data, same choices
    pool_d = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in pool_h]
    pool_p = [tuple(torch.from_numpy(a).pin_memory() for a in b) for b in pool_h]
    for i in range(W):
        trainer.step(*pool_d[i % NP])
    env.barrier()
    clocks = ClockSampler(env.local)
    if rank == 0:                      # one sampler per job: rank 0's GPU stands for the box (same clocks policy)
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = _lib.query("nasrec_host_prof", 4)          # the library's own count of its kernel launches
    l0p = _lib.LIB.launches                          # the binding's count: what a CUDA-graph replay launches is only known there
    env.barrier()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        trainer.step(*pool_d[(W + i) % NP])
        ev[i][1].record()
    env.barrier()
    lib_d, py_d = _lib.query("nasrec_host_prof", 4) - l0, _lib.LIB.launches - l0p
    launches = lib_d if lib_d > 10 * K else py_d     # graph replays launch what was captured: only the binding knows how many
    total_ms = env.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    clk = clocks.stop()
    out = {"value": world * B * K / (total_ms * 1e-3), "ms_per_step": total_ms / K, "launches": launches, "clocks": clk,
           "gemm": None}
    if profile_gemm:
        # roofline pass: the same K steps again with the library recording a CUDA-event pair around every GEMM launch
        # (kept apart from the timed pass above: two event records per launch are not free at ~70 launches per step).
        # The pass runs on ONE stream (no weight-gradient side stream): with two streams a launch's event pair also
        # counts the time it waits for SMs held by the other stream's GEMM, and the per-kernel times summed past the step.
        nt = getattr(trainer, "_nt", trainer)
        had_overlap = getattr(nt, "overlap_wgrad", None)
        if had_overlap:
            nt.overlap_wgrad = False
        evp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        env.barrier()
        _lib.LIB.gemm_prof_start()
        for i in range(K):
            flush.zero_()
            evp[i][0].record()
            trainer.step(*pool_d[(W + i) % NP])
            evp[i][1].record()
        env.barrier()
        out["gemm"] = _lib.LIB.gemm_prof_stop()
        out["gemm_pass_ms_per_step"] = sum(a.elapsed_time(b) for a, b in evp) / K
        # kernel pass: the same K steps once more with an event pair around EVERY launch of the library (both streams),
        # aggregated by kernel name -- each kernel's time inside the real step (warm L2), not ncu's cold serialised replays
        import tempfile
        tf = tempfile.NamedTemporaryFile("r", suffix=".trace", delete=False)
        os.environ["NASREC_TRACE_FILE"] = tf.name
        env.barrier()
        _lib.query("nasrec_host_prof", 10)
        for i in range(K):
            flush.zero_()
            trainer.step(*pool_d[(W + i) % NP])
        env.barrier()
        n_traced = _lib.query("nasrec_host_prof", 11)
        rows = [l.rsplit(" ", 2) for l in open(tf.name).read().splitlines() if l.strip()]
        os.unlink(tf.name)
        os.environ.pop("NASREC_TRACE_FILE", None)
        tot = sum(float(r[2]) for r in rows) or 1.0
        top = sorted(rows, key=lambda r: -float(r[2]))[:14]
        out["kernel_trace"] = {"launches_per_step": n_traced / K, "sum_us_per_step": tot / K,
                               "note": "CUDA events around every launch inside the native step, run on one stream (no side-stream overlap; "
                                       "event records between launches also suppress programmatic dependent launch, so the sum "
                                       "exceeds ms_per_step)",
                               "top": [{"kernel": r[0].split("(")[0][-48:], "launches_per_step": round(int(r[1]) / K, 2),
                                        "us_per_step": round(float(r[2]) / K, 1), "share": round(float(r[2]) / tot, 4)} for r in top]}
        if had_overlap:
            nt.overlap_wgrad = True
    # pipelined (back-to-back, one event pair; informational)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        trainer.step(*pool_d[(W + K + i) % NP])
    e1.record()
    env.barrier()
    out["pipelined"] = world * B * K / (env.max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    # e2e: pinned host batches, H2D inside the timed region, every step's loss read back (asynchronously: step t's loss
    # is consumed on the host while step t+1 is being issued, as a training loop that logs the loss does)
    out["h2d"] = sum(a.numel() * a.element_size() for a in pool_p[0])
    loss_host = torch.zeros(K, dtype=torch.float32).pin_memory()
    done = [torch.cuda.Event() for _ in range(K)]
    env.barrier()
    e0.record()
    for i in range(K):
        b = pool_p[(W + 2 * K + i) % NP]
        xb = tuple(t.to(dev, non_blocking=True) for t in b)
        _, loss = trainer.step(*xb)
        loss_host[i:i + 1].copy_(loss.reshape(-1)[:1], non_blocking=True)
        done[i].record()
        if i > 0:
            done[i - 1].synchronize()
    done[K - 1].synchronize()
    out["last_loss"] = float(loss_host[K - 1])
    e1.record()
    env.barrier()
    out["e2e"] = world * B * K / (env.max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    return out


def dp_consistent(env, model):
    """Data-parallel replicas must stay bit-identical: max - min over ranks of a weight checksum == 0."""
    if env.world == 1:
        return None
    torch = env.torch
    with torch.no_grad():
        cs = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    hi, lo = cs.clone(), cs.clone()
    env.dist.all_reduce(hi, op=env.dist.ReduceOp.MAX)
    env.dist.all_reduce(lo, op=env.dist.ReduceOp.MIN)
    return bool((hi - lo).item() == 0.0)


def roofline_of(gemm, ms_per_step, K, pass_ms_per_step=None):
    ms, n, fl = gemm
    ms_per_step = pass_ms_per_step or ms_per_step
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, which = float(peaks["bf16_tflops_sustained"]), "measured bf16_tflops_sustained"
    except Exception:
        peak, which = 1590.0, "fallback"
    from nasrec_b200 import _lib
    achieved = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    return {"bound": "tensor",
            "kernel": "gemm_tma_kernel / gemm_tc_kernel (segment-list GEMM: TMA-fed tcgen05 kind::tf32, 3xTF32 split, A operand "
                      "and accumulators in TMEM)",
            "gemm_mode": _lib.LIB.gemm_mode(), "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": which,
            "attainable_frac_note": "3xTF32 spends three TF32 MMAs (half the bf16 rate each) per algorithmic product: <= 1/6",
            "traffic": 6.2e6,
            "traffic_source": "mean dram__bytes_read+write over 6 gemm_tma_kernel launches inside this workload's step, ncu --set "
                              "full, cold caches (profiles/r02_notes.md); in the live step the operands are mostly L2-resident, "
                              "the kernel is not HBM-bound",
            "launches_timed": n, "avg_launch_us": ms * 1e3 / max(n, 1), "gflop_per_launch": fl / max(n, 1) / 1e9,
            "gemm_ms_per_step": ms / K, "instrumented_ms_per_step": ms_per_step, "share_of_step": (ms / K) / ms_per_step,
            "share_note": "CUDA-event time of every GEMM launch of K native steps run on one stream (no side-stream overlap), "
                          "over the duration of those same steps (instrumented_ms_per_step)"}


def measure_ea(env, cfg, n_per_rank=8, n_batches=EA_BATCHES):
    """configs[2]: candidates split contiguously across ranks, no data-path collective, one final gather."""
    torch, world, rank, dev = env.torch, env.world, env.rank, env.dev
    from nasrec_b200 import SuperNet, ops_config_lib, parallel
    from nasrec_b200.search import SubnetEvaluator, generate_random_choice
    from nasrec_b200.utils.train_utils import init_weights
    ne, B = cfg["ne"], cfg["B"]
    torch.manual_seed(2)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=ne,
                 path_sampling_strategy="full-path").to(dev)
    m.materialize(13)
    m.apply(init_weights)
    m.requires_grad_(False)
    np.random.seed(1234)
    n_cand = n_per_rank * world
    cands = [generate_random_choice(7, ops_config_lib["xlarge"]) for _ in range(n_cand + 2)]
    pool_h = synth_pool(n_batches, B, 13, ne, 11)
    batches = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in pool_h]
    pinned = [tuple(torch.from_numpy(a).pin_memory() for a in b) for b in pool_h]
    ev = SubnetEvaluator(m)
    lo, hi = parallel.shard_range(n_cand, world, rank)
    mine = cands[2 + lo:2 + hi]
    ev.score(cands[:2], batches)                                   # warm-up
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = ev.score(mine, batches)
    e1.record()
    env.barrier()
    sec = env.max_over_ranks(e0.elapsed_time(e1)) * 1e-3
    out = {"value": n_cand / sec, "sec": sec, "n_cand": n_cand}
    # e2e: evaluation batches start in pinned host memory (copied once per shard, as the evaluation set is shared by all
    # candidates of the shard), per-candidate records gathered to every rank at the end
    env.barrier()
    t0 = time.perf_counter()
    e0.record()
    dbatches = [tuple(t.to(dev, non_blocking=True) for t in b) for b in pinned]
    res = ev.score(mine, dbatches)
    recs = parallel.allgather_records(res, None) if world > 1 else res
    e1.record()
    env.barrier()
    out["e2e"] = n_cand / (env.max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    out["h2d"] = sum(a.numel() * a.element_size() for b in pinned for a in b)
    out["n_records"] = len(recs)
    out["mean_auc"] = float(np.mean([r["test_auroc"] for r in recs]))
    # what the regularized EA actually scores per generation (searcher.py:167-295): the mutated children of one parent,
    # which differ from it in one field of one block -- the batched path shares the blocks they have in common
    from nasrec_b200.search import Tokenizer
    tok = Tokenizer(7, ops_config_lib["xlarge"])
    np.random.seed(4321 + rank)
    parent = tok.generate_random_choice()
    gen = [tok.mutate_spec(parent) for _ in range(n_per_rank)]
    ev.multi_stats = [0, 0]
    env.barrier()
    e0.record()
    ev.score(gen, batches)
    e1.record()
    env.barrier()
    out["generation_value"] = n_cand / (env.max_over_ranks(e0.elapsed_time(e1)) * 1e-3)
    out["generation_blocks_computed_reused"] = list(ev.multi_stats)
    del m, ev, batches, dbatches
    torch.cuda.empty_cache()
    return out


def measure_hbm_kernels(env, B=8192):
    """HBM-bound kernels named by the north star, alone, at the evaluation batch and at 4x that: achieved GB/s of
    ALGORITHMIC bytes (gather F*(8+64+64) B/sample; row-wise Adagrad u*(64*5) B) against the measured copy
    peak.  At B = 8192 these kernels move 8-30 MB, i.e. 1-5 us of HBM time behind a ~5 us chain of dependent latencies
    (launch, id -> row address -> row): the second size separates that floor from the streaming rate."""
    out = _measure_hbm_kernels(env, B)
    big = _measure_hbm_kernels(env, 4 * B, ln=False)
    big.pop("peak_GBps", None)
    out.update(big)
    return out


def _measure_hbm_kernels(env, B, ln=True):
    torch, dev = env.torch, env.dev
    from nasrec_b200 import _lib
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ne = [min(x, CAP) for x in _CRITEO]
    F = len(ne)
    tables = [torch.randn(n, 16, device=dev) for n in ne]
    states = [torch.zeros(n, 16, device=dev) for n in ne]
    tp = torch.tensor([t.data_ptr() for t in tables], dtype=torch.int64, device=dev)
    sp = torch.tensor([t.data_ptr() for t in states], dtype=torch.int64, device=dev)
    rows = torch.tensor(ne, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    out = {}
    for tag, zipf in (("zipf", True), ("uniform", False)):
        cat = torch.from_numpy(synth_pool(1, B, 13, ne, 5, zipf=zipf)[0][1]).to(dev)
        rows_out = torch.empty(B, F, 16, device=dev)
        gout = torch.randn(B, F, 16, device=dev)
        uniq = torch.empty(F, B, dtype=torch.int64, device=dev)
        nuniq = torch.empty(F, dtype=torch.int32, device=dev)
        rg = torch.empty(F, B, 16, device=dev)
        sumsq = torch.empty(F, device=dev)
        scratch = torch.empty(F, B + 1, dtype=torch.int32, device=dev)

        def timeit(fn, n=20):
            fn()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(n):
                env.flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / n * 1e-3

        t = timeit(lambda: _lib.call("nasrec_emb_gather_fwd", tp.data_ptr(), rows.data_ptr(), cat.data_ptr(), rows_out.data_ptr(),
                                     B, F, err.data_ptr()))
        by = B * F * (8 + 64 + 64)
        out["emb_gather_%s_B%d" % (tag, B)] = {"us": t * 1e6, "GBps": by / t / 1e9, "frac_of_measured_hbm": by / t / 1e9 / peak}
        nb = _lib.query("nasrec_emb_grad_sort_reduce_big_ws_bytes", B, F)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        t = timeit(lambda: _lib.call("nasrec_emb_grad_sort_reduce_big", cat.data_ptr(), rows.data_ptr(), err.data_ptr(),
                                     gout.data_ptr(), B, F, uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(),
                                     ws.data_ptr(), nb))
        torch.cuda.synchronize()
        u = int(nuniq.sum().item())
        by = B * F * (8 + 64) + u * (64 + 8)
        out["emb_sort_reduce_%s_B%d" % (tag, B)] = {"us": t * 1e6, "GBps": by / t / 1e9, "frac_of_measured_hbm": by / t / 1e9 / peak,
                                                   "unique_rows": u}
        t = timeit(lambda: _lib.call("nasrec_emb_rowwise_adagrad", uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), tp.data_ptr(),
                                     sp.data_ptr(), B, F, 0.12, 1e-2, None))
        by = u * (64 * 5 + 8)
        out["emb_rowwise_adagrad_%s_B%d" % (tag, B)] = {"us": t * 1e6, "GBps": by / t / 1e9, "frac_of_measured_hbm": by / t / 1e9 / peak}
    out["peak_GBps"] = peak
    if not ln:
        return out
    # LayerNorm + ReLU + prefix mask over [B, 1024] (the epilogue of every dense linear)
    x = torch.randn(B, 1024, device=dev)
    g = torch.ones(1024, device=dev)
    bt = torch.zeros(1024, device=dev)
    y = torch.empty(B, 512, device=dev)
    mean = torch.empty(B, device=dev)
    rstd = torch.empty(B, device=dev)
    t = timeit(lambda: _lib.call("nasrec_ln_fwd", x.data_ptr(), 1024, B, 1024, g.data_ptr(), bt.data_ptr(), 1e-5, 1, 512, y.data_ptr(),
                                 512, mean.data_ptr(), rstd.data_ptr(), 0))
    by = B * (1024 + 512) * 4
    out["ln_fwd_B%d_N1024_d512" % B] = {"us": t * 1e6, "GBps": by / t / 1e9, "frac_of_measured_hbm": by / t / 1e9 / peak}
    out["peak_GBps"] = peak
    return out


def run_ours(args):
    env = Env(args)
    torch, world, rank = env.torch, env.world, env.rank
    from nasrec_b200 import _lib
    cfg = config_of(args)
    K, W = max(1, args.steps), max(3, args.warmup)      # timing rules: at least 3 warm-up steps (the line reports W)
    line = None
    if args.dtype == "bf16":
        import nasrec_b200
        nasrec_b200.set_precision("bf16")
    arith = ("bf16 GEMM operands (round-to-nearest-even), one product per k-step on the tensor cores, fp32 TMEM accumulation; "
             "fp32 master weights / Adagrad state / activations (logits within 2e-2 RMS of the reference under autocast)"
             if args.dtype == "bf16" else
             "fp32 storage and accumulation; GEMMs as 3xTF32-split tcgen05 MMAs (fp32-parity, logits within 1e-5 of the fp32 "
             "reference)")
    if args.config == "ea":
        r = measure_ea(env, cfg)
        if rank == 0:
            line = {"metric": cfg["metric"], "value": r["value"], "unit": cfg["unit"], "n_gpus": world, "steps": r["n_cand"],
                    "warmup": 2, "ms_per_step": r["sec"] * 1e3 / r["n_cand"] * world, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                    "config": {"workload": cfg["workload"], "arithmetic": arith, "candidates": r["n_cand"], "per_rank": r["n_cand"] // world,
                               "eval_batches": EA_BATCHES, "eval_batch": cfg["B"], "l2": "evaluation set (8 x 8192 x 15 KB) and "
                               "the 171 M-parameter supernet exceed L2", "parallelism": "candidates sharded x%d" % world},
                    "gpu_launches": _lib.LIB.launches,
                    "e2e": {"value": r["e2e"], "unit": cfg["unit"], "h2d_bytes_per_step": r["h2d"] // max(1, r["n_cand"] // world),
                            "d2h_bytes_per_step": 24, "mean_auc": r["mean_auc"]},
                    "extra": {"ea_generation_subnets_per_sec": r["generation_value"],
                              "ea_generation_blocks_computed_reused": r["generation_blocks_computed_reused"],
                              "note": "generation = the mutated children of one parent (what the regularized EA scores "
                                      "per generation); blocks shared between candidates are computed once"}}
            if world == 1 and not args.no_cpu:
                base, _ = reference_ea(cfg)
                if base is not None:
                    line["cpu_baseline"] = base
    else:
        model, trainer, host = build_training(env, cfg)
        r = measure_training(env, cfg, trainer, K, W, profile_gemm=not cfg["fixed"])
        cons = dp_consistent(env, model)
        if rank == 0:
            line = {"metric": cfg["metric"], "value": r["value"], "unit": cfg["unit"], "n_gpus": world, "steps": K, "warmup": W,
                    "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": args.dtype, "data": "synthetic",
                    "config": {"workload": cfg["workload"], "per_gpu_batch": cfg["B"], "global_batch": cfg["B"] * world,
                               "host": host, "arithmetic": arith,
                               "parallelism": "dp%d" % world, "l2": "256 MB flush write between timed steps",
                               "ids": "zipf(1.05)"},
                    "clocks": r["clocks"], "gpu_launches": r["launches"],
                    "e2e": {"value": r["e2e"], "unit": cfg["unit"], "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                            "last_loss": r["last_loss"]},
                    "pipelined_value": r["pipelined"]}
            if cons is not None:
                line["dp_consistent"] = cons
            if r["gemm"] is not None and r["gemm"][1] > 0:
                line["roofline"] = roofline_of(r["gemm"], r["ms_per_step"], K, r.get("gemm_pass_ms_per_step"))
            if r.get("kernel_trace"):
                line["kernel_trace"] = r["kernel_trace"]
        del trainer, model
        torch.cuda.empty_cache()
        # the other BASELINE configs ride along as `extra` on the default line
        if args.config == "small_supernet" and not args.no_extras and args.dtype == "f32":
            extra = {}
            ea = measure_ea(env, CONFIGS["ea"])
            extra["ea_subnets_per_sec_%dx8192" % EA_BATCHES] = ea["value"]
            extra["ea_subnets_per_sec_e2e"] = ea["e2e"]
            extra["ea_candidates"] = ea["n_cand"]
            extra["ea_mean_auc"] = ea["mean_auc"]
            extra["ea_generation_subnets_per_sec"] = ea["generation_value"]
            extra["ea_generation_blocks_computed_reused"] = ea["generation_blocks_computed_reused"]
            if world == 1:
                for name, tables in (("criteo_full_best", "capped"), ("criteo_full_best", "full"), ("kdd_xlarge", "capped")):
                    a2 = argparse.Namespace(config=name, tables=tables)
                    c2 = config_of(a2)
                    m2, t2, _h = build_training(env, c2)
                    r2 = measure_training(env, c2, t2, max(10, K // 2), W, profile_gemm=False)
                    key = name + ("_%s_tables" % tables if name == "criteo_full_best" else "")
                    extra[key + "_samples_per_sec"] = r2["value"]
                    extra[key + "_samples_per_sec_e2e"] = r2["e2e"]
                    del m2, t2
                    torch.cuda.empty_cache()
                # bf16 compute (configs[3]): the Avazu NASRec-Full best model with full-size tables, and the EA scoring
                import nasrec_b200
                with nasrec_b200.precision("bf16"):
                    c2 = config_of(argparse.Namespace(config="avazu_full_best", tables="full"))
                    m2, t2, _h = build_training(env, c2)
                    r2 = measure_training(env, c2, t2, max(10, K // 2), W, profile_gemm=False)
                    extra["avazu_full_best_bf16_samples_per_sec"] = r2["value"]
                    extra["avazu_full_best_bf16_samples_per_sec_e2e"] = r2["e2e"]
                    del m2, t2
                    torch.cuda.empty_cache()
                    ea16 = measure_ea(env, CONFIGS["ea"])
                    extra["ea_subnets_per_sec_%dx8192_bf16" % EA_BATCHES] = ea16["value"]
                extra["hbm_kernels"] = measure_hbm_kernels(env)
            if rank == 0:
                line["extra"] = extra
        if world == 1 and rank == 0 and not args.no_cpu:
            base, _ = reference_train(cfg, 12, 3)
            line["cpu_baseline"] = base
            if args.config == "small_supernet" and not args.no_extras and args.dtype == "f32":
                for name, tables in (("criteo_full_best", "capped"), ("criteo_full_best", "full")):
                    b2, _ = reference_train(config_of(argparse.Namespace(config=name, tables=tables)), 6, 2)
                    line["extra"]["%s_%s_tables_cpu_reference_samples_per_sec" % (name, tables)] = b2["value"]
                    line["extra"]["%s_%s_tables_cpu_kind" % (name, tables)] = b2["kind"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="small_supernet", choices=sorted(CONFIGS))
    ap.add_argument("--tables", default="capped", choices=["capped", "full"], help="criteo_full_best: embedding table sizes")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="GEMM arithmetic: f32 = 3xTF32-split (fp32 parity); bf16 = bf16 operands, fp32 accumulate (use_amp)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs' side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
