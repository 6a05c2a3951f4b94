"""Mirror of the step body in ``nasrec/utils/train_utils.py`` plus the fused trainer.

* ``init_weights`` / ``warmup_supernet_model`` / ``warmup_model`` / ``accuracy`` /
  ``get_l2_loss`` keep the reference's names and semantics (train_utils.py:70-127,
  393-433).
* ``reference_style_step`` is the reference's own step body (train_utils.py:262-286)
  run on the CUDA model with stock ``torch.optim.Adagrad`` / ``clip_grad_norm_`` --
  the drop-in path.
* ``FusedTrainer`` is the B200 fast path for the same step: forward tape -> fused
  BCE -> backward tape -> one global-norm reduction -> multi-tensor Adagrad on the
  dense parameters that took part + row-wise Adagrad on the touched embedding rows.
  It is mathematically the reference step when wd == 0: untouched rows
  and unused parameters have zero gradient, for which Adagrad is the identity.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import ctypes

import torch
import torch.nn as nn

from .. import _lib
from .. import engine as eng
from ..engine import Tape, Var
from ..supernet.modules import Run
from ..supernet.supernet import SuperNet


def init_weights(m):
    """train_utils.py:70-89 (exact type matches, so LayerNorm and MHA.out_proj keep their init)."""
    if type(m) == nn.Embedding:
        torch.nn.init.xavier_normal_(m.weight)
    elif type(m) == nn.Linear:
        torch.nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)
    elif type(m) == nn.MultiheadAttention:
        for p in m.parameters():
            if len(p.size()) > 1:
                torch.nn.init.xavier_uniform_(p)
            else:
                torch.nn.init.zeros_(p)


def get_l2_loss(model: nn.Module, reg: float, no_reg_param_name=None, gpu=None):
    """train_utils.py:91-115."""
    if reg == 0:
        return torch.tensor(0.0).to(gpu)
    reg_loss = None
    for n, m in model.named_parameters():
        if len(m.shape) == 1 or ((no_reg_param_name is not None) and n.startswith(no_reg_param_name)):
            continue
        r = torch.square(torch.norm(m, p=2)) * reg
        reg_loss = r if reg_loss is None else reg_loss + r
    return reg_loss


def accuracy(gt, pred):
    """train_utils.py:118-126."""
    pred_binary = torch.gt(pred, 0.5).float()
    return (pred_binary == gt).sum() / pred_binary.size(0)


def warmup_model(model: nn.Module, train_loader, gpu):
    """train_utils.py:393-410."""
    model = model.to(gpu)
    int_x, cat_x, _ = next(iter(train_loader))
    with torch.no_grad():
        model(int_x.to(gpu), cat_x.to(gpu))
    return model


def warmup_supernet_model(model: nn.Module, train_loader, gpu):
    """train_utils.py:413-433: one full-path forward materialises every lazy layer."""
    assert isinstance(model, SuperNet), NotImplementedError(
        "For 'warmup_supernet_model', the passed in model must be a 'SuperNet' object.")
    model = model.to(gpu)
    int_x, cat_x, _ = next(iter(train_loader))
    model.configure_path_sampling_strategy("full-path")
    with torch.no_grad():
        model(int_x.to(gpu), cat_x.to(gpu))
    return model


def reference_style_step(model, optimizer, loss_fn, int_x, cat_x, y, grad_clip_value: Optional[float] = 5.0):
    """train_utils.py:262-286 verbatim in structure, on the CUDA model."""
    optimizer.zero_grad()
    res = model(int_x, cat_x)
    loss = loss_fn(res, y)
    loss.backward()
    if grad_clip_value is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip_value)
    optimizer.step()
    return res, loss


def test_one_epoch(model, test_loader, loss_fn=None, gpu=None, max_steps: int = -1, use_amp: bool = False):
    """train_utils.py:129-178: forward over the evaluation loader, then accuracy / ROC-AUC /
    log-loss over the concatenated predictions -- computed on the device
    (``nasrec_binary_metrics``), so only three scalars reach the host.  ``loss_fn`` is accepted
    for signature compatibility; the metric is BCE-with-logits, the only loss the reference wires."""
    from ..search import binary_metrics_device
    if use_amp:          # train_utils.py:146 autocast: GEMMs on bf16 operands (nasrec_b200.precision)
        from .. import precision
        with precision("bf16"):
            return test_one_epoch(model, test_loader, loss_fn, gpu, max_steps, False)
    was_training = model.training
    model.eval()
    preds, labels = [], []
    with torch.no_grad():
        for counter, (int_x, cat_x, y) in enumerate(test_loader, start=1):
            preds.append(model(int_x.to(gpu, non_blocking=True), cat_x.to(gpu, non_blocking=True)).reshape(-1))
            labels.append(y.to(gpu, non_blocking=True).reshape(-1))
            if max_steps != -1 and counter >= max_steps:
                break
    model.train(was_training)
    res = binary_metrics_device(torch.cat(preds), torch.cat(labels))       # synchronises (three scalars reach the host)
    tables = getattr(model, "_tables", None)
    if tables is not None:
        tables.check()             # an out-of-range id raises IndexError, as nn.Embedding does in the reference
    return res


def train_and_test_one_epoch(model, epoch: int, optimizer, lr_scheduler, train_loader, test_loader, loss_fn,
                             l2_loss_fn, train_batch_size: int, gpu, display_interval: int = 100,
                             test_interval: int = 2000, max_train_steps: int = -1, max_eval_steps: int = -1,
                             test_only_at_last_step: bool = False, grad_clip_value: Optional[float] = None,
                             tb_writer=None, use_amp: bool = False):
    """train_utils.py:181-390 with the same arguments and ``logs`` dictionary.

    ``optimizer`` is either a torch optimizer -- the reference's own step body on the CUDA model
    (``reference_style_step`` + the L2 term) -- or a ``FusedTrainer``, which replaces that body by
    the fused step when ``l2_loss_fn`` is None / weight decay is 0 (the shipped recipes)."""
    from ..search import binary_metrics_device
    if use_amp:          # train_utils.py:247-286: autocast forward + (here unnecessary) loss scaling -> bf16 GEMM operands
        from .. import precision
        with precision("bf16"):
            return train_and_test_one_epoch(model, epoch, optimizer, lr_scheduler, train_loader, test_loader, loss_fn,
                                            l2_loss_fn, train_batch_size, gpu, display_interval, test_interval,
                                            max_train_steps, max_eval_steps, test_only_at_last_step, grad_clip_value,
                                            tb_writer, False)
    logs = {k: [] for k in ("train_loss", "train_AUROC", "train_Accuracy", "test_loss", "test_AUROC", "test_Accuracy",
                            "epoch", "iters")}
    fused = isinstance(optimizer, FusedTrainer)
    if fused and l2_loss_fn is not None:
        # The fused step has no L2 term (train_utils.py:283 adds get_l2_loss to the loss; main_train.py:317 defaults
        # --wd to 1e-8).  Training on without it would silently differ from the reference: refuse instead.
        with torch.no_grad():
            l2_now = float(l2_loss_fn(model))
        if l2_now != 0.0:
            raise ValueError("the fused step (FusedTrainer / NativeTrainer) has no weight-decay term but l2_loss_fn(model) = %g; "
                             "pass wd = 0 (l2_loss_fn=None) or a torch optimizer, which runs the reference's own step "
                             "body including the L2 loss" % l2_now)
    model.train()
    for batch_num, (int_x, cat_x, y) in enumerate(train_loader):
        int_x, cat_x, y = (t.to(gpu, non_blocking=True) for t in (int_x, cat_x, y))
        full = len(y) == train_batch_size                       # the ragged last batch is only evaluated
        if fused:
            lr = None
            if lr_scheduler is not None:
                lr = lr_scheduler.get_last_lr()[0] if hasattr(lr_scheduler, "get_last_lr") else lr_scheduler.lr()
            if full:
                res, loss = optimizer.step(int_x, cat_x, y, lr=lr)
            else:
                with torch.no_grad():
                    res = model(int_x, cat_x)
                    loss = torch.nn.functional.binary_cross_entropy_with_logits(res, y)
            l2_loss = torch.zeros((), device=res.device)
        else:
            optimizer.zero_grad()
            res = model(int_x, cat_x)
            loss = loss_fn(res, y)
            l2_loss = l2_loss_fn(model) if l2_loss_fn is not None else torch.zeros((), device=res.device)
            if full:
                (loss + l2_loss).backward()
                if grad_clip_value is not None:
                    torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip_value)
                optimizer.step()
        last = batch_num == max_train_steps - 1
        if batch_num % display_interval == 0 or last:
            if bool(torch.isnan(loss)):                          # train_utils.py:294-300 (diverged KDD runs)
                logs["test_loss"].append(999.99)
                logs["test_AUROC"].append(-1)
                logs["test_Accuracy"].append(-1)
                return logs
            yv = y.detach().reshape(-1)
            if bool((yv == yv[0]).all()):                        # sklearn raises on one-class batches
                train_acc = float(((res.detach().reshape(-1) > 0).float() == yv).float().mean())
                train_auroc = 1.0
            else:
                train_acc, train_auroc, _ = binary_metrics_device(res.detach(), yv)
            logs["train_loss"].append(float(loss.detach()))
            logs["train_AUROC"].append(train_auroc)
            logs["train_Accuracy"].append(train_acc)
            logs["epoch"].append(epoch)
            logs["iters"].append(batch_num)
            if tb_writer is not None:
                tb_writer.add_scalar("Loss/train/epoch{}".format(epoch), float(loss), batch_num * train_batch_size)
        if (batch_num % test_interval == 0 or last) and ((not test_only_at_last_step) or last):
            test_acc, test_auroc, test_loss = test_one_epoch(model, test_loader, loss_fn, gpu, max_steps=max_eval_steps)
            logs["test_loss"].append(test_loss)
            logs["test_AUROC"].append(test_auroc)
            logs["test_Accuracy"].append(test_acc)
            if tb_writer is not None:
                tb_writer.add_scalar("Loss/test/epoch{}".format(epoch), test_loss, batch_num * train_batch_size)
            model.train()
        if max_train_steps != -1 and batch_num >= max_train_steps - 1:
            return logs
        if lr_scheduler is not None:
            lr_scheduler.step() if not hasattr(lr_scheduler, "advance") else lr_scheduler.advance()
    return logs


class FusedTrainer:
    """Fused step for SuperNet training (weight sharing or fixed), wd == 0."""

    def __init__(self, model: SuperNet, lr: float, eps: float = 1e-2, clip: Optional[float] = 5.0):
        self.model = model
        self.lr = lr
        self.eps = eps
        self.clip = clip
        self.state: Dict[int, torch.Tensor] = {}
        self._emb_state: Optional[List[torch.Tensor]] = None
        self._emb_ptrs = None
        self.last_total_norm: Optional[torch.Tensor] = None
        self.defer_sparse = False
        # set by GraphedFusedTrainer during capture: weight-gradient GEMMs fork onto this stream
        self.side_stream: Optional[torch.cuda.Stream] = None
        # pre-split weight planes (nasrec_b200/planes.py): forward / dgrad GEMMs take the TMA-fed kernel, the Adagrad
        # kernel keeps the planes in step.  Costs 2x the dense weights in HBM; False = LDG-producer kernel everywhere.
        self.use_planes = True
        self._plane_params = None

    def _state_of(self, p: torch.Tensor) -> torch.Tensor:
        s = self.state.get(id(p))
        if s is None:
            s = torch.zeros_like(p)
            self.state[id(p)] = s
        return s

    def _emb_tables(self):
        ws = [m.weight for m in self.model._embedding]
        key = tuple(w.data_ptr() for w in ws)
        if self._emb_ptrs is None or self._emb_ptrs[0] != key:
            dev = ws[0].device
            states = [self._state_of(w) for w in ws]
            self._emb_ptrs = (key, torch.tensor(list(key), dtype=torch.int64, device=dev),
                              torch.tensor([s.data_ptr() for s in states], dtype=torch.int64, device=dev))
        return self._emb_ptrs[1], self._emb_ptrs[2]

    def forward_backward(self, int_x, cat_x, y, grad_scale: float = 1.0):
        """One forward + backward; returns (logits, loss[1], run, sparse grads)."""
        model = self.model
        macro, micro = model._sample()
        if model._needs_materialize():
            model.materialize(int_x.shape[1])
        tape = Tape(True)
        sink = eng.SparseSink()
        sink.defer = self.defer_sparse
        run = Run(tape, sparse_sink=sink)
        cat = cat_x if cat_x.dtype == torch.int64 else cat_x.long()
        eng.PLANES = self._planes()            # forward and dgrad GEMMs fetch pre-split weight planes by TMA
        try:
            return self._forward_backward(model, run, tape, sink, int_x, cat, y, macro, micro, grad_scale)
        finally:
            eng.PLANES = None

    def _planes(self):
        """The model's PlaneCache with every dense 2-D weight's planes in step (apply() keeps them so; anything else that
        touched a weight is noticed through Tensor._version / data_ptr / the library's epoch and rebuilt here)."""
        if not self.use_planes:
            return None
        from ..planes import PlaneCache
        cache = PlaneCache.of(self.model)
        if self._plane_params is None:
            emb = {id(m.weight) for m in self.model._embedding}
            self._plane_params = [p for p in self.model.parameters() if id(p) not in emb and PlaneCache.wanted(p)]
        cache.sync(self._plane_params)
        return cache

    def _forward_backward(self, model, run, tape, sink, int_x, cat, y, macro, micro, grad_scale):
        logits = model._run_network(run, Var(int_x.contiguous()), cat.contiguous(), macro, micro)
        loss, dl = eng.bce_with_logits(logits.t, y, grad_scale=grad_scale)
        logits.g = dl
        if self.side_stream is None:
            tape.backward()
        else:
            lib = _lib.LIB.load().cdll
            eng.OVERLAP_KEEP = keep = []
            lib.nasrec_set_side_stream(ctypes.c_void_p(self.side_stream.cuda_stream))
            try:
                tape.backward()
                rc = lib.nasrec_side_join(ctypes.c_void_p(_lib.stream_ptr()))
            finally:
                lib.nasrec_set_side_stream(None)
                eng.OVERLAP_KEEP = None
            if rc != 0:
                raise RuntimeError("nasrec_side_join failed: %d" % rc)
            self._keep = keep            # released when the next step replaces it (after the join in stream order)
        return logits.t, loss, run, (sink[0] if sink else None)

    def apply(self, run: Run, sparse, lr: Optional[float] = None):
        """Global-norm clip + Adagrad on everything that received a gradient."""
        lr = self.lr if lr is None else lr
        emb_ids = {id(m.weight) for m in self.model._embedding}
        dense = [h for h in run.touched() if h.g is not None and id(h.p) not in emb_ids]
        dev = self.model._final.weight.device
        out = torch.empty(2, dtype=torch.float32, device=dev)
        sizes = [h.g.numel() for h in dense]
        if self.clip is not None:
            gp = _lib.ptr_array([h.g.data_ptr() for h in dense])
            sz = _lib.i64_array(sizes)
            nws = _lib.query("nasrec_sumsq_ws_floats", sz, len(dense))
            partial = torch.empty(nws, dtype=torch.float32, device=dev)
            _lib.call("nasrec_grad_norm_clip", gp, sz, len(dense),
                      sparse.sumsq.data_ptr() if sparse is not None else None,
                      sparse.F if sparse is not None else 0, float(self.clip), partial.data_ptr(), out.data_ptr())
            coef = out.data_ptr() + 4
            self.last_total_norm = out[0:1]
        else:
            coef = None
        if dense:
            cache = None
            if self.use_planes:
                from ..planes import PlaneCache
                cache = self.model.__dict__.get(PlaneCache.ATTR)
            pls = [cache.planes(h.p) if cache is not None else None for h in dense]
            if any(pl is not None for pl in pls):
                # same update, and the hi / lo planes of every tensor that has them are rewritten in the same pass
                _lib.call("nasrec_adagrad_multi_planes", _lib.ptr_array([h.t.data_ptr() for h in dense]),
                          _lib.ptr_array([h.g.data_ptr() for h in dense]),
                          _lib.ptr_array([self._state_of(h.p).data_ptr() for h in dense]), _lib.i64_array(sizes),
                          len(dense), float(lr), float(self.eps), coef,
                          _lib.ptr_array([pl[0].data_ptr() if pl else 0 for pl in pls]),
                          _lib.ptr_array([pl[1].data_ptr() if pl else 0 for pl in pls]),
                          _lib.i32_array([h.t.shape[1] if (pl and h.t.dim() == 2) else 0 for h, pl in zip(dense, pls)]),
                          _lib.i32_array([pl[3] if pl else 0 for pl in pls]),
                          _lib.i64_array([pl[2] if pl else 0 for pl in pls]))
            else:
                _lib.call("nasrec_adagrad_multi", _lib.ptr_array([h.t.data_ptr() for h in dense]),
                          _lib.ptr_array([h.g.data_ptr() for h in dense]),
                          _lib.ptr_array([self._state_of(h.p).data_ptr() for h in dense]), _lib.i64_array(sizes),
                          len(dense), float(lr), float(self.eps), coef)
        if sparse is not None:
            tp, sp = self._emb_tables()
            _lib.call("nasrec_emb_rowwise_adagrad", sparse.uniq.data_ptr(), sparse.nuniq.data_ptr(),
                      sparse.row_grad.data_ptr(), tp.data_ptr(), sp.data_ptr(), sparse.B, sparse.F, float(lr),
                      float(self.eps), coef)

    def prime(self, int_x, cat_x, y):
        """Trigger every lazy host-side initialisation (pointer tables, optimizer state,
        kernel attributes) WITHOUT changing the model: one forward/backward whose gradients
        are dropped.  Needed before CUDA-graph capture of step()."""
        _, _, run, _ = self.forward_backward(int_x, cat_x, y)
        for h in run.touched():
            if h.g is not None or h.req:
                self._state_of(h.p)
        self._emb_tables()

    def step(self, int_x, cat_x, y, lr: Optional[float] = None):
        with _lib.pin_stream():
            logits, loss, run, sparse = self.forward_backward(int_x, cat_x, y)
            self.apply(run, sparse, lr)
        return logits, loss
