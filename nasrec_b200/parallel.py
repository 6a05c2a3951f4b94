"""Single-node multi-GPU: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference has no collectives at all (its only multi-GPU mechanism is one OS
process per EA candidate, searcher/searcher.py:126-156).  The hot path shards in
exactly two ways:

* data-parallel supernet training: same-seed choice on every rank, one all-reduce
  of the *active subnet's* dense gradients, an all-gather of the (ids, d_embedding)
  pairs followed by the same deterministic sorted-row Adagrad on every replica
  (replicas stay bit-identical), the clip norm computed on the already-global grads;
* EA candidate scoring: candidates partitioned across ranks, no data-path
  collective, one final gather of (loss, AUC, accuracy) per candidate.

The helpers are device-agnostic so the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import engine as eng
from .utils.train_utils import FusedTrainer


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of candidates (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_world(group=None) -> Tuple[int, int]:
    """(rank, world) of the current process; (0, 1) outside torch.distributed."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allgather_records(local: list, group=None) -> list:
    """Every rank contributes a list of per-candidate records (host objects); every rank gets
    the concatenation in rank order == candidate order (shards are contiguous)."""
    world = dist.get_world_size(group)
    parts: List[Optional[list]] = [None] * world
    dist.all_gather_object(parts, list(local), group=group)
    return [r for part in parts for r in part]


def allreduce_flat(tensors: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
    """Sum-all-reduce a list of tensors as ONE bucket; returns views into the bucket
    (replacing the inputs), so no copy-back pass is needed."""
    if not tensors:
        return []
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, o = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat[o:o + n].view(t.shape))
        o += n
    return out


def allgather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equally-shaped per-rank tensors along dim 0, in rank order."""
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def gather_results(local: Sequence[Sequence[float]], n_total: int, group=None) -> Optional[torch.Tensor]:
    """Final EA gather: every rank contributes [n_local, k] floats; rank 0 gets [n_total, k]
    in candidate order (shards are contiguous, so rank order == candidate order)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = len(local[0]) if len(local) else 0
    kk = torch.tensor([k], dtype=torch.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    kk = kk.to(dev)
    dist.all_reduce(kk, op=dist.ReduceOp.MAX, group=group)
    k = int(kk.item())
    pad = max(shard_range(n_total, world, r)[1] - shard_range(n_total, world, r)[0] for r in range(world))
    buf = torch.zeros(pad, k, dtype=torch.float64, device=dev)
    if len(local):
        buf[:len(local)] = torch.tensor(local, dtype=torch.float64, device=dev)
    allb = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf, group=group)
    if rank != 0:
        return None
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        rows.append(allb[r][:hi - lo])
    return torch.cat(rows).cpu()


class DataParallelTrainer(FusedTrainer):
    """Data-parallel fused step.  Every rank must seed numpy identically so that the
    sampled subnet (and therefore the gradient support) coincides on all ranks."""

    def __init__(self, model, lr, eps: float = 1e-2, clip: Optional[float] = 5.0, group=None):
        super().__init__(model, lr, eps, clip)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.defer_sparse = self.world > 1

    def step(self, int_x, cat_x, y, lr: Optional[float] = None):
        if self.world == 1:
            return super().step(int_x, cat_x, y, lr)
        # local loss is a mean over the local batch; scale so the summed gradient is the
        # gradient of the mean over the GLOBAL batch (BCEWithLogitsLoss 'mean', train_utils.py:266)
        logits, loss, run, raw = self.forward_backward(int_x, cat_x, y, grad_scale=1.0 / self.world)
        emb_ids = {id(m.weight) for m in self.model._embedding}
        dense = [h for h in run.touched() if h.g is not None and id(h.p) not in emb_ids]
        for h, g in zip(dense, allreduce_flat([h.g for h in dense], self.group)):
            h.g = g
        sparse = None
        if raw is not None:
            cat_l, gout_l = raw
            sparse = eng.reduce_sparse(allgather_cat(cat_l, self.group), allgather_cat(gout_l, self.group),
                                       getattr(self.model, "_tables", None))
        self.apply(run, sparse, lr)
        return logits, loss


class NativeDataParallelTrainer(DataParallelTrainer):
    """DataParallelTrainer on the C++ step executor (nasrec_b200/native.py).  The executor lays the
    step's dense parameter gradients out back to back in one arena, in first-touch order -- identical on
    every rank because the sampled subnet is -- so the flat bucket that is all-reduced IS the gradient
    storage: no concatenation, no copy back.  The embedding gradient is all-gathered raw and reduced by
    the same deterministic sorted-row kernel on every replica."""

    def __init__(self, model, lr, eps: float = 1e-2, clip: Optional[float] = 5.0, group=None):
        super().__init__(model, lr, eps, clip, group)
        from .native import NativeTrainer
        self._nt = NativeTrainer(model, lr, eps, clip)
        self._nt.late_join = False                        # the gradient bucket is all-reduced between backward and apply
        self._nt.state = self.state                       # one set of Adagrad accumulators for both paths
        self.overlap_comm = True
        self._cb_net = None
        self._pending: list = []

    def _seal(self, off: int, nbytes: int):
        """Called by the executor (on this thread) when gradient-bucket bytes [off, off+nbytes) are final."""
        chunk = self._cb_net.pg[off: off + nbytes].view(torch.float32)
        self._pending.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def step(self, int_x, cat_x, y, lr: Optional[float] = None):
        net = self._nt._native(int_x)
        if net is None:
            return super().step(int_x, cat_x, y, lr)
        if self.world == 1:
            out = self._nt.step(int_x, cat_x, y, lr)
            self.last_total_norm = self._nt.last_total_norm
            return out
        from . import _lib
        net.refresh()
        cat = (cat_x if cat_x.dtype == torch.int64 else cat_x.long()).contiguous()
        net.reserve(self.world * cat.shape[0])       # the sparse reduction will see the all-gathered batch
        with _lib.pin_stream():
            # Sealed ranges of the gradient bucket are all-reduced while backward is still running: NCCL works on
            # its own stream, ordered after the kernels already queued here (late blocks hold the widest weights and
            # run their backward first, so most of the exchange hides behind the rest of the backward pass).
            self._pending = []
            # the ids are known before the step starts: gather them behind the forward pass
            cat_all = torch.empty((self.world * cat.shape[0], cat.shape[1]), dtype=cat.dtype, device=cat.device)
            ids_handle = dist.all_gather_into_tensor(cat_all, cat, group=self.group, async_op=True)
            if self.overlap_comm and net is not self._cb_net:
                net.set_seal_callback(self._seal)
                self._cb_net = net
            self._nt._fork_on(net)
            try:
                logits, loss = net.forward_backward(self._nt._choice(), int_x.contiguous(), cat, y.contiguous(),
                                                    grad_scale=1.0 / self.world)
            finally:
                self._nt._fork_off()
            if self.overlap_comm:
                for h in self._pending:
                    h.wait()
            else:
                bucket = net.grad_bucket()
                if bucket.numel():
                    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group)
            raw = net.sparse_raw(cat)
            ids_handle.wait()
            if raw is not None:
                net.sparse_reduce(cat_all, allgather_cat(raw, self.group))
            norm = net.apply(self.lr if lr is None else lr, self.eps, self.clip)
        self.last_total_norm = norm[0:1]
        return logits, loss
