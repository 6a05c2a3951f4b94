"""Developer probe (torchrun, N GPUs): where does a data-parallel step spend its time?
Host wall time and device time per phase of DataParallelTrainer.step, rank 0 and max over ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from nasrec_b200 import SuperNet, ops_config_lib, _lib
from nasrec_b200 import engine as eng
from nasrec_b200.parallel import DataParallelTrainer, allreduce_flat, allgather_cat
from nasrec_b200.utils.train_utils import init_weights

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ne = [min(x, bench.CAP) for x in bench._CRITEO]
torch.manual_seed(1234); np.random.seed(1234)
model = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
                 path_sampling_strategy="default", anypath_choice="binomial-0.5", supernet_training_steps=0).to(dev)
model.materialize(13); model.apply(init_weights)
tr = DataParallelTrainer(model, lr=0.12)
pool = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in bench.synth_pool(32, 512, 13, ne, seed=1234 + rank)]
for i in range(5):
    tr.step(*pool[i])
dist.barrier(); torch.cuda.synchronize()
names = ["fwd_bwd", "allreduce", "allgather", "reduce_sparse", "apply"]
host = {n: 0.0 for n in names}; devt = {n: 0.0 for n in names}
K = 30
for i in range(K):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    t = [0.0] * 6
    int_x, cat_x, y = pool[(5 + i) % 32]
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs[0].record(); t[0] = time.perf_counter()
    logits, loss, run, raw = tr.forward_backward(int_x, cat_x, y, grad_scale=1.0 / world)
    evs[1].record(); t[1] = time.perf_counter()
    emb_ids = {id(m.weight) for m in model._embedding}
    dense = [h for h in run.touched() if h.g is not None and id(h.p) not in emb_ids]
    for h, g in zip(dense, allreduce_flat([h.g for h in dense])):
        h.g = g
    evs[2].record(); t[2] = time.perf_counter()
    cat_l, gout_l = raw
    gc, gg = allgather_cat(cat_l), allgather_cat(gout_l)
    evs[3].record(); t[3] = time.perf_counter()
    sparse = eng.reduce_sparse(gc, gg)
    evs[4].record(); t[4] = time.perf_counter()
    tr.apply(run, sparse, None)
    evs[5].record(); t[5] = time.perf_counter()
    torch.cuda.synchronize()
    for k, n in enumerate(names):
        host[n] += (t[k + 1] - t[k]) * 1e3
        devt[n] += evs[k].elapsed_time(evs[k + 1])
out = torch.tensor([host[n] / K for n in names] + [devt[n] / K for n in names], dtype=torch.float64, device=dev)
mx = out.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    print("cpus", os.cpu_count(), "world", world, "dense tensors", len(dense), "dense floats", sum(h.g.numel() for h in dense))
    for k, n in enumerate(names):
        print("%-14s host %.3f ms (max %.3f)   device-span %.3f ms (max %.3f)" % (n, out[k], mx[k], out[5 + k], mx[5 + k]))
dist.destroy_process_group()
