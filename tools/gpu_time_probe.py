"""GPU box: device-only time of a supernet training step, measured by CUDA-graph replay of steps with
pinned choices (no host in the loop), next to the eager (host-issued) time of the same choices."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nasrec_b200 import SuperNet, ops_config_lib, _lib
from nasrec_b200.utils.train_utils import FusedTrainer, init_weights
from nasrec_b200.utils.graph import GraphedFusedTrainer
dev = torch.device("cuda")
ops = sys.argv[1] if len(sys.argv) > 1 else "autoctr"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ne = [min(x, 500000) for x in bench._CRITEO]
torch.manual_seed(0); np.random.seed(0)
m = SuperNet(num_blocks=7, ops_config=ops_config_lib[ops], use_layernorm=True, num_embeddings=ne,
             path_sampling_strategy="default", anypath_choice="binomial-0.5").to(dev)
m.materialize(13); m.apply(init_weights)
pool = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in bench.synth_pool(4, B, 13, ne, 1)]
tr = FusedTrainer(m, lr=0.12)
for i in range(5): tr.step(*pool[i % 4])
choices = []
for i in range(6):
    m._sample(); choices.append(m.choice)
m.configure_path_sampling_strategy("fixed-path")
g_ms, e_ms = [], []
for ch in choices:
    m.configure_choice(ch)
    gt = GraphedFusedTrainer(FusedTrainer(m, lr=0.12))
    gt.step(*pool[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): gt.step(*pool[i % 4])
    e1.record(); torch.cuda.synchronize()
    g_ms.append(e0.elapsed_time(e1) / 20)
    et = FusedTrainer(m, lr=0.12)
    for i in range(3): et.step(*pool[i % 4])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(20): et.step(*pool[i % 4])
    torch.cuda.synchronize(); e_ms.append((time.perf_counter() - t0) / 20 * 1e3)
print("%s B=%d: device-only (graph replay) ms/step per choice: %s  mean %.3f" % (ops, B, ["%.2f" % x for x in g_ms], np.mean(g_ms)))
print("%s B=%d: eager (host-issued)        ms/step per choice: %s  mean %.3f" % (ops, B, ["%.2f" % x for x in e_ms], np.mean(e_ms)))
