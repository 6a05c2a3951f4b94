"""Is the native training step host- or device-bound?  Host enqueue time per step (no sync) vs drained time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nasrec_b200 import SuperNet, ops_config_lib
from nasrec_b200.native import NativeTrainer
from nasrec_b200.utils.train_utils import init_weights
dev = torch.device("cuda", 0)
ne = [min(x, bench.CAP) for x in bench._CRITEO]
torch.manual_seed(1234); np.random.seed(1234)
m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
             path_sampling_strategy="default", anypath_choice="binomial-0.5", supernet_training_steps=0).to(dev)
m.materialize(13); m.apply(init_weights)
tr = NativeTrainer(m, lr=0.12)
pool = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in bench.synth_pool(32, 512, 13, ne, seed=1)]
for i in range(10): tr.step(*pool[i % 32])
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for i in range(N): tr.step(*pool[i % 32])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.3f ms/step; drained %.3f ms/step" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
# split: python-side vs C call
import nasrec_b200.native as nat
tc = [0.0]
orig = nat.NativeNet.forward_backward
def timed(self, *a, **k):
    s = time.perf_counter(); r = orig(self, *a, **k); tc[0] += time.perf_counter() - s; return r
nat.NativeNet.forward_backward = timed
t0 = time.perf_counter()
for i in range(N): tr.step(*pool[i % 32])
t1 = time.perf_counter()
torch.cuda.synchronize()
print("of which forward_backward C call %.3f ms/step (host), rest of step() %.3f ms" % (tc[0] / N * 1e3, ((t1 - t0) - tc[0]) / N * 1e3))

# inside the C call: time in cudaLaunchKernelEx vs the executor's own host code
from nasrec_b200 import _lib
nat.NativeNet.forward_backward = orig
_lib.query("nasrec_host_prof", 1)
t0 = time.perf_counter()
for i in range(N): tr.step(*pool[i % 32])
t1 = time.perf_counter()
torch.cuda.synchronize()
ns, cnt = _lib.query("nasrec_host_prof", 2), _lib.query("nasrec_host_prof", 3)
_lib.query("nasrec_host_prof", 0)
print("per step: %.1f launches, %.3f ms inside cudaLaunchKernelEx (%.2f us each), step host total %.3f ms" % (cnt / N, ns / N / 1e6, ns / max(cnt, 1) / 1e3, (t1 - t0) / N * 1e3))
print("tensor maps: hits %d encodes %d tma launches %d" % tuple(_lib.query("nasrec_tensor_map_stats", i) for i in range(3)))

# true host cost: time 2 steps right after a synchronize (the launch queue is empty, nothing blocks on the device)
hs = []
for rep in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.step(*pool[rep % 32]); tr.step(*pool[(rep + 1) % 32])
    hs.append((time.perf_counter() - t0) / 2)
    torch.cuda.synchronize()
hs.sort()
print("host cost of a step with an empty launch queue: median %.3f ms (min %.3f)" % (hs[len(hs) // 2] * 1e3, hs[0] * 1e3))
_lib.query("nasrec_host_prof", 1)
torch.cuda.synchronize()
for rep in range(20):
    tr.step(*pool[rep % 32]); tr.step(*pool[(rep + 1) % 32])
    torch.cuda.synchronize()
ns, cnt = _lib.query("nasrec_host_prof", 2), _lib.query("nasrec_host_prof", 3)
print("  with an empty queue: %.2f us per cudaLaunchKernelEx over %d launches" % (ns / max(cnt, 1) / 1e3, cnt))
