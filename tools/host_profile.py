"""GPU box: where does the HOST time of a supernet training step go? (cProfile over 30 steps)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nasrec_b200 import SuperNet, ops_config_lib, _lib
from nasrec_b200.utils.train_utils import FusedTrainer, init_weights
dev = torch.device("cuda")
ne = [min(x, 500000) for x in bench._CRITEO]
torch.manual_seed(0); np.random.seed(0)
m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
             path_sampling_strategy="default", anypath_choice="binomial-0.5").to(dev)
m.materialize(13); m.apply(init_weights)
tr = FusedTrainer(m, lr=0.12)
pool = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in bench.synth_pool(8, 512, 13, ne, 1)]
for i in range(10): tr.step(*pool[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50): tr.step(*pool[i % 8])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host-side issue time %.3f ms/step ; incl. drain %.3f ms/step" % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for i in range(30): tr.step(*pool[i % 8])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
