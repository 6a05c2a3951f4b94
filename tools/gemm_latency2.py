import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
M, N, K = 512, 16, 16
x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
sp, ns = _lib.segs([(x.data_ptr(), K, K, 0)])
for mode in (3, 0):
    _lib.LIB.set_gemm_mode(mode)
    f = lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), K, 0, N, None, C.data_ptr(), N, M)
    for _ in range(10): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): f()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("mode %d: host enqueue %.2f us/call, total incl. drain %.2f us/call" % (mode, (t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
# reference: a trivial torch kernel
y = torch.zeros(64, device=dev)
for _ in range(10): y.add_(1.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): y.add_(1.0)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("torch add_: host %.2f us/call, total %.2f us/call" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
# act kernel through the library (tiny param list)
a = torch.randn(512, 16, device=dev); b = torch.empty_like(a)
g = lambda: _lib.call("nasrec_act_fwd", a.data_ptr(), 16, 512, 16, 1, b.data_ptr(), 16, 0)
for _ in range(10): g()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): g()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("nasrec_act_fwd: host %.2f us/call, total %.2f us/call" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
