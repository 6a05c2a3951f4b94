"""Developer probe: launch latency floor of the segment GEMM (tiny shapes), warm back-to-back vs isolated,
tensor-core vs FFMA path, with and without another kernel interleaved (instruction-cache effect)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
def probe(M, N, K, mode):
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    big = torch.randn(4096, 1024, device=dev)
    sp, ns = _lib.segs([(x.data_ptr(), K, K, 0)])
    _lib.LIB.set_gemm_mode(mode)
    f = lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), K, 0, N, None, C.data_ptr(), N, M)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): f()
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 200 * 1e3
    iso = 0.0
    for _ in range(30):
        torch.cuda.synchronize(); e0.record(); f(); e1.record(); torch.cuda.synchronize(); iso += e0.elapsed_time(e1) * 1e3
    iso /= 30
    inter = 0.0
    for _ in range(30):
        big.mul_(1.0001); torch.relu_(big); torch.cuda.synchronize()
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); inter += e0.elapsed_time(e1) * 1e3
    inter /= 30
    print("M=%d N=%d K=%d mode=%d: back-to-back %.1f us, isolated %.1f us, after other kernels %.1f us" % (M, N, K, mode, b2b, iso, inter))
for shape in ((512, 16, 16), (512, 128, 128), (512, 1024, 1037), (512, 1024, 4109)):
    for mode in (3, 0):
        probe(*shape, mode)
