"""Single-shape driver for ncu: runs the tcgen05 GEMM a few times (fwd, dgrad, wgrad) on
workload-like operands (aligned activations, reference-layout weights with odd row stride)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, K0, K1 = 1024, 13, 1024
Ktot = K0 + K1                                  # 1037: the reference's row stride for block 1
x0 = torch.randn(M, K0, device=dev); x1 = torch.randn(M, K1, device=dev)
W = torch.randn(N, Ktot, device=dev) / 32
C = torch.empty(M, N, device=dev); dC = torch.randn(M, N, device=dev)
dx0 = torch.empty_like(x0); dx1 = torch.empty_like(x1); dW = torch.zeros_like(W)
sp, ns = _lib.segs([(x0.data_ptr(), K0, K0, 0), (x1.data_ptr(), K1, K1, K0)])
dsp, _ = _lib.segs([(dx0.data_ptr(), K0, K0, 0), (dx1.data_ptr(), K1, K1, K0)])
_lib.LIB.set_gemm_mode(mode)
def run():
    _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, None, C.data_ptr(), N, M)
    _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, 0, dsp, ns, M, 0)
    _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, None, C.data_ptr(), N, M)),
                 ("dgrad", lambda: _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, 0, dsp, ns, M, 0)),
                 ("wgrad", lambda: _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0))):
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print("M=%d mode=%d %s: %.1f us  %.1f TFLOP/s" % (M, mode, name, us, 2.0 * M * N * Ktot / us / 1e6))
