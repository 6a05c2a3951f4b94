"""Single-shape driver (also for ncu): the tensor-core GEMM fwd / dgrad / wgrad on workload-like operands
(aligned activations, reference-layout weights with the 1037-float row stride + their hi/lo planes), timed with
the TMA-fed kernel and with the LDG-producer kernel.  usage: gemm_prof.py [M] [mode] [N]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
K0, K1 = 13, 1024
Ktot = K0 + K1                                  # 1037: the reference's row stride for block 1
x0 = torch.randn(M, 16, device=dev); x1 = torch.randn(M, K1, device=dev)
W = torch.randn(N, Ktot, device=dev) / 32
ldp = (Ktot + 3 + 3) & ~3
hi = torch.zeros(N, ldp, device=dev); lo = torch.zeros(N, ldp, device=dev)
_lib.call("nasrec_planes_refresh", W.data_ptr(), Ktot, N, Ktot, K0, hi.data_ptr(), lo.data_ptr(), ldp)
C = torch.empty(M, N, device=dev); dC = torch.randn(M, N, device=dev)
dx0 = torch.empty_like(x0); dx1 = torch.empty_like(x1); dW = torch.zeros_like(W)
sp, ns = _lib.segs([(x0.data_ptr(), 16, K0, 0), (x1.data_ptr(), K1, K1, K0)])
dsp, _ = _lib.segs([(dx0.data_ptr(), 16, K0, 0), (dx1.data_ptr(), K1, K1, K0)])
_lib.LIB.set_gemm_mode(mode)
_lib.LIB.set_weight_planes(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, N, Ktot, K0)
ops = (("fwd", lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, None, C.data_ptr(), N, M)),
       ("dgrad", lambda: _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, 0, dsp, ns, M, 0)),
       ("wgrad", lambda: _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for tma in (1, 0):
    _lib.LIB.set_gemm_tma(bool(tma))
    for _ in range(3):
        for _n, fn in ops: fn()
    torch.cuda.synchronize()
    for name, fn in ops:
        t0 = time.perf_counter()
        e0.record()
        for _ in range(50): fn()
        e1.record()
        host = (time.perf_counter() - t0) / 50 * 1e6
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print("M=%d N=%d mode=%d tma=%d %s: %.1f us  %.1f TFLOP/s  (host %.1f us/call)" % (M, N, mode, tma, name, us, 2.0 * M * N * Ktot / us / 1e6, host), flush=True)
print("tensor maps: hits %d encodes %d tma launches %d" % tuple(_lib.query("nasrec_tensor_map_stats", i) for i in range(3)))
