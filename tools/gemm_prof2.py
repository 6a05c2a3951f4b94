"""Device-bound GEMM timing: the fwd / dgrad / wgrad entry points captured in a CUDA graph (20 calls per replay) so
that the host enqueue cost does not cap the measurement.  usage: gemm_prof2.py M N K1 [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
M = int(sys.argv[1]); N = int(sys.argv[2]); K1 = int(sys.argv[3]); mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
K0 = 13
Ktot = K0 + K1
x0 = torch.randn(M, 16, device=dev); x1 = torch.randn(M, K1, device=dev)
W = torch.randn(N, Ktot, device=dev) / 32
ldp = (Ktot + 3 + 3) & ~3
hi = torch.zeros(N, ldp, device=dev); lo = torch.zeros(N, ldp, device=dev)
_lib.call("nasrec_planes_refresh", W.data_ptr(), Ktot, N, Ktot, K0, hi.data_ptr(), lo.data_ptr(), ldp)
C = torch.empty(M, N, device=dev); dC = torch.randn(M, N, device=dev)
dx0 = torch.empty_like(x0); dx1 = torch.empty_like(x1); dW = torch.zeros_like(W)
ws = torch.empty(64 << 20, device=dev)
_lib.LIB.load(); _lib.LIB.cdll.nasrec_set_workspace(ws.data_ptr(), ws.numel())
sp, ns = _lib.segs([(x0.data_ptr(), 16, K0, 0), (x1.data_ptr(), K1, K1, K0)])
dsp, _ = _lib.segs([(dx0.data_ptr(), 16, K0, 0), (dx1.data_ptr(), K1, K1, K0)])
_lib.LIB.set_gemm_mode(mode)
_lib.LIB.set_weight_planes(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, N, Ktot, K0)
ops = (("fwd", lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, None, C.data_ptr(), N, M)),
       ("dgrad", lambda: _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, 0, dsp, ns, M, 0)),
       ("wgrad", lambda: _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
REP = 20
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for name, fn in ops:
        for _ in range(3): fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(REP): fn()
        g.replay(); st.synchronize()
        e0.record(st)
        for _ in range(5): g.replay()
        e1.record(st); st.synchronize()
        us = e0.elapsed_time(e1) / (5 * REP) * 1e3
        print("graph M=%d N=%d K=%d mode=%d dbg=%s bn=%s %s: %.2f us  %.1f TFLOP/s" % (M, N, Ktot, mode, os.environ.get("NASREC_GEMM_DBG", "0"), os.environ.get("NASREC_TC_BN", "-"), name, us, 2.0 * M * N * Ktot / us / 1e6), flush=True)
