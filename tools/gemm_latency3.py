"""Warm-instruction-cache floor of the tensor-core GEMM: 200 identical tiny launches replayed as one CUDA graph
(no host cost), vs the same launches interleaved with an unrelated kernel (LayerNorm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
def run(M, N, K, interleave):
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    z = torch.randn(M, 1024, device=dev); y = torch.empty(M, 1024, device=dev); g = torch.ones(1024, device=dev)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    sp, ns = _lib.segs([(x.data_ptr(), K, K, 0)])
    def body():
        _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), K, 0, N, None, C.data_ptr(), N, M)
        if interleave:
            _lib.call("nasrec_ln_fwd", z.data_ptr(), 1024, M, 1024, g.data_ptr(), g.data_ptr(), 1e-5, 1, 1024, y.data_ptr(), 1024,
                      mean.data_ptr(), rstd.data_ptr(), 0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with _lib.pin_stream():
            for _ in range(3): body()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        with _lib.pin_stream():
            for _ in range(200): body()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1e3
_lib.LIB.set_gemm_mode(3)
for shape in ((512, 16, 16), (512, 128, 128), (512, 1024, 256), (512, 1024, 1037)):
    a = run(*shape, False); b = run(*shape, True)
    print("M=%d N=%d K=%d: gemm only %.2f us/launch; gemm+ln pair %.2f us" % (*shape, a, b))
# LN alone
def ln_only():
    M = 512
    z = torch.randn(M, 1024, device=dev); y = torch.empty(M, 1024, device=dev); g = torch.ones(1024, device=dev)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    f = lambda: _lib.call("nasrec_ln_fwd", z.data_ptr(), 1024, M, 1024, g.data_ptr(), g.data_ptr(), 1e-5, 1, 1024, y.data_ptr(), 1024, mean.data_ptr(), rstd.data_ptr(), 0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with _lib.pin_stream():
            f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        with _lib.pin_stream():
            for _ in range(200): f()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print("ln_fwd alone %.2f us/launch" % (e0.elapsed_time(e1) / 200 * 1e3))
ln_only()
