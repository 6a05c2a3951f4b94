"""Debug driver: every GEMM entry point through the TMA path on small shapes, error vs fp64 with a coarse error map."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
torch.manual_seed(0)

def planes(W, first=0):
    rows, cols = W.shape
    ldp = (cols + 6) & ~3
    hi = torch.zeros(rows, ldp, device=dev); lo = torch.zeros(rows, ldp, device=dev)
    _lib.call("nasrec_planes_refresh", W.data_ptr(), cols, rows, cols, first, hi.data_ptr(), lo.data_ptr(), ldp)
    _lib.LIB.set_weight_planes(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, rows, cols, first)
    return hi, lo

def report(name, got, ref):
    torch.cuda.synchronize()
    got = got.double(); ref = ref.double()
    e = (got - ref).abs()
    rel = float(e.max() / ref.abs().max())
    msg = "%-28s rel %.2e" % (name, rel)
    if rel > 1e-5:
        g2 = got.reshape(-1, got.shape[-1]) if got.dim() > 1 else got.reshape(1, -1)
        e2 = e.reshape(g2.shape)
        bad_r = (e2.max(dim=1).values > 1e-4 * float(ref.abs().max())).nonzero().flatten()
        bad_c = (e2.max(dim=0).values > 1e-4 * float(ref.abs().max())).nonzero().flatten()
        msg += "  bad rows %d/%d [%s..] bad cols %d/%d [%s..]  got_absmax %.3g ref_absmax %.3g" % (
            len(bad_r), e2.shape[0], bad_r[:6].tolist(), len(bad_c), e2.shape[1], bad_c[:6].tolist(), float(got.abs().max()), float(ref.abs().max()))
    print(msg, flush=True)

tma0 = lambda: _lib.query("nasrec_tensor_map_stats", 2)
def run(M, N, K):
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); dC = torch.randn(M, N, device=dev)
    keep = planes(W)
    sp, ns = _lib.segs([(x.data_ptr(), K, K, 0)])
    C = torch.zeros(M, N, device=dev)
    t = tma0()
    _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), K, 0, N, None, C.data_ptr(), N, M)
    report("fwd M%d N%d K%d tma=%d" % (M, N, K, tma0() - t), C, x.double() @ W.double().t())
    dx = torch.zeros(M, K, device=dev)
    dsp, _ = _lib.segs([(dx.data_ptr(), K, K, 0)])
    t = tma0()
    _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), K, 0, dsp, ns, M, 0)
    report("dgrad M%d N%d K%d tma=%d" % (M, N, K, tma0() - t), dx, dC.double() @ W.double())
    dW = torch.zeros(N, K, device=dev)
    t = tma0()
    _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), K, 0, M, 0)
    report("wgrad M%d N%d K%d tma=%d" % (M, N, K, tma0() - t), dW, dC.double().t() @ x.double())

def run3(B, P, R):
    E = 16
    x = torch.randn(B, R, E, device=dev); W = torch.randn(P, R, device=dev); dZ = torch.randn(B, P, E, device=dev)
    keep = planes(W)
    sp, ns = _lib.segs([(x.data_ptr(), R * E, R, 0)])
    Z = torch.zeros(B, P, E, device=dev)
    t = tma0()
    _lib.call("nasrec_sproj_fwd", sp, ns, W.data_ptr(), R, P, None, Z.data_ptr(), P * E, B)
    report("sfwd B%d P%d R%d tma=%d" % (B, P, R, tma0() - t), Z, torch.einsum("pr,bre->bpe", W.double(), x.double()))
    dx = torch.zeros_like(x)
    dsp, _ = _lib.segs([(dx.data_ptr(), R * E, R, 0)])
    t = tma0()
    _lib.call("nasrec_sproj_dgrad", dZ.data_ptr(), P * E, P, W.data_ptr(), R, dsp, ns, B, 0)
    report("sdgrad B%d P%d R%d tma=%d" % (B, P, R, tma0() - t), dx, torch.einsum("bpe,pr->bre", dZ.double(), W.double()))
    dW = torch.zeros_like(W)
    ws = torch.empty(_lib.query("nasrec_sproj_wgrad_ws_floats", P, R, B), device=dev)
    t = tma0()
    _lib.call("nasrec_sproj_wgrad", dZ.data_ptr(), P * E, P, sp, ns, dW.data_ptr(), R, B, 0, ws.data_ptr())
    report("swgrad B%d P%d R%d tma=%d" % (B, P, R, tma0() - t), dW, torch.einsum("bpe,bre->pr", dZ.double(), x.double()))

for mode in (1, 3):
    _lib.LIB.set_gemm_mode(mode)
    print("mode", mode)
    run(128, 32, 32)
    run(128, 32, 64)
    run(256, 128, 256)
    run3(8, 32, 32)
    run3(16, 64, 64)
    run3(64, 32, 32)
