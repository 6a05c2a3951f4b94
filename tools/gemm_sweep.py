"""Debug sweep: seg-linear fwd / dgrad(accumulate) / wgrad against fp64 over model-like shapes under the GEMM plan forced
by NASREC_TC_BN / NASREC_TC_NS; prints the worst elementwise-relative-to-max error per case."""
import sys, os, itertools
import os.path
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import _lib as L
L.LIB.ensure_workspace()
g = torch.Generator().manual_seed(1)
def rel(a, ref): return float((a.double() - ref).abs().max() / ref.abs().max())
cases = [(512, 1024, [1035]), (512, 1024, [1033]), (512, 1024, [1040]), (512, 1024, [1037]), (512, 1024, [1056]), (512, 1024, [1028]), (512, 1024, [1048]), (512, 128, [1024]), (512, 64, [416]), (512, 64, [386]), (512, 1024, [13, 1024]), (512, 256, [13, 416, 1024]),
         (512, 16, [1024, 1024]), (512, 1024, [128]), (512, 128, [128, 1024, 1024, 1024]), (512, 45, [26 * 16])]
if os.environ.get('SWEEP_FIRST'): cases = cases[:int(os.environ['SWEEP_FIRST'])]
for M, N, widths in cases:
    offs, o = [], 0
    for j, w in enumerate(widths):
        offs.append(o); o += 13 if (j == 0 and w == 13) else ((w + 3) & ~3)
    Ktot = o if len(widths) > 1 else widths[0]
    xs = [torch.randn(M, (w + 3) & ~3, generator=g).cuda() for w in widths]
    W = (torch.randn(N, Ktot, generator=g) / np.sqrt(sum(widths))).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ldc = (N + 3) & ~3
    dC = torch.randn(M, ldc, generator=g).cuda()[:, :N]
    first = 13 if widths[0] == 13 else 0
    ldp = (Ktot + ((4 - first % 4) % 4) + 3) & ~3
    hi = torch.zeros(N, ldp, device="cuda"); lo = torch.zeros(N, ldp, device="cuda")
    L.call("nasrec_planes_refresh", W.data_ptr(), Ktot, N, Ktot, first, hi.data_ptr(), lo.data_ptr(), ldp)
    L.LIB.set_weight_planes(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, N, Ktot, first)
    sp, ns = L.segs([(x.data_ptr(), x.stride(0), w, off) for x, w, off in zip(xs, widths, offs)])
    out = []
    for tma in (1, 0):
        L.LIB.set_gemm_tma(bool(tma))
        C = torch.zeros(M, N, device="cuda")
        L.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, bias.data_ptr(), C.data_ptr(), N, M)
        ref = sum(x[:, :w].double() @ W[:, off:off + w].double().t() for x, w, off in zip(xs, widths, offs)) + bias.double()
        e_f = rel(C, ref)
        dxs = [torch.full_like(x, 0.5) for x in xs]
        dsp, _ = L.segs([(d.data_ptr(), d.stride(0), w, off) for d, w, off in zip(dxs, widths, offs)])
        L.call("nasrec_seg_linear_dgrad", dC.data_ptr(), ldc, N, W.data_ptr(), Ktot, 0, dsp, ns, M, 1)
        e_d = max(rel(d[:, :w], dC.double() @ W[:, off:off + w].double() + 0.5) for d, w, off in zip(dxs, widths, offs))
        dW = torch.zeros_like(W)
        L.call("nasrec_seg_linear_wgrad", dC.data_ptr(), ldc, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)
        e_w = max(rel(dW[:, off:off + w], dC.double().t() @ x[:, :w].double()) for x, w, off in zip(xs, widths, offs))
        out.append((e_f, e_d, e_w))
    nbad = int(((C.double() - ref).abs() > 1e-5 * ref.abs().max()).sum())
    flag = ("  <-- BAD (fwd elements off: %d)" % nbad) if max(max(o) for o in out) > 5e-6 else ""
    print("bn=%s ns=%s M=%d N=%d widths=%s  tma fwd/dgrad/wgrad %.1e %.1e %.1e | ldg %.1e %.1e %.1e%s" % (
        os.environ.get("NASREC_TC_BN", "-"), os.environ.get("NASREC_TC_NS", "-"), M, N, widths, *out[0], *out[1], flag), flush=True)
