"""GPU: accuracy and timing of the segment-GEMM entry points in every arithmetic mode,
against an fp64 reference.  Prints one line per (case, mode)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)

def segs_of(tensors, widths, offs):
    return _lib.segs([(t.data_ptr(), t.stride(0), w, o) for t, w, o in zip(tensors, widths, offs)])

def err(a, ref):
    a = a.double().cpu(); ref = ref.cpu()
    return float((a - ref).abs().max() / ref.abs().max())

def timeit(fn, n=200):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def case_2d(M, N, widths, ldw_extra=0, n_off=0, bias=True):
    K = sum(widths)
    offs, o = [], 3
    for w in widths:
        offs.append(o); o += w + 5
    Ktot = o + ldw_extra
    xs = [torch.randn(M, w + 3, generator=g).to(dev) for w in widths]      # padded row stride
    W = (torch.randn(N + n_off, Ktot, generator=g) / np.sqrt(K)).to(dev)
    b = torch.randn(N + n_off, generator=g).to(dev) if bias else None
    ref = sum(x[:, :w].double() @ W[n_off:, o:o + w].double().t() for x, w, o in zip(xs, widths, offs))
    if bias: ref = ref + b[n_off:].double()
    dC = torch.randn(M, N, generator=g).to(dev)
    dref = [dC.double() @ W[n_off:, o:o + w].double() for w, o in zip(widths, offs)]
    wref = [dC.double().t() @ x[:, :w].double() for x, w in zip(xs, widths)]
    sp, ns = segs_of(xs, widths, offs)
    for mode in (0, 1, 3, 4):
        _lib.LIB.set_gemm_mode(mode)
        C = torch.zeros(M, N, device=dev)
        f = lambda: _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, n_off, N, b.data_ptr() if bias else None, C.data_ptr(), N, M)
        f(); torch.cuda.synchronize()
        e_f = err(C, ref)
        dxs = [torch.zeros_like(x) for x in xs]
        dsp, _ = segs_of(dxs, widths, offs)
        _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, n_off, dsp, ns, M, 0)
        torch.cuda.synchronize()
        e_d = max(err(dx[:, :w], r) for dx, w, r in zip(dxs, widths, dref))
        dW = torch.zeros_like(W)
        _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, n_off, M, 0)
        torch.cuda.synchronize()
        e_w = max(err(dW[n_off:, o:o + w], r) for w, o, r in zip(widths, offs, wref))
        untouched = float(dW[:n_off].abs().max()) if n_off else 0.0
        us = timeit(f)
        print("2d M=%d N=%d K=%s mode=%d  fwd %.2e dgrad %.2e wgrad %.2e  (rows<n_off untouched: %g)  fwd %.1f us  %.1f TFLOP/s"
              % (M, N, widths, mode, e_f, e_d, e_w, untouched, us, 2.0 * M * N * K / us / 1e6), flush=True)

def case_3d(B, P, rows, bias=True):
    S = sum(rows)
    offs, o = [], 2
    for r in rows:
        offs.append(o); o += r + 1
    Stot = o
    xs = [torch.randn(B, r + 2, 16, generator=g).to(dev) for r in rows]
    W = (torch.randn(P, Stot, generator=g) / np.sqrt(S)).to(dev)
    b = torch.randn(P, generator=g).to(dev) if bias else None
    ref = sum(torch.einsum("ps,bse->bpe", W[:, o:o + r].double(), x[:, :r].double()) for x, r, o in zip(xs, rows, offs))
    if bias: ref = ref + b.double()[None, :, None]
    dZ = torch.randn(B, P, 16, generator=g).to(dev)
    dref = [torch.einsum("ps,bpe->bse", W[:, o:o + r].double(), dZ.double()) for r, o in zip(rows, offs)]
    wref = [torch.einsum("bpe,bse->ps", dZ.double(), x[:, :r].double()) for x, r in zip(xs, rows)]
    sp, ns = _lib.segs([(x.data_ptr(), x.stride(0), r, o) for x, r, o in zip(xs, rows, offs)])
    for mode in (0, 1, 3, 4):
        _lib.LIB.set_gemm_mode(mode)
        Z = torch.zeros(B, P, 16, device=dev)
        f = lambda: _lib.call("nasrec_sproj_fwd", sp, ns, W.data_ptr(), Stot, P, b.data_ptr() if bias else None, Z.data_ptr(), P * 16, B)
        f(); torch.cuda.synchronize()
        e_f = err(Z, ref)
        dxs = [torch.zeros_like(x) for x in xs]
        dsp, _ = _lib.segs([(x.data_ptr(), x.stride(0), r, o) for x, r, o in zip(dxs, rows, offs)])
        _lib.call("nasrec_sproj_dgrad", dZ.data_ptr(), P * 16, P, W.data_ptr(), Stot, dsp, ns, B, 0)
        torch.cuda.synchronize()
        e_d = max(err(dx[:, :r], rf) for dx, r, rf in zip(dxs, rows, dref))
        dW = torch.zeros_like(W)
        ws = torch.empty(_lib.query("nasrec_sproj_wgrad_ws_floats", P, S, B), device=dev)
        fw = lambda: _lib.call("nasrec_sproj_wgrad", dZ.data_ptr(), P * 16, P, sp, ns, dW.data_ptr(), Stot, B, 0, ws.data_ptr())
        fw(); torch.cuda.synchronize()
        e_w = max(err(dW[:, o:o + r], rf) for r, o, rf in zip(rows, offs, wref))
        us, usw = timeit(f), timeit(fw)
        print("3d B=%d P=%d rows=%s mode=%d  fwd %.2e dgrad %.2e wgrad %.2e  fwd %.1f us wgrad %.1f us"
              % (B, P, rows, mode, e_f, e_d, e_w, us, usw), flush=True)

def case_model(M, N, nd, nsrc, accumulate=False):
    """Operands laid out exactly as in the supernet: contiguous sources, reference weight layout."""
    widths = [nd] + [1024] * (nsrc - 1)
    offs = [0] + [nd + 1024 * j for j in range(nsrc - 1)]
    Ktot = nd + 1024 * (nsrc - 1)
    xs = [torch.randn(M, w, generator=g).to(dev) for w in widths]
    W = (torch.randn(N, Ktot, generator=g) / np.sqrt(Ktot)).to(dev)
    ref = sum(x.double() @ W[:, o:o + w].double().t() for x, w, o in zip(xs, widths, offs))
    dC = torch.randn(M, N, generator=g).to(dev)
    dref = [dC.double() @ W[:, o:o + w].double() for w, o in zip(widths, offs)]
    wref = [dC.double().t() @ x.double() for x in xs]
    sp, ns = _lib.segs([(x.data_ptr(), w, w, o) for x, w, o in zip(xs, widths, offs)])
    for mode in (0, 3):
        _lib.LIB.set_gemm_mode(mode)
        C = torch.zeros(M, N, device=dev)
        _lib.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, None, C.data_ptr(), N, M)
        base = [torch.randn_like(x) if accumulate else torch.zeros_like(x) for x in xs]
        dxs = [b.clone() for b in base]
        dsp, _ = _lib.segs([(x.data_ptr(), w, w, o) for x, w, o in zip(dxs, widths, offs)])
        _lib.call("nasrec_seg_linear_dgrad", dC.data_ptr(), N, N, W.data_ptr(), Ktot, 0, dsp, ns, M, int(accumulate))
        dW = torch.zeros_like(W)
        _lib.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)
        torch.cuda.synchronize()
        e_f = err(C, ref)
        e_d = max(err(dx - (b if accumulate else 0), r) for dx, b, r in zip(dxs, base, dref))
        e_w = max(err(dW[:, o:o + w], r) for w, o, r in zip(widths, offs, wref))
        print("model-layout M=%d N=%d nsrc=%d acc=%d mode=%d  fwd %.2e dgrad %.2e wgrad %.2e" % (M, N, nsrc, accumulate, mode, e_f, e_d, e_w), flush=True)


if __name__ == "__main__":
    a = torch.randn(4096, 4096, device=dev)
    t0 = time.time()
    while time.time() - t0 < 1.5:          # ramp the clocks before timing anything
        (a @ a).sum().item()
    if len(sys.argv) > 1 and sys.argv[1] == "model":
        for M in (5, 512):
            for N in (1024, 16, 128):
                for nsrc in (1, 2, 7):
                    case_model(M, N, 13, nsrc)
        case_model(5, 1024, 13, 7, accumulate=True)
        _lib.LIB.set_gemm_mode(3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        case_2d(5, 1024, [13])
        case_2d(5, 16, [13, 1024, 1024])
        case_2d(5, 1024, [13, 1024])
        case_2d(6, 16, [13])
        case_2d(64, 13, [16])
        case_3d(5, 64, [26])
        case_3d(5, 45, [26, 64, 8])
        _lib.LIB.set_gemm_mode(3)
        sys.exit(0)
    case_2d(128, 64, [32])
    case_2d(512, 1024, [13, 1000])
    case_2d(37, 16, [300], n_off=5)
    case_2d(512, 1024, [1035], bias=False)
    case_2d(8192, 1024, [13, 256, 768])
    case_2d(257, 1, [100, 28])
    case_3d(64, 64, [26])
    case_3d(512, 64, [26, 32, 8, 64])
    case_3d(33, 45, [98])
    _lib.LIB.set_gemm_mode(0)
