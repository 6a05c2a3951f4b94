"""GPU: the TMA-fed tensor-core GEMM (csrc/gemm_tma.cuh) against the LDG-producer kernel (csrc/gemm_tc.cuh) and fp64.

Both kernels implement the same arithmetic (rn split into tf32 hi/lo, three products per k-step, accumulators
rotated over k-tiles, fixed-order split-K), so on identical operands they must agree BIT FOR BIT: any difference is
a layout bug (tensor map, swizzle, shared-memory descriptor, MN-major transposition).  The fp64 check holds both
to the 3xTF32 error level.  Replaces: nn.LazyLinear on the padded concat, nasrec/supernet/modules.py:171,223,340,359.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from nasrec_b200 import _lib
    return _lib


def _planes(W, first=0):
    L = _lib()
    rows, cols = W.shape
    ldp = (cols + ((4 - first % 4) % 4) + 3) & ~3
    hi = torch.zeros(rows, ldp, device=W.device)
    lo = torch.zeros(rows, ldp, device=W.device)
    L.call("nasrec_planes_refresh", W.data_ptr(), cols, rows, cols, first, hi.data_ptr(), lo.data_ptr(), ldp)
    return hi, lo, ldp, first


def _announce(W, pl):
    hi, lo, ldp, first = pl
    _lib().LIB.set_weight_planes(W.data_ptr(), hi.data_ptr(), lo.data_ptr(), ldp, W.shape[0], W.shape[1], first)


def _tma_launches():
    return _lib().query("nasrec_tensor_map_stats", 2)


def _both(fn):
    """Run fn() with the TMA path off, then on; returns (ldg_result, tma_result) and asserts TMA was really taken."""
    L = _lib()
    small = L.LIB.set_small_k(0)          # this file is about the two tensor-core kernels: no CUDA-core shortcut for short K
    try:
        L.LIB.set_gemm_tma(False)
        a = fn()
        L.LIB.set_gemm_tma(True)
        before = _tma_launches()
        b = fn()
        torch.cuda.synchronize()
        assert _tma_launches() > before, "the launch did not qualify for the TMA path"
    finally:
        L.LIB.set_small_k(small)
    return a, b


def _rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max())


def test_planes_split_is_exact():
    W = torch.randn(37, 1037, device="cuda")
    hi, lo, ldp, _ = _planes(W, first=13)
    assert ldp == 1040
    assert torch.equal(hi[:, :13] + lo[:, :13], W[:, :13]) and torch.equal(hi[:, 16:] + lo[:, 16:], W[:, 13:])
    assert float(hi[:, 13:16].abs().max()) == 0.0 and float(lo[:, 13:16].abs().max()) == 0.0
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0           # hi is a tf32 number
    assert float(lo.abs().max()) <= float(W.abs().max()) * 2.0 ** -11


@pytest.mark.parametrize("M,N,n_off,widths", [
    (512, 1024, 0, [13, 256, 1024]),
    (3, 1024, 0, [13, 64]),
    (1000, 1024, 0, [16, 768, 1024, 512]),
    (257, 64, 16, [13, 128]),               # SigmoidGating self-linear slice (rows n_off.., modules.py:578)
    (512, 16, 0, [13, 1024, 32]),           # DotProduct dense projection, N = 16
    (8192, 128, 0, [512]),                  # dense -> sparse merger at the evaluation batch
    (130, 1, 0, [1024, 1024, 128]),         # _final
])
def test_segment_linear_tma_equals_ldg(M, N, n_off, widths):
    L = _lib()
    g = torch.Generator().manual_seed(1)
    offs, o = [], 0
    for j, w in enumerate(widths):
        offs.append(o)
        o += 13 if j == 0 and w == 13 else 1024
    Ktot = o
    xs = [torch.randn(M, (w + 3) & ~3, generator=g).cuda() for w in widths]
    W = (torch.randn(N + n_off, Ktot, generator=g) / np.sqrt(sum(widths))).cuda()
    bias = torch.randn(N + n_off, generator=g).cuda()
    ldc = (N + 3) & ~3                                   # dY rows 16-byte aligned, as the executor's arena allocates them
    dC = torch.randn(M, ldc, generator=g).cuda()[:, :N]
    pl = _planes(W, first=13 if widths[0] == 13 else 0)
    sp, ns = L.segs([(x.data_ptr(), x.stride(0), w, off) for x, w, off in zip(xs, widths, offs)])

    def fwd():
        _announce(W, pl)
        C = torch.zeros(M, N, device="cuda")
        L.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, n_off, N, bias.data_ptr(), C.data_ptr(), N, M)
        return C

    def dgrad():
        _announce(W, pl)
        dxs = [torch.full_like(x, 0.5) for x in xs]
        dsp, _ = L.segs([(d.data_ptr(), d.stride(0), w, off) for d, w, off in zip(dxs, widths, offs)])
        L.call("nasrec_seg_linear_dgrad", dC.data_ptr(), ldc, N, W.data_ptr(), Ktot, n_off, dsp, ns, M, 1)
        return dxs

    def wgrad():
        dW = torch.zeros_like(W)
        L.call("nasrec_seg_linear_wgrad", dC.data_ptr(), ldc, N, sp, ns, dW.data_ptr(), Ktot, n_off, M, 0)
        return dW

    a, b = _both(fwd)
    assert torch.equal(a, b)
    ref = sum(x[:, :w].double() @ W[n_off:, off:off + w].double().t() for x, w, off in zip(xs, widths, offs)) + bias[n_off:].double()
    assert _rel(b, ref) < 5e-6
    a, b = _both(dgrad)
    for da, db, w, off in zip(a, b, widths, offs):
        assert torch.equal(da, db)
        ref = dC.double() @ W[n_off:, off:off + w].double() + 0.5
        assert _rel(db[:, :w], ref) < 5e-6
    a, b = _both(wgrad)
    assert torch.equal(a, b)
    for x, w, off in zip(xs, widths, offs):
        ref = dC.double().t() @ x[:, :w].double()
        assert _rel(b[n_off:, off:off + w], ref) < 5e-6
    if n_off:
        assert float(b[:n_off].abs().max()) == 0.0


@pytest.mark.parametrize("B,P,rows", [
    (8, 64, [26]),
    (257, 45, [26, 64, 8]),                 # DotProduct sparse projection over stem + one block (+ merger rows)
    (512, 64, [26, 48, 8, 64]),
    (3, 16, [10]),
])
def test_sparse_projection_tma_equals_ldg(B, P, rows):
    L = _lib()
    g = torch.Generator().manual_seed(2)
    E = 16
    offs, o = [], 0
    for j, r in enumerate(rows):
        offs.append(o)
        o += r if j == 0 else 72
    Stot = o
    xs = [torch.randn(B, r + 2, E, generator=g).cuda() for r in rows]        # batch stride wider than the live rows
    W = (torch.randn(P, Stot, generator=g) / np.sqrt(sum(rows))).cuda()
    bias = torch.randn(P, generator=g).cuda()
    dZ = torch.randn(B, P, E, generator=g).cuda()
    pl = _planes(W, first=rows[0])
    sp, ns = L.segs([(x.data_ptr(), x.stride(0), r, off) for x, r, off in zip(xs, rows, offs)])

    def fwd():
        _announce(W, pl)
        Z = torch.zeros(B, P, E, device="cuda")
        L.call("nasrec_sproj_fwd", sp, ns, W.data_ptr(), Stot, P, bias.data_ptr(), Z.data_ptr(), P * E, B)
        return Z

    def dgrad():
        _announce(W, pl)
        dxs = [torch.zeros_like(x) for x in xs]
        dsp, _ = L.segs([(d.data_ptr(), d.stride(0), r, off) for d, r, off in zip(dxs, rows, offs)])
        L.call("nasrec_sproj_dgrad", dZ.data_ptr(), P * E, P, W.data_ptr(), Stot, dsp, ns, B, 0)
        return dxs

    def wgrad():
        dW = torch.zeros_like(W)
        ws = torch.empty(L.query("nasrec_sproj_wgrad_ws_floats", P, sum(rows), B), device="cuda")
        L.call("nasrec_sproj_wgrad", dZ.data_ptr(), P * E, P, sp, ns, dW.data_ptr(), Stot, B, 0, ws.data_ptr())
        return dW

    a, b = _both(fwd)
    assert torch.equal(a, b)
    ref = sum(torch.einsum("pr,bre->bpe", W[:, off:off + r].double(), x[:, :r].double()) for x, r, off in zip(xs, rows, offs))
    assert _rel(b, ref + bias.double()[None, :, None]) < 5e-6
    a, b = _both(dgrad)
    for da, db, r, off in zip(a, b, rows, offs):
        assert torch.equal(da, db)
        assert _rel(db[:, :r], torch.einsum("bpe,pr->bre", dZ.double(), W[:, off:off + r].double())) < 5e-6
    a, b = _both(wgrad)
    assert torch.equal(a, b)
    for x, r, off in zip(xs, rows, offs):
        assert _rel(b[:, off:off + r], torch.einsum("bpe,bre->pr", dZ.double(), x[:, :r].double())) < 5e-6


def test_adagrad_keeps_planes_in_step():
    """nasrec_adagrad_multi_planes == nasrec_adagrad_multi on the weights, and the planes it writes are the exact
    split of the updated weights (train_utils.py:286 torch.optim.Adagrad.step)."""
    L = _lib()
    shapes = [(64, 1037), (16, 128), (1, 2176)]
    ws = [torch.randn(*s, device="cuda") for s in shapes]
    gs = [torch.randn(*s, device="cuda") for s in shapes]
    ss = [torch.rand(*s, device="cuda") for s in shapes]
    w2, s2 = [w.clone() for w in ws], [s.clone() for s in ss]
    firsts = [13, 0, 0]
    pls = [_planes(w, f) for w, f in zip(ws, firsts)]
    sizes = L.i64_array([w.numel() for w in ws])
    L.call("nasrec_adagrad_multi", L.ptr_array([w.data_ptr() for w in w2]), L.ptr_array([g.data_ptr() for g in gs]),
           L.ptr_array([s.data_ptr() for s in s2]), sizes, 3, 0.12, 1e-2, None)
    L.call("nasrec_adagrad_multi_planes", L.ptr_array([w.data_ptr() for w in ws]), L.ptr_array([g.data_ptr() for g in gs]),
           L.ptr_array([s.data_ptr() for s in ss]), sizes, 3, 0.12, 1e-2, None,
           L.ptr_array([p[0].data_ptr() for p in pls]), L.ptr_array([p[1].data_ptr() for p in pls]),
           L.i32_array([w.shape[1] for w in ws]), L.i32_array(firsts), L.i64_array([p[2] for p in pls]))
    for w, wr, s, sr, pl, f in zip(ws, w2, ss, s2, pls, firsts):
        assert torch.equal(w, wr) and torch.equal(s, sr)
        hi, lo, _, _ = _planes(w, f)
        assert torch.equal(pl[0], hi) and torch.equal(pl[1], lo)


@pytest.mark.parametrize("M,N,widths", [(512, 1024, [16]), (512, 1024, [13]), (512, 16, [13]), (300, 64, [13, 48])])
def test_small_k_cuda_core_path(M, N, widths):
    """K <= 64 goes to the CUDA-core kernel (nasrec_set_small_k): same contract, against fp64, and the tensor-core path
    agrees to the 3xTF32 error level.  Replaces the narrow linears of modules.py:171 (dense stem), :340 / :385 (16-wide
    DotProduct / FM projections)."""
    L = _lib()
    g = torch.Generator().manual_seed(5)
    offs, o = [], 0
    for w in widths:
        offs.append(o)
        o += w
    Ktot = o
    xs = [torch.randn(M, (w + 3) & ~3, generator=g).cuda() for w in widths]
    W = (torch.randn(N, Ktot, generator=g) / np.sqrt(Ktot)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    dC = torch.randn(M, N, generator=g).cuda()
    sp, ns = L.segs([(x.data_ptr(), x.stride(0), w, off) for x, w, off in zip(xs, widths, offs)])
    ref = sum(x[:, :w].double() @ W[:, off:off + w].double().t() for x, w, off in zip(xs, widths, offs)) + bias.double()
    outs = []
    for k in (64, 0):
        old = L.LIB.set_small_k(k)
        try:
            before = _tma_launches()
            C = torch.zeros(M, N, device="cuda")
            L.call("nasrec_seg_linear_fwd", sp, ns, W.data_ptr(), Ktot, 0, N, bias.data_ptr(), C.data_ptr(), N, M)
            torch.cuda.synchronize()
            if k:
                assert _tma_launches() == before
        finally:
            L.LIB.set_small_k(old)
        assert _rel(C, ref) < 5e-6
        outs.append(C)
    assert _rel(outs[0], outs[1].double()) < 5e-6


@pytest.mark.parametrize("B,P,rows", [(512, 64, [26]), (257, 45, [26, 30]), (64, 16, [10, 8, 40])])
def test_small_k_sparse_projection(B, P, rows):
    """Sparse-axis projections with short contractions (K = rows <= 64 forward, K = P <= 64 backward) on the one-shot
    CUDA-core kernel, several gradient targets in one launch (ElasticLinear3D / DotProduct, modules.py:223-262, 340-359)."""
    L = _lib()
    g = torch.Generator().manual_seed(7)
    E = 16
    offs, o = [], 0
    for r in rows:
        offs.append(o)
        o += r
    Stot = o
    xs = [torch.randn(B, r + 2, E, generator=g).cuda() for r in rows]
    W = (torch.randn(P, Stot, generator=g) / np.sqrt(Stot)).cuda()
    bias = torch.randn(P, generator=g).cuda()
    dZ = torch.randn(B, P, E, generator=g).cuda()
    sp, ns = L.segs([(x.data_ptr(), x.stride(0), r, off) for x, r, off in zip(xs, rows, offs)])
    old = L.LIB.set_small_k(64)
    try:
        before = _tma_launches()
        Z = torch.zeros(B, P, E, device="cuda")
        L.call("nasrec_sproj_fwd", sp, ns, W.data_ptr(), Stot, P, bias.data_ptr(), Z.data_ptr(), P * E, B)
        dxs = [torch.full_like(x, 0.25) for x in xs]
        dsp, _ = L.segs([(d.data_ptr(), d.stride(0), r, off) for d, r, off in zip(dxs, rows, offs)])
        L.call("nasrec_sproj_dgrad", dZ.data_ptr(), P * E, P, W.data_ptr(), Stot, dsp, ns, B, 1)
        torch.cuda.synchronize()
        if Stot <= 64:
            assert _tma_launches() == before
    finally:
        L.LIB.set_small_k(old)
    ref = sum(torch.einsum("pr,bre->bpe", W[:, off:off + r].double(), x[:, :r].double()) for x, r, off in zip(xs, rows, offs))
    assert _rel(Z, ref + bias.double()[None, :, None]) < 5e-6
    for d, r, off in zip(dxs, rows, offs):
        assert _rel(d[:, :r], torch.einsum("bpe,pr->bre", dZ.double(), W[:, off:off + r].double()) + 0.25) < 5e-6
        assert float((d[:, r:] - 0.25).abs().max()) == 0.0


@pytest.mark.parametrize("shapes", [
    [(512, 1024, [13, 1024]), (512, 64, [416]), (512, 16, [1024, 1024, 128])],      # a block's worth: split-K clusters + flat grid
    [(512, 1024, [1024] * 3), (512, 1024, [13]), (512, 128, [1024]), (512, 1024, [1035]), (512, 256, [13, 416, 1024])] * 3,
])
def test_deferred_weight_gradients_batched_launch(shapes):
    """nasrec_wgrad_defer / nasrec_wgrad_flush: queued weight gradients of several linears run as one grid over all their
    output tiles (flat tile index, one plan for the batch); every dW equals the fp64 product and the one-by-one launches
    to the 3xTF32 error level.  Replaces the per-parameter weight.grad of loss.backward() (train_utils.py:283)."""
    L = _lib()
    g = torch.Generator().manual_seed(11)
    cases = []
    for M, N, widths in shapes:
        xs = [torch.randn(M, (w + 3) & ~3, generator=g).cuda() for w in widths]
        offs, o = [], 0
        for w in widths:
            offs.append(o)
            o += w
        dC = torch.randn(M, N, generator=g).cuda()
        cases.append((M, N, widths, xs, offs, o, dC))

    def run(defer):
        outs = []
        old = L.LIB.load().cdll.nasrec_wgrad_defer(1 if defer else 0)
        try:
            for M, N, widths, xs, offs, Ktot, dC in cases:
                dW = torch.zeros(N, Ktot, device="cuda")
                sp, ns = L.segs([(x.data_ptr(), x.stride(0), w, off) for x, w, off in zip(xs, widths, offs)])
                L.call("nasrec_seg_linear_wgrad", dC.data_ptr(), N, N, sp, ns, dW.data_ptr(), Ktot, 0, M, 0)
                outs.append(dW)
            if defer:
                assert L.query("nasrec_wgrad_pending") == len(cases)
                assert float(outs[0].abs().max()) == 0.0                  # nothing ran yet
                L.call("nasrec_wgrad_flush")
                assert L.query("nasrec_wgrad_pending") == 0
        finally:
            L.LIB.cdll.nasrec_wgrad_defer(old)
        torch.cuda.synchronize()
        return outs

    one_by_one, batched = run(False), run(True)
    for (M, N, widths, xs, offs, Ktot, dC), a, b in zip(cases, one_by_one, batched):
        for x, w, off in zip(xs, widths, offs):
            ref = dC.double().t() @ x[:, :w].double()
            assert _rel(b[:, off:off + w], ref) < 5e-6
            assert _rel(a[:, off:off + w], ref) < 5e-6
