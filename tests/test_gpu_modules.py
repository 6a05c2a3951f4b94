"""GPU: each building block of the CUDA path against the oracle (the reference's
module math restated on CPU), forward and backward, through the public module
API (which calls the C ABI).  fp32 tolerance: 2e-5 relative on outputs, 2e-4 on
gradients (different summation order than MKL; LayerNorm amplifies)."""
import numpy as np
import pytest
import torch

from oracle import nasrec_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-5
BWD_TOL = 2e-4


def _randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() >= 2:
                a = (6.0 / (p.shape[0] + p.shape[-1])) ** 0.5
                p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * a).to(p.device))
            elif "norm" in n or "_ln" in n:
                if n.endswith("weight"):
                    p.copy_((1 + 0.1 * torch.randn(p.shape, generator=g)).to(p.device))
                else:
                    p.copy_((0.1 * torch.randn(p.shape, generator=g)).to(p.device))
            else:
                p.copy_((0.1 * torch.randn(p.shape, generator=g)).to(p.device))


def _check(mod, inputs, my_call, oracle_call, seed=0):
    """inputs: CPU float tensors. Runs both sides with a random linear loss."""
    dev = torch.device("cuda")
    xs_gpu = [x.clone().to(dev).requires_grad_(True) for x in inputs]
    out = my_call(mod, *xs_gpu)            # first call materialises the lazy layers
    _randomize(mod, seed)
    for p in mod.parameters():
        p.grad = None
    xs_gpu = [x.clone().to(dev).requires_grad_(True) for x in inputs]
    out = my_call(mod, *xs_gpu)
    R = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1))
    (out * R.to(dev)).sum().backward()
    torch.cuda.synchronize()

    sd = {"p." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xs_cpu = [x.clone().requires_grad_(True) for x in inputs]
    ref = oracle_call(sd, *xs_cpu)
    (ref * R).sum().backward()

    assert out.shape == ref.shape
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < FWD_TOL
    for a, b in zip(xs_gpu, xs_cpu):
        if b.grad is None:
            continue
        assert a.grad is not None
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < BWD_TOL
    named = dict(mod.named_parameters())
    for k, v in sd.items():
        name = k[2:]
        if name not in named:
            continue
        if v.grad is None or float(v.grad.abs().max()) == 0.0:
            g = named[name].grad
            assert g is None or float(g.abs().max()) == 0.0, name
            continue
        assert named[name].grad is not None, name
        assert rel_err(named[name].grad.cpu().numpy(), v.grad.numpy()) < BWD_TOL, name


@pytest.mark.parametrize("ln,fixed,K,maxd,d", [(True, False, 300, 1024, 64), (True, False, 1037, 1024, 1024),
                                               (False, True, 813, 128, 128), (True, True, 45, 96, 96),
                                               (False, False, 77, 256, 32)])
def test_elastic_linear(ln, fixed, K, maxd, d):
    from nasrec_b200.supernet.modules import ElasticLinear
    mod = ElasticLinear(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd, activation="relu").cuda()
    x = torch.randn(37, K)
    _check(mod, [x], lambda m, a: m(a, d), lambda sd, a: orc.fc(sd, "p", a, d, maxd, ln, fixed))


@pytest.mark.parametrize("ln,fixed,S,maxs,s", [(True, False, 98, 64, 32), (True, False, 458, 64, 64),
                                               (False, True, 194, 48, 48), (True, False, 26, 64, 16)])
def test_elastic_linear_3d(ln, fixed, S, maxs, s):
    from nasrec_b200.supernet.modules import ElasticLinear3D
    mod = ElasticLinear3D(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxs, activation="relu",
                          embedding_dim=16).cuda()
    x = torch.randn(21, S, 16)
    _check(mod, [x], lambda m, a: m(a, s), lambda sd, a: orc.efc(sd, "p", a, s, maxs, ln, fixed))


@pytest.mark.parametrize("ln,fixed,Kd,S,maxd,d", [(True, False, 1037, 98, 1024, 256), (False, True, 160, 162, 768, 768),
                                                  (False, True, 13, 26, 32, 32), (True, False, 13, 26, 1024, 16)])
def test_dot_product(ln, fixed, Kd, S, maxd, d):
    from nasrec_b200.supernet.modules import DotProduct
    mod = DotProduct(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd, embedding_dim=16).cuda()
    x, sp = torch.randn(19, Kd), torch.randn(19, S, 16) * 0.5
    _check(mod, [x, sp], lambda m, a, b: m(a, b, d),
           lambda sd, a, b: orc.dot_product(sd, "p", a, b, d, maxd, ln, fixed))


@pytest.mark.parametrize("ln,fixed,Kl,Kr,maxd,d", [(True, False, 1037, 1037, 1024, 128), (False, True, 32, 13, 768, 768),
                                                   (True, True, 13, 40, 64, 64), (False, True, 64, 64, 64, 64)])
def test_sum(ln, fixed, Kl, Kr, maxd, d):
    from nasrec_b200.supernet.modules import Sum
    mod = Sum(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd, activation="relu").cuda()
    l, r = torch.randn(23, Kl), torch.randn(23, Kr)
    _check(mod, [l, r], lambda m, a, b: m(a, b, d), lambda sd, a, b: orc.sum_node(sd, "p", a, b, d, maxd, ln, fixed))


@pytest.mark.parametrize("ln,fixed,Kl,Kr,maxd,d", [(True, False, 269, 269, 1024, 512), (False, True, 32, 13, 768, 768),
                                                   (False, True, 13, 13, 128, 128), (True, True, 20, 48, 48, 48)])
def test_sigmoid_gating(ln, fixed, Kl, Kr, maxd, d):
    from nasrec_b200.supernet.modules import SigmoidGating
    mod = SigmoidGating(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd, activation="relu").cuda()
    l, r = torch.randn(23, Kl), torch.randn(23, Kr)
    _check(mod, [l, r], lambda m, a, b: m(a, b, d),
           lambda sd, a, b: orc.sigmoid_gating(sd, "p", a, b, d, maxd, ln, fixed))


@pytest.mark.parametrize("ln,fixed,S,maxs,s", [(True, False, 98, 64, 48), (True, False, 26, 64, 64),
                                               (False, True, 128, 16, 16), (False, True, 104, 48, 48),
                                               (True, False, 170, 64, 16)])
def test_transformer(ln, fixed, S, maxs, s):
    from nasrec_b200.supernet.modules import Transformer
    mod = Transformer(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxs, activation="relu",
                      embedding_dim=16).cuda()
    x = torch.randn(17, S, 16)
    _check(mod, [x], lambda m, a: m(a, s), lambda sd, a: orc.transformer(sd, "p", a, s, maxs, ln, fixed))


@pytest.mark.parametrize("ln,fixed,S,maxd,d", [(True, False, 64, 1024, 256), (False, True, 32, 768, 768),
                                               (False, True, 16, 16, 16)])
def test_factorization_machine(ln, fixed, S, maxd, d):
    from nasrec_b200.supernet.modules import FactorizationMachine3D
    mod = FactorizationMachine3D(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd).cuda()
    x = torch.randn(29, S, 16) * 0.3
    _check(mod, [x], lambda m, a: m(a, d), lambda sd, a: orc.fm3d(sd, "p", a, d, maxd, ln, fixed))


def test_embedding_gather_sort_reduce_adagrad_bit_exact():
    """Integer/byte side: bit-exact gathered rows, row sets and (for duplicate-free
    rows) gradients; duplicates summed in ascending sample order == oracle order."""
    from nasrec_b200 import _lib
    dev = torch.device("cuda")
    rs = np.random.RandomState(0)
    sizes = [7, 3, 50, 1000, 4, 123457]
    F = len(sizes)
    for B in (1, 33, 512, 1000, 4096):
        tables = [torch.from_numpy(rs.randn(n, 16).astype(np.float32)).to(dev) for n in sizes]
        cat = np.stack([rs.randint(0, n, B) for n in sizes], 1).astype(np.int64)
        cat_d = torch.from_numpy(cat).to(dev)
        ptrs = torch.tensor([t.data_ptr() for t in tables], dtype=torch.int64, device=dev)
        rows = torch.tensor(sizes, dtype=torch.int64, device=dev)
        out = torch.empty(B, F, 16, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("nasrec_emb_gather_fwd", ptrs.data_ptr(), rows.data_ptr(), cat_d.data_ptr(), out.data_ptr(), B, F,
                  err.data_ptr())
        want = orc.embedding_gather([t.cpu().numpy() for t in tables], cat)
        assert np.array_equal(out.cpu().numpy(), want)
        assert int(err.item()) == 0
        g = rs.randn(B, F, 16).astype(np.float32)
        g_d = torch.from_numpy(g).to(dev)
        uniq = torch.full((F, B), -1, dtype=torch.int64, device=dev)
        nuniq = torch.zeros(F, dtype=torch.int32, device=dev)
        rg = torch.zeros(F, B, 16, device=dev)
        sumsq = torch.zeros(F, device=dev)
        scratch = torch.zeros(F, B + 1, dtype=torch.int32, device=dev)
        _lib.call("nasrec_emb_grad_sort_reduce", cat_d.data_ptr(), g_d.data_ptr(), B, F, uniq.data_ptr(),
                  nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(), scratch.data_ptr())
        torch.cuda.synchronize()
        ref = orc.embedding_grad_rows(cat, g)
        for f, (rows_ref, acc_ref) in enumerate(ref):
            U = int(nuniq[f].item())
            assert U == len(rows_ref)
            assert np.array_equal(uniq[f, :U].cpu().numpy(), rows_ref)           # sorted unique row set, bit-exact
            assert np.array_equal(rg[f, :U].cpu().numpy(), acc_ref)              # same summation order -> bit-exact
            assert abs(float(sumsq[f].item()) - float((acc_ref.astype(np.float64) ** 2).sum())) < 1e-3 * max(
                1.0, float((acc_ref ** 2).sum()))
        # row-wise Adagrad == dense Adagrad on the dense gradient
        states = [torch.zeros_like(t) for t in tables]
        sp = torch.tensor([s.data_ptr() for s in states], dtype=torch.int64, device=dev)
        coef = torch.tensor([0.5], device=dev)
        before = [t.clone() for t in tables]
        _lib.call("nasrec_emb_rowwise_adagrad", uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), ptrs.data_ptr(),
                  sp.data_ptr(), B, F, 0.12, 1e-2, coef.data_ptr())
        for f in range(F):
            dense = np.zeros((sizes[f], 16), np.float32)
            dense[ref[f][0]] = ref[f][1] * np.float32(0.5)
            st = dense * dense
            want = before[f].cpu().numpy() - np.float32(0.12) * (dense / (np.sqrt(st) + np.float32(1e-2)))
            assert np.allclose(tables[f].cpu().numpy(), want, rtol=1e-6, atol=1e-7)
            assert np.allclose(states[f].cpu().numpy(), st, rtol=1e-6, atol=1e-9)


def test_out_of_range_id_sets_error_flag():
    from nasrec_b200 import _lib
    dev = torch.device("cuda")
    t = torch.randn(5, 16, device=dev)
    ptrs = torch.tensor([t.data_ptr()], dtype=torch.int64, device=dev)
    rows = torch.tensor([5], dtype=torch.int64, device=dev)
    cat = torch.tensor([[1], [7]], dtype=torch.int64, device=dev)
    out = torch.empty(2, 1, 16, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("nasrec_emb_gather_fwd", ptrs.data_ptr(), rows.data_ptr(), cat.data_ptr(), out.data_ptr(), 2, 1,
              err.data_ptr())
    assert int(err.item()) == 1
    assert torch.equal(out[0, 0], t[1])


def test_bad_arguments_are_rejected():
    from nasrec_b200 import _lib
    with pytest.raises(ValueError):
        _lib.call("nasrec_ln_fwd", None, 4, 1, 4, None, None, 1e-5, 0, 4, None, 4, None, None, 0)


def test_bce_norm_clip_adagrad_kernels():
    from nasrec_b200 import _lib
    from nasrec_b200 import engine as eng
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    z = torch.randn(777, generator=g) * 3
    y = (torch.rand(777, generator=g) < 0.3).float()
    loss, dl = eng.bce_with_logits(z.to(dev), y.to(dev))
    zr = z.clone().requires_grad_(True)
    lr_ = torch.nn.functional.binary_cross_entropy_with_logits(zr, y)
    lr_.backward()
    assert abs(float(loss.item()) - float(lr_.detach())) < 1e-6
    assert rel_err(dl.cpu().numpy(), zr.grad.numpy()) < 1e-5
    # global norm + clip + Adagrad against torch
    shapes = [(1024, 300), (16,), (45, 98), (70000,), (1,)]
    ws = [torch.randn(s, generator=g) for s in shapes]
    gs = [torch.randn(s, generator=g) * 2 for s in shapes]
    params = [torch.nn.Parameter(w.clone()) for w in ws]
    for p, gr in zip(params, gs):
        p.grad = gr.clone()
    opt = torch.optim.Adagrad(params, lr=0.12, eps=1e-2)
    total = torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step()
    wd = [w.clone().to(dev) for w in ws]
    gd = [x.clone().to(dev) for x in gs]
    sd = [torch.zeros_like(w) for w in wd]
    sizes = _lib.i64_array([w.numel() for w in wd])
    nws = _lib.query("nasrec_sumsq_ws_floats", sizes, len(wd))
    partial = torch.empty(nws, device=dev)
    out = torch.empty(2, device=dev)
    _lib.call("nasrec_grad_norm_clip", _lib.ptr_array([x.data_ptr() for x in gd]), sizes, len(wd), None, 0, 5.0,
              partial.data_ptr(), out.data_ptr())
    assert abs(float(out[0].item()) - float(total)) < 1e-4 * float(total)
    _lib.call("nasrec_adagrad_multi", _lib.ptr_array([x.data_ptr() for x in wd]),
              _lib.ptr_array([x.data_ptr() for x in gd]), _lib.ptr_array([x.data_ptr() for x in sd]), sizes, len(wd),
              0.12, 1e-2, out.data_ptr() + 4)
    for a, p in zip(wd, params):
        assert rel_err(a.cpu().numpy(), p.detach().numpy()) < 1e-5   # torch CPU vector width differs per host
