import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nasrec_oracle as orc
from tests.test_gpu_modules import _randomize
from tests.helpers import rel_err
from nasrec_b200.supernet.modules import ElasticLinear
ln, fixed, K, maxd, d = True, False, 300, 1024, 64
mod = ElasticLinear(fixed=fixed, use_layernorm=ln, max_dims_or_dims=maxd, activation="relu").cuda()
x = torch.randn(37, K)
xg = x.clone().cuda().requires_grad_(True)
out = mod(xg, d); _randomize(mod, 0)
for rep in range(2):
    for p in mod.parameters(): p.grad = None
    xg = x.clone().cuda().requires_grad_(True)
    out = mod(xg, d)
    R = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * R.cuda()).sum().backward(); torch.cuda.synchronize()
    sd = {"p." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xc = x.clone().requires_grad_(True)
    ref = orc.fc(sd, "p", xc, d, maxd, ln, fixed); (ref * R).sum().backward()
    print("rep", rep, "out", rel_err(out.detach().cpu().numpy(), ref.detach().numpy()), "x.grad", rel_err(xg.grad.cpu().numpy(), xc.grad.numpy()))
    for n, p in mod.named_parameters():
        print("   ", n, rel_err(p.grad.cpu().numpy(), sd["p." + n].grad.numpy()))
