"""Debug: one training step of the B=512 autoctr supernet through the native executor; dumps every parameter after the
step to an .npz (argv[1]) so that two runs under different GEMM plans can be compared tensor by tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import SuperNet, ops_config_lib
from nasrec_b200.utils.train_utils import init_weights
from nasrec_b200.native import NativeTrainer
from oracle import nasrec_oracle as orc
CRITEO = [1461, 584, 10131227, 2202609, 306, 25, 12518, 634, 4, 93146, 5684, 8351593, 3195, 28, 14993, 5461307, 11, 5653, 2174, 5, 7046548, 19, 16, 286182, 106, 142573]
ne = [min(x, 500000) for x in CRITEO]
torch.manual_seed(3); np.random.seed(3)
m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
             path_sampling_strategy="default", anypath_choice="binomial-0.5", supernet_training_steps=0).to("cuda")
m.materialize(13); m.apply(init_weights)
tr = NativeTrainer(m, lr=0.12, overlap_wgrad=not os.environ.get("NO_OVERLAP"))
for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    int_x, cat_x, y = orc.synth_batch(512, 13, ne, seed=1200 + step, zipf=True)
    logits, loss = tr.step(int_x.cuda(), cat_x.cuda(), y.cuda())
torch.cuda.synchronize()
print("loss", float(loss), "choice", m.choice)
np.savez(sys.argv[1], **{k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
if os.environ.get("WITH_ORACLE"):
    torch.manual_seed(3); np.random.seed(3)
    m2 = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
                  path_sampling_strategy="default", anypath_choice="binomial-0.5", supernet_training_steps=0).to("cuda")
    m2.materialize(13); m2.apply(init_weights)
    sd0 = {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()}
    ref = orc.OracleTrainer(sd0, dict(ops="autoctr", use_layernorm=True, fixed=False, num_blocks=7), lr=0.12)
    int_x, cat_x, y = orc.synth_batch(512, 13, ne, seed=1200, zipf=True)
    ref.step(m.choice if len(sys.argv) <= 2 or sys.argv[2] == "1" else None, int_x, cat_x, y)
    new = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    rows = []
    for k in new:
        if k in ref.params:
            d = (new[k] - ref.params[k].detach()).abs()
            rows.append((float(d.max()), k, tuple(new[k].shape), int((d > 1e-5).sum())))
    for d, k, s, n in sorted(rows, reverse=True)[:14]: print("vs oracle %.3e %s %s n>1e-5: %d" % (d, k, s, n))
if os.environ.get("ORACLE_SPLIT"):
    # conditioning probe, CPU only: the fp32 oracle against ITSELF with every linear whose K >= 256 computed as two half-K
    # products added (what a split-K GEMM does to the summation order).  Same weights, same batch, same code otherwise.
    import torch.nn.functional as F
    real_linear = F.linear
    def split_linear(x, w, b=None):
        K = w.shape[1]
        if K < 256 or x.dim() != 2:
            return real_linear(x, w, b)
        h = (K // 2 + 31) // 32 * 32
        z = real_linear(x[:, :h], w[:, :h]) + real_linear(x[:, h:], w[:, h:])
        return z if b is None else z + b
    ref2 = orc.OracleTrainer(sd0, dict(ops="autoctr", use_layernorm=True, fixed=False, num_blocks=7), lr=0.12)
    F.linear = split_linear
    try:
        ref2.step(m.choice, int_x, cat_x, y)
    finally:
        F.linear = real_linear
    rows = []
    for k in ref.params:
        d = (ref2.params[k].detach() - ref.params[k].detach()).abs()
        rows.append((float(d.max()), k, tuple(d.shape), int((d > 1e-5).sum())))
    for d, k, s, n in sorted(rows, reverse=True)[:8]: print("oracle(split-K linears) vs oracle %.3e %s %s n>1e-5: %d" % (d, k, s, n))
if os.environ.get("ORACLE_FLIP"):
    # kink probe, CPU only: find the ReLU input closest to zero in the oracle's forward, push it across zero (a constant
    # offset of twice its value on that single element), redo the oracle step, and compare with the GPU's weights.
    orig_relu = torch.relu
    rec = []
    def spy(x):
        a = x.detach().abs().flatten()
        i = int(a.argmin())
        rec.append((float(a[i]) / max(float(a.median()), 1e-30), len(rec), i, float(x.detach().flatten()[i])))
        return orig_relu(x)
    torch.relu = spy
    try:
        with torch.no_grad():
            orc.supernet_forward(sd0, dict(ops="autoctr", use_layernorm=True, fixed=False, num_blocks=7), m.choice, int_x, cat_x)
    finally:
        torch.relu = orig_relu
    rec.sort()
    print("closest ReLU inputs (|x|/median, call, flat index, x):", rec[:4])
    for margin, call, idx, val in rec[:int(os.environ["ORACLE_FLIP"])]:
        cnt = [0]
        def flip(x, call=call, idx=idx, val=val):
            c = cnt[0]; cnt[0] += 1
            if c == call:
                off = torch.zeros_like(x).flatten(); off[idx] = -2.0 * val
                x = x + off.view_as(x)
            return orig_relu(x)
        ref3 = orc.OracleTrainer(sd0, dict(ops="autoctr", use_layernorm=True, fixed=False, num_blocks=7), lr=0.12)
        torch.relu = flip
        try:
            ref3.step(m.choice, int_x, cat_x, y)
        finally:
            torch.relu = orig_relu
        rows = []
        for k in new:
            if k in ref3.params:
                d = (new[k] - ref3.params[k].detach()).abs()
                rows.append((float(d.max()), k, int((d > 1e-5).sum())))
        print("GPU vs oracle with ReLU call %d element %d (x = %.3e, margin %.2e) flipped:" % (call, idx, val, margin), sorted(rows, reverse=True)[:3])
