"""Attention core (csrc/attn.cu) forward / backward launch times by batch and token count, CUDA events, L2 flushed.
Run once as is and once with NASREC_ATTN_OLD=1 (one thread per token) to compare; NASREC_ATTN_FWD4_MAXB moves the batch
size above which forward keeps the one-thread-per-token kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
dev = torch.device("cuda")
E = 16
sizes = [768, 48, 256, 16, 16, 16, 256, 16, 256, 16, 16, 16]
g = torch.Generator(device="cpu").manual_seed(3)
params = [(torch.randn(n, generator=g) * 0.2).to(dev) for n in sizes]
params[4] += 1.0; params[10] += 1.0
pp = _lib.ptr_array([p.data_ptr() for p in params])
flush = torch.empty(64 << 20, device=dev)
tag = "old" if os.environ.get("NASREC_ATTN_OLD") else "new(maxb=%s)" % os.environ.get("NASREC_ATTN_FWD4_MAXB", "1024")
for B, L in ((256, 16), (256, 32), (256, 48), (256, 64), (512, 32), (2048, 32), (8192, 32)):
    s = L
    x = torch.randn(B, s, E, device=dev); dy = torch.randn(B, s, E, device=dev)
    y = torch.empty(B, s, E, device=dev); dx = torch.empty(B, s, E, device=dev)
    dpar = torch.empty(_lib.ATTN_PARAMS, device=dev)
    ws = torch.empty(_lib.query("nasrec_attn_bwd_ws_floats", B), device=dev)
    fns = {"fwd": lambda: _lib.call("nasrec_attn_fwd", x.data_ptr(), s * E, L, s, pp, y.data_ptr(), s * E, B),
           "bwd": lambda: _lib.call("nasrec_attn_bwd", dy.data_ptr(), s * E, x.data_ptr(), s * E, L, s, pp, dx.data_ptr(), s * E,
                                    dpar.data_ptr(), 0, ws.data_ptr(), B)}
    res = {}
    for name, fn in fns.items():
        fn(); torch.cuda.synchronize()
        tot, n = 0.0, 8
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot / n * 1e3
    print("%s B=%d L=%d: fwd %.1f us, bwd (+param reduce) %.1f us  checksum %.6e %.6e" % (tag, B, L, res["fwd"], res["bwd"], float(y.double().sum()), float(dx.double().sum() + dpar.double().sum())), flush=True)
