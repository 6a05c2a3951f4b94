"""Sorted-row reduction of the embedding gradient: one CTA per table (emb.cu) vs the multi-CTA path (emb_big.cu) at the
global batch sizes of data-parallel training (N x 512 ids per table), Zipf ids."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from nasrec_b200 import _lib
dev = torch.device("cuda")
ne = [min(x, bench.CAP) for x in bench._CRITEO]
F = len(ne)
rows = torch.tensor(ne, dtype=torch.int64, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(64 << 20, device=dev)
for B in (512, 1024, 2048, 4096, 8192):
    cat = torch.from_numpy(bench.synth_pool(1, B, 13, ne, 5, zipf=True)[0][1]).to(dev)
    gout = torch.randn(B, F, 16, device=dev)
    uniq = torch.empty(F, B, dtype=torch.int64, device=dev); nuniq = torch.empty(F, dtype=torch.int32, device=dev)
    rg = torch.empty(F, B, 16, device=dev); sumsq = torch.empty(F, device=dev)
    scratch = torch.empty(F, B + 1, dtype=torch.int32, device=dev)
    nb = _lib.query("nasrec_emb_grad_sort_reduce_big_ws_bytes", B, F)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    fns = {"one_cta": lambda: _lib.call("nasrec_emb_grad_sort_reduce_checked", cat.data_ptr(), rows.data_ptr(), err.data_ptr(), gout.data_ptr(), B, F, uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(), scratch.data_ptr()),
           "big": lambda: _lib.call("nasrec_emb_grad_sort_reduce_big", cat.data_ptr(), rows.data_ptr(), err.data_ptr(), gout.data_ptr(), B, F, uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(), ws.data_ptr(), nb)}
    res = {}
    for name, fn in fns.items():
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot / 10 * 1e3
    print("B=%d ids per table: one CTA per table %.1f us, multi-CTA %.1f us" % (B, res["one_cta"], res["big"]), flush=True)
