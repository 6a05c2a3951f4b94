"""CPU: the device code of csrc/attn.cu compiled for the host (one OS thread per CUDA thread, pthread barrier for
__syncthreads, tests/host_emul/attn_emul.cpp): the four-threads-per-token attention kernels must reproduce the
one-thread-per-token kernels -- which the GPU tests pin to the oracle -- bit for bit (same operations, same order per
output), over ragged token counts, masked rows, several samples per CTA and a null dX."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_code():
    src = open(os.path.join(ROOT, "nasrec_b200", "csrc", "attn.cu")).read()
    a = src.index("namespace {")
    b = src.index("int attn_grid(int B)")
    body = src[a:b]
    body = body.replace("extern __shared__ float sm[];", "float* sm = g_dyn;")
    assert "__shfl" not in body and "atomic" not in body          # nothing the thread emulation cannot express
    return body


def _fm_device_code():
    src = open(os.path.join(ROOT, "nasrec_b200", "csrc", "interact.cu")).read()
    a = src.index("namespace {")
    b = src.index("// ------------------------------------------------------------------ gating / concat")
    body = src[a:b]
    assert "__shfl" not in body and "atomic" not in body
    return body


def _build(tmp_path, name="attn", code=None):
    (tmp_path / ("%s_device_code.inc" % ("attn" if name == "attn" else "interact"))).write_text(code or _device_code())
    exe = tmp_path / (name + "_emul")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", str(tmp_path),
           "-I", os.path.join(ROOT, "tests", "host_emul"),
           os.path.join(ROOT, "tests", "host_emul", name + "_emul.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_four_threads_per_token_kernels_match_one_thread_per_token_kernels_and_the_oracle(tmp_path):
    import numpy as np
    import torch
    from oracle import nasrec_oracle as orc
    exe = _build(tmp_path)
    dump = tmp_path / "case2.bin"
    r = subprocess.run([str(exe), str(dump)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-1000:]
    assert len(re.findall(r" 0 differ", r.stdout)) == 21          # 7 cases x (y, dX, parameter-gradient partials)

    # the emulated kernels against the oracle's restatement of the reference (modules.py:664-688) on the dumped case:
    # tokens s..L-1 are the reference's zero rows and take part as keys / values
    raw = np.fromfile(dump, dtype=np.float32)
    B, L, s, ybs = (int(v) for v in raw[:4].view(np.int32))
    off = [4]

    def take(n):
        a = raw[off[0]:off[0] + n]
        off[0] += n
        return torch.from_numpy(a.copy())
    P = take(1696)
    x = take(B * s * 16).reshape(B, s, 16)
    dy = take(B * ybs).reshape(B, ybs)[:, :s * 16].reshape(B, s, 16)
    y = take(B * ybs).reshape(B, ybs)[:, :s * 16].reshape(B, s, 16)
    dx = take(B * s * 16).reshape(B, s, 16)
    dpar = take(1696)
    names = [("m.in_proj_weight", (48, 16)), ("m.in_proj_bias", (48,)), ("m.out_proj.weight", (16, 16)),
             ("m.out_proj.bias", (16,)), ("ln1.weight", (16,)), ("ln1.bias", (16,)), ("fc1.weight", (16, 16)),
             ("fc1.bias", (16,)), ("fc2.weight", (16, 16)), ("fc2.bias", (16,)), ("ln2.weight", (16,)), ("ln2.bias", (16,))]
    sd, o = {}, 0
    for k, shp in names:
        n = int(np.prod(shp))
        sd[k] = P[o:o + n].reshape(shp).clone().requires_grad_(True)
        o += n
    assert o == 1696
    xr = x.clone().requires_grad_(True)
    p = torch.cat([xr, torch.zeros(B, L - s, 16)], 1)
    a = orc._ln(sd, "ln1", orc.mha_self(sd, "m", p) + p)
    f = orc._lin(sd, "fc2", torch.relu(orc._lin(sd, "fc1", a)))
    out = orc._ln(sd, "ln2", a + f)[:, :s]
    (out * dy).sum().backward()
    ref = torch.cat([sd[k].grad.reshape(-1) for k, _ in names])
    print("max abs diff y / dX / dparam:", float((out.detach() - y).abs().max()), float((xr.grad - dx).abs().max()),
          float((ref - dpar).abs().max()))
    assert float((out.detach() - y).abs().max()) < 2e-5
    assert float((xr.grad - dx).abs().max()) < 1e-4 * max(1.0, float(xr.grad.abs().max()))
    assert float((ref - dpar).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fm_backward_one_cta_per_sample_matches_thread_per_column_kernel(tmp_path):
    """interact.cu: fm_bwd_rows_kernel (used up to B = 2048) against fm_bwd_kernel, bit for bit, incl. in-place accumulate."""
    exe = _build(tmp_path, "fm", _fm_device_code())
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-1000:]
    assert len(re.findall(r" 0 differ", r.stdout)) == 7


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_dot_product_triangle_and_fm_kernels_match_the_reference_formulas(tmp_path):
    """interact.cu on the host: DotProduct's strict lower triangle (tril_indices row-major order, modules.py:366-383) and
    the FM reductions (modules.py:736-738), forward and backward, against direct double-precision evaluation."""
    (tmp_path / "interact_device_code.inc").write_text(_fm_device_code())
    exe = tmp_path / "interact_emul"
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", str(tmp_path),
           "-I", os.path.join(ROOT, "tests", "host_emul"),
           os.path.join(ROOT, "tests", "host_emul", "interact_emul.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-1000:]
