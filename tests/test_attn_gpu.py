"""GPU parity of the varlen tcgen05 attention forward against per-sequence fp32 softmax attention."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attn(qkv, cu, H):
    T, D3 = qkv.shape
    D = D3 // 3
    d = D // H
    out = torch.empty(T, D, device=qkv.device)
    lse = torch.empty(H, T, device=qkv.device)
    q, k, v = qkv.float().split(D, dim=1)
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        for h in range(H):
            qs, ks, vs = (t[s:e, h * d:(h + 1) * d] for t in (q, k, v))
            sc = qs @ ks.t() / math.sqrt(d)
            lse[h, s:e] = torch.logsumexp(sc, -1)
            out[s:e, h * d:(h + 1) * d] = torch.softmax(sc, -1) @ vs
    return out, lse


@pytest.mark.parametrize("D,H", [(192, 2), (32, 2), (192, 12), (768, 12), (64, 2), (256, 2)])
@pytest.mark.parametrize("counts,npatch", [([1, 3, 10, 5], 196), ([2, 10, 1], 36), ([1], 4), ([7, 2, 2, 9, 1, 4], 196)])
def test_attn_fwd(D, H, counts, npatch):
    from chadavit_b200 import ops
    lay = ops.PackedLayout(counts, npatch, "cuda")
    g = torch.Generator(device="cpu").manual_seed(D + H + npatch)
    qkv = (torch.randn(lay.T, 3 * D, generator=g) * 1.5).to(torch.bfloat16).cuda()
    out, lse = ops.attn_fwd(qkv, lay, H)
    ops.sync_check()
    ref, ref_lse = _ref_attn(qkv, lay.cu_host.tolist(), H)
    err = (out.float() - ref).abs().max().item()
    lerr = (lse - ref_lse).abs().max().item()
    print(f"attn D={D} H={H} counts={counts} N={npatch}: out err {err:.3e}  lse err {lerr:.3e}")
    assert err < 3e-2 and lerr < 2e-2


@pytest.mark.parametrize("D,H", [(192, 2), (32, 2), (192, 12), (768, 12)])
@pytest.mark.parametrize("counts,npatch", [([1, 3, 10, 5], 196), ([2, 10, 1], 36), ([1], 4)])
def test_attn_bwd(D, H, counts, npatch):
    from chadavit_b200 import ops
    lay = ops.PackedLayout(counts, npatch, "cuda")
    g = torch.Generator(device="cpu").manual_seed(7 * D + H + npatch)
    qkv = (torch.randn(lay.T, 3 * D, generator=g) * 1.2).to(torch.bfloat16).cuda()
    dout = torch.randn(lay.T, D, generator=g).to(torch.bfloat16).cuda()
    out, lse = ops.attn_fwd(qkv, lay, H)
    dqkv = ops.attn_bwd(dout, qkv, out, lse, lay, H)
    ops.sync_check()
    qr = qkv.float().requires_grad_()
    d = D // H
    cu = lay.cu_host.tolist()
    q, k, v = qr.split(D, dim=1)
    outs = []
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        hs = []
        for h in range(H):
            qs, ks, vs = (t[s:e, h * d:(h + 1) * d] for t in (q, k, v))
            hs.append(torch.softmax(qs @ ks.t() / math.sqrt(d), -1) @ vs)
        outs.append(torch.cat(hs, 1))
    torch.cat(outs, 0).backward(dout.float())
    ref = qr.grad
    names = ("dq", "dk", "dv")
    errs = [(dqkv.float()[:, i * D:(i + 1) * D] - ref[:, i * D:(i + 1) * D]).abs().max().item() for i in range(3)]
    scale = ref.abs().max().item()
    print(f"attn bwd D={D} H={H} counts={counts} N={npatch}: " + " ".join(f"{n} {e:.3e}" for n, e in zip(names, errs)) + f" (ref max {scale:.2f})")
    assert max(errs) < 0.03 * max(1.0, scale)
