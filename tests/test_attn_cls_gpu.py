"""GPU parity of the CLS-query attention of the last block (cb_attn_cls_fwd / cb_attn_cls_bwd, through the C ABI) against fp32
torch softmax attention of the same bf16 operands, and against the dense varlen kernels on the CLS rows
(reference: nn.MultiheadAttention inside the encoder layer, src/backbones/vit/chada_vit.py:105-111, with `return x[:, 0]`, :289)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(qkv, cu, H, d):
    """fp32 attention of every sequence's first query against all of its keys: (out [B, H*d], differentiable in qkv)."""
    D = H * d
    outs = []
    for b in range(len(cu) - 1):
        seq = qkv[cu[b]:cu[b + 1]]
        S = seq.shape[0]
        q = seq[0, :D].view(H, d)
        k = seq[:, D:2 * D].view(S, H, d)
        v = seq[:, 2 * D:].view(S, H, d)
        s = torch.einsum("hd,shd->hs", q, k) * d ** -0.5
        p = torch.softmax(s, dim=-1)
        outs.append(torch.einsum("hs,shd->hd", p, v).reshape(D))
    return torch.stack(outs)


@pytest.mark.parametrize("H,d", [(2, 96), (2, 16), (12, 64), (3, 32), (1, 128)])
@pytest.mark.parametrize("counts,npatch", [([1, 3, 10, 2], 196), ([1, 1, 1, 1, 1, 10, 4], 36), ([7], 4), ([10] * 3 + [1], 200)])
def test_attn_cls_fwd_bwd(H, d, counts, npatch):
    from chadavit_b200 import ops
    lay = ops.PackedLayout(counts, npatch, torch.device("cuda"))
    T, D = lay.T, H * d
    g = torch.Generator(device="cpu").manual_seed(1000 * H + d + len(counts))
    qkv = (torch.randn(T, 3 * D, generator=g) * 0.8).to(torch.bfloat16).cuda()
    dout = torch.randn(lay.B, D, generator=g).to(torch.bfloat16).cuda()
    cu = lay.cu_host.tolist()

    out, lse = ops.attn_cls_fwd(qkv, lay, H)
    ops.sync_check()
    qf = qkv.float().requires_grad_()
    ref = _reference(qf, cu, H, d)
    (ref * dout.float()).sum().backward()
    e_out = (out.float() - ref).abs().max().item()
    assert e_out <= 1e-2 * max(1.0, ref.abs().max().item()), e_out

    # the dense varlen kernel on the same rows (bf16 probabilities in its P V product)
    dense, _ = ops.attn_fwd(qkv, lay, H, need_lse=False)
    e_dense = (dense[lay.cls_rows64()].float() - out.float()).abs().max().item()
    assert e_dense <= 2e-2 * max(1.0, ref.abs().max().item()), e_dense

    dqkv = ops.attn_cls_bwd(dout, qkv, out, lse, lay, H)
    ops.sync_check()
    assert dqkv.shape == qkv.shape and dqkv.dtype == torch.bfloat16
    gq = qf.grad
    e = ((dqkv.float() - gq).norm() / gq.norm()).item()
    e_max = (dqkv.float() - gq).abs().max().item()
    print(f"attn_cls H={H} d={d} B={lay.B} T={T}: out max err {e_out:.2e}, vs dense kernel {e_dense:.2e}, dqkv rel err {e:.2e} (max abs {e_max:.2e})")
    assert e <= 1e-2
    assert e_max <= 2e-2 * max(1.0, gq.abs().max().item())
    # dQ of every non-CLS row is exactly zero and is written (the buffer comes from torch.empty)
    mask = torch.ones(T, dtype=torch.bool, device="cuda")
    mask[lay.cls_rows64()] = False
    assert bool((dqkv[mask][:, :D] == 0).all())
