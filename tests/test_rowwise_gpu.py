"""GPU parity of the row-wise kernels (LayerNorm fwd/bwd, column sums, tokenizer) against the fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


@pytest.mark.parametrize("D", [64, 192, 256])
def test_layernorm2_fwd(D):
    """norm2 of block i fused with norm1 of block i+1 (chada_vit.py:100, :96) == two torch LayerNorms."""
    from chadavit_b200 import ops
    T = 1237
    x = _rand((T, D), 11, 2.0, dtype=torch.float32)
    ga, gb = 1 + 0.2 * _rand((D,), 12, dtype=torch.float32), 1 + 0.2 * _rand((D,), 13, dtype=torch.float32)
    ba, bb = 0.1 * _rand((D,), 14, dtype=torch.float32), 0.1 * _rand((D,), 15, dtype=torch.float32)
    y1, y2, ma, ra, mb, rb = ops.layernorm2_fwd(x, ga, ba, 1e-5, gb, bb, 1e-6)
    ops.sync_check()
    r1 = F.layer_norm(x, (D,), ga, ba, 1e-5)
    r2 = F.layer_norm(r1, (D,), gb, bb, 1e-6)
    assert (y1 - r1).abs().max().item() < 1e-5 and (y2.float() - r2).abs().max().item() < 0.03
    _, _, m_ref, r_ref = ops.layernorm_fwd(x, ga, ba, 1e-5)
    _, _, m2_ref, r2_ref = ops.layernorm_fwd(y1, gb, bb, 1e-6)
    assert (ma - m_ref).abs().max().item() < 1e-6 and (ra - r_ref).abs().max().item() < 1e-4 * r_ref.abs().max().item()
    assert (mb - m2_ref).abs().max().item() < 1e-5 and (rb - r2_ref).abs().max().item() < 1e-4 * r2_ref.abs().max().item()


@pytest.mark.parametrize("D", [32, 64, 192, 256, 768])
def test_layernorm_fwd_bwd(D):
    from chadavit_b200 import ops
    T = 1237
    x = _rand((T, D), 1, 2.0, dtype=torch.float32)
    gamma = 1 + 0.2 * _rand((D,), 2, dtype=torch.float32)
    beta = 0.1 * _rand((D,), 3, dtype=torch.float32)
    y, y32, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-5, out_f32=True)
    xr = x.clone().requires_grad_()
    gr, br = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    yr = F.layer_norm(xr, (D,), gr, br, 1e-5)
    assert (y32 - yr).abs().max().item() < 1e-5 and (y.float() - yr).abs().max().item() < 0.03
    dy = _rand((T, D), 4, dtype=torch.float32)
    dres = _rand((T, D), 5, dtype=torch.float32)
    yr.backward(dy)
    dg, db, dc = (torch.zeros(D, device="cuda") for _ in range(3))
    dx32, dx16 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma=dg, dbeta=db, dcolsum=dc, dres=dres, want_bf16=True)
    ops.sync_check()
    assert (dx32 - (xr.grad + dres)).abs().max().item() < 1e-4
    assert (dx16.float() - dx32).abs().max().item() < 0.05
    assert (dg - gr.grad).abs().max().item() < 1e-3 * gr.grad.abs().max().item() + 1e-3
    assert (db - br.grad).abs().max().item() < 1e-3 * br.grad.abs().max().item() + 1e-3
    assert (dc - xr.grad.sum(0)).abs().max().item() < 1e-2
    # gathered rows (final norm + CLS select) and scatter in the backward
    idx = torch.tensor([0, 5, 77, 1236], dtype=torch.int32, device="cuda")
    _, g32, m2, r2 = ops.layernorm_fwd(x, gamma, beta, 1e-6, in_idx=idx, out_bf16=False, out_f32=True)
    ref = F.layer_norm(x[idx.long()], (D,), gamma, beta, 1e-6)
    assert (g32 - ref).abs().max().item() < 1e-5
    dyf = _rand((4, D), 6, dtype=torch.float32)
    dg2, db2 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dxs, _ = ops.layernorm_bwd(dyf, x, gamma, m2, r2, dgamma=dg2, dbeta=db2, idx=idx)
    xr2 = x.clone().requires_grad_()
    F.layer_norm(xr2[idx.long()], (D,), gamma, beta, 1e-6).backward(dyf)
    assert (dxs - xr2.grad).abs().max().item() < 1e-4


def test_colsum_cast():
    from chadavit_b200 import ops
    x = _rand((3001, 2048), 7)
    out = torch.ones(2048, device="cuda")
    ops.colsum(x, out)
    ref = x.float().sum(0) + 1
    assert (out - ref).abs().max().item() < 1e-2
    f = _rand((1000003,), 8, dtype=torch.float32)
    assert torch.equal(ops.cast_bf16(f), f.to(torch.bfloat16))


@pytest.mark.parametrize("D,hw,counts,max_ch", [(32, 224, [1, 3, 5, 10], 10), (192, 96, [2, 10, 1], 10), (192, 224, [3, 1], 3)])
def test_tokenizer_fwd_bwd(D, hw, counts, max_ch):
    """Packed tokens == the reference's padded tokens at every real position; index maps bit-exact."""
    from chadavit_b200 import ops
    from oracle import chada_oracle as O, det
    npatch = (hw // 16) ** 2
    P = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.backbone_shapes(D, depth=0, max_ch=max_ch), 3).items()}
    x = torch.from_numpy(det.det_pixels(sum(counts), hw, hw, 5))
    # oracle (padded) -> gather real rows
    Pq = dict(P)
    Pq["token_learner.proj.weight"] = P["token_learner.proj.weight"].to(torch.bfloat16).float()
    emb, mask = O.tokenize_padded(x.to(torch.bfloat16).float(), counts, Pq, 16, max_ch)
    cu, rows = O.packed_index(counts, npatch)
    ref = emb.reshape(-1, D)[rows]
    assert not mask.reshape(-1)[rows].any() and int((~mask).sum()) == len(rows)   # bookkeeping: real rows == unmasked rows
    lay = ops.PackedLayout(counts, npatch, "cuda", max_channels=10)
    assert lay.cu_host.tolist() == cu
    dev = {k: v.cuda() for k, v in P.items()}
    pos_patch = O.interp_pos_embed(P["pos_embed"], npatch, hw, hw, 16)[0, 0].contiguous().cuda()
    pos0 = P["pos_embed"][0, 0, 0].contiguous().cuda()
    cls_tok = P["cls_token"][0, 0].contiguous().cuda()
    chan = dev["channel_token"][0, :, 0].contiguous() if max_ch == 10 else None
    w_bf = dev["token_learner.proj.weight"].reshape(D, 256).to(torch.bfloat16)
    tok, patches = ops.tokenize_fwd(x.cuda(), lay, 16, w_bf, dev["token_learner.proj.bias"], pos_patch, pos0, cls_tok, chan)
    ops.sync_check()
    assert tok.dtype == torch.float32
    err = (tok.cpu() - ref).abs().max().item()
    print(f"tokenizer D={D} hw={hw}: max err {err:.3e}")
    assert err < 2e-2
    # backward: parameter gradients for a random dTok
    g = torch.Generator(device="cpu").manual_seed(1)
    dtok = torch.randn(lay.T, D, generator=g).to(torch.bfloat16)
    Pg = {k: v.clone().requires_grad_() for k, v in Pq.items()}
    emb2, _ = O.tokenize_padded(x.to(torch.bfloat16).float(), counts, Pg, 16, max_ch)
    (emb2.reshape(-1, D)[rows] * dtok.float()).sum().backward()
    dw = torch.zeros(D, 256, device="cuda"); dbias = torch.zeros(D, device="cuda")
    dpos = torch.zeros(npatch, D, device="cuda"); dcls = torch.zeros(D, device="cuda"); dpos0 = torch.zeros(D, device="cuda")
    dchan = torch.zeros(10, D, device="cuda") if max_ch == 10 else None
    ops.tokenize_bwd(dtok.cuda(), patches, lay, dw_pe=dw, db_pe=dbias, dpos_patch=dpos, dpos0=dpos0, dcls_tok=dcls, dchan_tok=dchan)
    ops.sync_check()
    assert (dw.cpu() - Pg["token_learner.proj.weight"].grad.reshape(D, 256)).abs().max().item() < 0.05
    assert (dbias.cpu() - Pg["token_learner.proj.bias"].grad).abs().max().item() < 0.05
    assert (dcls.cpu() - Pg["cls_token"].grad[0, 0]).abs().max().item() < 1e-3
    assert (dpos0.cpu() - Pg["pos_embed"].grad[0, 0, 0]).abs().max().item() < 1e-3
    if hw == 224:
        assert (dpos.cpu() - Pg["pos_embed"].grad[0, 0, 1:]).abs().max().item() < 2e-3
    if max_ch == 10:
        assert (dchan.cpu() - Pg["channel_token"].grad[0, :, 0]).abs().max().item() < 0.05
