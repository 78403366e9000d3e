"""GPU parity of the tcgen05 GEMM (through the C ABI) against fp32 matmul of the same bf16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _ref(A, B, a_mn, b_mn):
    a = A.float().t() if a_mn else A.float()
    b = B.float().t() if b_mn else B.float()
    return a @ b.t()


SHAPES = [
    # M, N, K
    (128, 128, 64), (256, 192, 192), (300, 576, 192), (1000, 2048, 192), (520, 192, 2048),
    (130, 96, 32), (77, 32, 2048), (128, 4096, 256), (197, 256, 2048), (64, 2048, 2048),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_kmajor(M, N, K):
    from chadavit_b200 import ops
    A, B = _rand((M, K), 1), _rand((N, K), 2, 0.1)
    C = ops.gemm(A, B, flags=ops.EPI_OUT_F32)
    ops.sync_check()
    ref = _ref(A, B, False, False)
    err = (C - ref).abs().max().item()
    print(f"gemm K-major M={M} N={N} K={K}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})")
    assert err <= 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(256, 192, 320), (192, 2048, 1000), (2048, 192, 776), (576, 192, 640), (96, 32, 304), (32, 96, 200), (200, 192, 576)])
def test_gemm_mn_major(M, N, K, a_mn, b_mn):
    from chadavit_b200 import ops
    if a_mn and M % 32:
        pytest.skip("MN-major A needs M % 32 == 0")
    if a_mn and not b_mn:
        pytest.skip("combination not instantiated")
    A = _rand((K, M) if a_mn else (M, K), 3)
    B = _rand((K, N) if b_mn else (N, K), 4, 0.1)
    C = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, flags=ops.EPI_OUT_F32)
    ops.sync_check()
    ref = _ref(A, B, a_mn, b_mn)
    err = (C - ref).abs().max().item()
    print(f"gemm a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K}: max err {err:.3e} (ref max {ref.abs().max().item():.2f})")
    assert err <= 2e-3 * max(1.0, ref.abs().max().item())


def test_gemm_epilogues():
    from chadavit_b200 import ops
    M, N, K = 333, 192, 2048
    A, B = _rand((M, K), 5), _rand((N, K), 6, 0.05)
    bias = torch.randn(N, device="cuda")
    res = _rand((M, N), 7)
    res32 = res.float() * 1.7
    ref = _ref(A, B, False, False) + bias
    C = ops.gemm(A, B, bias=bias, aux=res32, flags=ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32)
    assert (C - (ref + res32)).abs().max().item() < 1e-2
    C = ops.gemm(A, B, bias=bias, flags=ops.EPI_RELU)
    assert (C.float() - ref.relu()).abs().max().item() < 0.05
    cs = torch.ones(N, device="cuda")
    C = ops.gemm(A, B, aux=res, flags=ops.EPI_RELU_MASK, colsum=cs)
    masked = _ref(A, B, False, False) * (res.float() > 0)
    assert (C.float() - masked).abs().max().item() < 0.05
    assert (cs - 1 - C.float().sum(0)).abs().max().item() < 1e-2       # fused bias-gradient column sums (of the stored bf16 values)
    assert (cs - 1 - masked.sum(0)).abs().max().item() < 0.5
    # split-K atomic accumulation into an existing fp32 buffer
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(A, B, flags=ops.EPI_ATOMIC, out=acc, k_splits=7)
    ops.sync_check()
    assert (acc - 1 - _ref(A, B, False, False)).abs().max().item() < 1e-2


@pytest.mark.parametrize("M,N,K", [(512, 576, 192), (4999, 576, 192), (68664, 576, 192), (1000, 192, 192), (777, 192, 64), (129 * 128 + 1, 384, 128)])
@pytest.mark.parametrize("relu", [False, True])
def test_gemm_bf16_tma_store_epilogue(M, N, K, relu):
    """bf16 outputs of the weights-resident 192-wide products (qkv projection, chada_vit.py:105-111 in_proj; the out-projection's
    input gradient) leave through swizzled shared-memory boxes and TMA stores.  Ragged last row tile (clipped by the tensor map),
    bias, alpha, ReLU, and an output that is a column slice of a wider buffer (ldc > N); the neighbouring columns stay untouched."""
    from chadavit_b200 import ops
    A, B = _rand((M, K), 31), _rand((N, K), 32, 0.1)
    bias = torch.randn(N, device="cuda")
    ref = 0.5 * _ref(A, B, False, False) + bias
    ref = ref.relu() if relu else ref
    C = ops.gemm(A, B, bias=bias, alpha=0.5, flags=ops.EPI_RELU if relu else 0)
    ops.sync_check()
    assert C.dtype == torch.bfloat16
    assert torch.equal(C, ref.to(torch.bfloat16)) or (C.float() - ref).abs().max().item() <= 8e-3 * max(1.0, ref.abs().max().item())
    wide = torch.full((M, N + 64), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, alpha=0.5, flags=ops.EPI_RELU if relu else 0, out=wide[:, 32:32 + N])
    ops.sync_check()
    assert torch.equal(wide[:, 32:32 + N], C)
    assert bool((wide[:, :32] == 7.0).all()) and bool((wide[:, 32 + N:] == 7.0).all())


def test_gemm_weight_grad_splitk():
    """dW[out,in] += dY^T X with both operands MN-major and split-K (the encoder's weight-gradient product)."""
    from chadavit_b200 import ops
    T, out_f, in_f = 5000, 2048, 192
    dY, X = _rand((T, out_f), 8, 0.1), _rand((T, in_f), 9)
    dW = torch.zeros(out_f, in_f, device="cuda")
    ops.gemm(dY, X, a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=dW, k_splits=ops.splitk_for(T, 16))
    ops.sync_check()
    ref = dY.float().t() @ X.float()
    err = (dW - ref).abs().max().item()
    print(f"dW split-K: max err {err:.3e} (ref max {ref.abs().max().item():.2f})")
    assert err <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("T,out_f,splits", [(5000, 2048, 0), (4999, 576, 0), (68664, 2048, 0), (300, 576, 1), (130, 64, 3)])
def test_gemm_weight_grad_with_bias_grad(T, out_f, splits):
    """The weight-gradient product also delivers the bias gradient dY.sum(0) (F.linear's autograd: grad_bias = grad_output.sum(0))
    through `colsum`: sums of the A operand over K, formed on the tensor pipe next to the tile.  Both outputs accumulate."""
    from chadavit_b200 import ops
    in_f = 192
    assert ops.gemm_rowsum_ok(in_f)
    dY, X = _rand((T, out_f), 18, 0.1), _rand((T, in_f), 19)
    dW = torch.full((out_f, in_f), 0.5, device="cuda")
    db = torch.full((out_f,), -2.0, device="cuda")
    ops.gemm(dY, X, a_mn=True, b_mn=True, flags=ops.EPI_ATOMIC, out=dW, k_splits=splits or ops.splitk_for(T, (out_f + 127) // 128 * 2), colsum=db)
    ops.sync_check()
    ref_w = dY.float().t() @ X.float()
    ref_b = dY.double().sum(0)
    err_w = (dW - 0.5 - ref_w).abs().max().item()
    err_b = (db.double() + 2.0 - ref_b).abs().max().item()
    print(f"dW + db T={T} out={out_f}: dW err {err_w:.3e} (max {ref_w.abs().max().item():.2f}), db err {err_b:.3e} (max {ref_b.abs().max().item():.2f})")
    assert err_w <= 2e-3 * max(1.0, ref_w.abs().max().item())
    assert err_b <= 1e-3 * max(1.0, ref_b.abs().max().item())   # fp32 sums of bf16 values: only the summation order differs


@pytest.mark.parametrize("T,F", [(1000, 2048), (128, 64), (129, 128), (257, 2048), (40000, 2048), (128 * 148 * 2 + 5, 512)])
@pytest.mark.parametrize("save_hidden", [True, False])
@pytest.mark.parametrize("gen", [0, 1, 3])
def test_ffn_fused(T, F, save_hidden, gen):
    """cb_ffn_fwd == linear1 -> ReLU -> (bf16 rounding of the hidden activations) -> linear2 + residual (chada_vit.py:113-116, :100).
    gen: the kernel forced through the `kernel` argument (0 = the library's own choice: 3 without / 1 with the hidden store)."""
    from chadavit_b200 import ops
    if gen == 3 and F % 128:
        pytest.skip("the cluster kernel walks the hidden dimension in chunks of 128")
    D = 192
    g = torch.Generator(device="cpu").manual_seed(T + F)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    y = r(T, D).to(torch.bfloat16).cuda()
    w1, w2 = (r(F, D) / D ** 0.5).to(torch.bfloat16).cuda(), (r(D, F) / F ** 0.5).to(torch.bfloat16).cuda()
    b1, b2, resid = r(F).cuda(), r(D).cuda(), r(T, D).cuda()
    z2, hid, bits = ops.ffn_fwd(y, w1, b1, w2, b2, resid, save_hidden=save_hidden, save_mask_bits=True, kernel=gen)
    ops.sync_check()
    h_ref = torch.relu(y.float() @ w1.float().t() + b1).to(torch.bfloat16)
    z_ref = h_ref.float() @ w2.float().t() + b2 + resid
    assert (z2 - z_ref).abs().max().item() < 2e-2 * max(1.0, z_ref.abs().max().item())
    # the ReLU mask as bits: bit j of bits[w, t] <=> hidden[t, 32 w + j] > 0
    assert bits.shape == (F // 32, (T + 31) // 32 * 32) and bits.dtype == torch.int32
    unpacked = ((bits[:, :T].t().unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(T, F).bool()
    if save_hidden:
        assert (hid.float() - h_ref.float()).abs().max().item() < 2e-2 * max(1.0, h_ref.float().abs().max().item())
        assert torch.equal(unpacked, hid > 0)                                   # exactly the stored activations' mask
    else:
        # against the fp32 reference the two can only differ where the pre-activation is at rounding distance from 0
        pre = y.float() @ w1.float().t() + b1
        assert (pre[unpacked != (h_ref > 0)].abs() < 1e-2).all()
        assert hid is None


@pytest.mark.parametrize("T", [200, 128 * 9 + 7, 40000])
def test_relu_mask_bits_equals_bf16_mask(T):
    """d(hidden) = (dz2 . W2) o (hidden > 0) with the mask as 1 bit per unit (CB_EPI_MASK_BITS, the layout cb_ffn_fwd writes) is
    bit-identical to the same product masked by the bf16 hidden activations, in the staged (small M) and the weights-resident
    row-direct epilogue, including the fused bias-gradient column sums."""
    from chadavit_b200 import ops
    D, F = 192, 2048
    g = torch.Generator(device="cpu").manual_seed(T)
    dz = (torch.randn(T, D, generator=g) * 0.1).to(torch.bfloat16).cuda()
    w2 = (torch.randn(D, F, generator=g) / F ** 0.5).to(torch.bfloat16).cuda()
    hid = torch.relu(torch.randn(T, F, generator=g)).to(torch.bfloat16).cuda()
    ld = (T + 31) // 32 * 32
    pos = torch.zeros(ld, F, dtype=torch.int64, device="cuda")
    pos[:T] = (hid > 0).long()
    words = (pos.view(ld, F // 32, 32) << torch.arange(32, device="cuda")).sum(-1)            # [ld, F/32], bit j = column 32 w + j
    bits = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).t().contiguous()
    cs_a, cs_b = torch.zeros(F, device="cuda"), torch.zeros(F, device="cuda")
    ref = ops.gemm(dz, w2, b_mn=True, aux=hid, flags=ops.EPI_RELU_MASK, colsum=cs_a)
    got = ops.gemm(dz, w2, b_mn=True, aux=bits, flags=ops.EPI_RELU_MASK | ops.EPI_MASK_BITS, colsum=cs_b)
    ops.sync_check()
    assert torch.equal(ref, got)
    assert (cs_a - cs_b).abs().max().item() <= 1e-3 * max(1.0, cs_a.abs().max().item())       # fp32 atomics: order differs
    full = (dz.float() @ w2.float()) * (hid > 0)
    assert (got.float() - full).abs().max().item() < 2e-2 * max(1.0, full.abs().max().item())


@pytest.mark.parametrize("T,F", [(1000, 2048), (128, 64), (129, 128), (257, 2048), (40000, 2048), (128 * 148 * 2 + 5, 512)])
def test_ffn_bwd_fused(T, F):
    """cb_ffn_bwd == autograd of linear2(relu(linear1(y))) + residual w.r.t. the hidden layer and y (chada_vit.py:113-116):
    dh = (dz2 W2) o (hidden > 0) rounded to bf16, dy = dh W1 + dz2 — against fp32 torch on the same bf16 operands and against
    the two cb_gemm_bf16 calls it replaces."""
    from chadavit_b200 import ops
    D = 192
    assert ops.ffn_bwd_fused_ok(D, F)
    y, w1, w2 = _rand((T, D), 31), _rand((F, D), 32, 0.08), _rand((D, F), 33, 0.05)
    b1, b2 = torch.randn(F, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
    y32 = y.float()
    _, hid, bits = ops.ffn_fwd(y, w1, b1, w2, b2, y32, save_hidden=True, save_mask_bits=True)
    dz2 = torch.randn(T, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(34)) * 0.3
    dz2h = dz2.to(torch.bfloat16)
    dy, dh = ops.ffn_bwd(dz2h, w2, w1, bits, dz2)
    ops.sync_check()
    ref_dh = ((dz2h.float() @ w2.float()) * (hid.float() > 0)).to(torch.bfloat16)
    ref_dy = ref_dh.float() @ w1.float() + dz2
    e_dh = (dh.float() - ref_dh.float()).abs().max().item()
    e_dy = (dy - ref_dy).abs().max().item()
    u_dh = ops.gemm(dz2h, w2, b_mn=True, aux=bits, flags=ops.EPI_RELU_MASK | ops.EPI_MASK_BITS)
    u_dy = ops.gemm(u_dh, w1, b_mn=True, aux=dz2, flags=ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32)
    ops.sync_check()
    d_dh = (dh.float() - u_dh.float()).abs().max().item()
    d_dy = (dy - u_dy).abs().max().item()
    print(f"ffn bwd T={T} F={F}: dh err {e_dh:.3e} (max {ref_dh.float().abs().max().item():.2f}), dy err {e_dy:.3e} (max {ref_dy.abs().max().item():.2f}); "
          f"vs unfused: dh {d_dh:.3e}, dy {d_dy:.3e}")
    assert e_dh <= 1.6e-2 * max(1.0, ref_dh.float().abs().max().item())      # one bf16 ulp of the largest element
    assert e_dy <= 3e-3 * max(1.0, ref_dy.abs().max().item())
    assert d_dh <= 1.6e-2 * max(1.0, ref_dh.float().abs().max().item()) and d_dy <= 3e-3 * max(1.0, ref_dy.abs().max().item())
    assert (dh.float() != 0).sum().item() == ((hid.float() > 0) & (dh.float() != 0)).sum().item()   # nothing leaks through the mask


@pytest.mark.parametrize("T,keep_z", [(512, True), (1000, False), (68664, True), (128 * 148 + 77, False)])
def test_gemm_ln_fused(T, keep_z):
    """cb_gemm_ln_fwd == out_proj + residual followed by LayerNorm (chada_vit.py:99, :105-111): against the two kernels it
    replaces (same bf16 operands, fp32 accumulation and two-pass statistics) and against torch in fp32."""
    import torch.nn.functional as F
    from chadavit_b200 import ops
    D = 192
    assert ops.gemm_ln_ok(T, D, D)
    att, w = _rand((T, D), 41), _rand((D, D), 42, 0.07)
    g = torch.Generator(device="cpu").manual_seed(43)
    bias, x = (torch.randn(D, generator=g) * 0.1).cuda(), (torch.randn(T, D, generator=g) * 1.5).cuda()
    gamma, beta = (1 + 0.2 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    z, y, y32, mean, rstd = ops.gemm_ln_fwd(att, w, bias, x, gamma, beta, 1e-5, keep_z=keep_z)
    ops.sync_check()
    z_ref = ops.gemm(att, w, bias=bias, aux=x, flags=ops.EPI_RESIDUAL_F32 | ops.EPI_OUT_F32)
    y_u, y32_u, mean_u, rstd_u = ops.layernorm_fwd(z_ref, gamma, beta, 1e-5, out_f32=True)
    ops.sync_check()
    assert (z is None) == (not keep_z)
    if keep_z:
        assert torch.equal(z, z_ref)                                    # the same accumulator, bias and residual arithmetic
    assert (y32 - y32_u).abs().max().item() < 2e-5 and (mean - mean_u).abs().max().item() < 1e-6
    assert ((rstd - rstd_u).abs() / rstd_u).max().item() < 1e-5
    assert (y.float() - y_u.float()).abs().max().item() <= 0.04        # one bf16 ulp where the fp32 values straddle a rounding boundary
    t = F.layer_norm(att.float() @ w.float().t() + bias + x, (D,), gamma, beta, 1e-5)
    assert (y32 - t).abs().max().item() < 2e-3
