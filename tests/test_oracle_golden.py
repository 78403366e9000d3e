"""CPU: the oracle restatement reproduces the golden outputs of the REFERENCE's own modules (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import chada_oracle as O
from oracle import det
from tests.helpers import backbone_case, cases, det_params, golden

G = golden()
CASES = cases()["backbone"]


@pytest.mark.parametrize("name", ["tiny_224_cls", "tiny_224_all", "tiny_96_cls", "tiny_maxch3"])
def test_backbone_forward_matches_reference(name):
    c = CASES[name]
    P, x, nhead, eps = backbone_case(c)
    with torch.no_grad():
        y = O.backbone_forward(x, 0, [c["counts"]], P, nhead=nhead, final_eps=eps, return_all_tokens=c["all_tokens"],
                               max_channels_model=c["max_ch"])
    assert list(y.shape) == G[f"bb.{name}.out_shape"].tolist()
    ref = torch.from_numpy(G[f"bb.{name}.out"])
    got = y if y.shape[0] <= 64 else y[::37]
    assert (got - ref).abs().max().item() < 2e-5
    assert abs(y.double().sum().item() - float(G[f"bb.{name}.out_sum"])) < 1e-2


@pytest.mark.parametrize("name", ["tiny_96_cls", "moyen_224_cls"])
def test_backbone_gradients_match_reference(name):
    c = CASES[name]
    P, x, nhead, eps = backbone_case(c)
    P = {k: v.requires_grad_() for k, v in P.items()}
    y = O.backbone_forward(x, 0, [c["counts"]], P, nhead=nhead, final_eps=eps)
    (y * torch.from_numpy(det.det_uniform(tuple(y.shape), 99, 1.0))).sum().backward()
    for k in ("cls_token", "channel_token", "norm.weight", "blocks.0.norm1.weight"):
        ref = torch.from_numpy(G[f"bb.{name}.grad.{k}"])
        assert (P[k].grad - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())
    for k in ("pos_embed", "blocks.3.linear1.weight", "blocks.11.linear2.weight", "token_learner.proj.weight"):
        ref = torch.from_numpy(G[f"bb.{name}.grad.{k}.sub"])   # pos_embed: the gradient flows through the bicubic resize (96^2 case)
        got = P[k].grad.reshape(-1)[::61]
        assert (got - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item()), k
        assert ((got - ref).norm() / ref.norm()).item() < 1e-3, k
    for k, v in P.items():                                     # every parameter: sum |grad| within 2e-3 of the reference's
        ref = float(G[f"bb.{name}.grad.{k}.abs"])
        assert abs(v.grad.double().abs().sum().item() - ref) <= 2e-3 * max(ref, 1e-3), k


def test_head_loss_ema_match_reference():
    Ph = det_params(O.head_shapes(32, 4096), 11)
    f = torch.from_numpy(det.det_uniform((6, 32), 21, 1.5))
    z = O.dino_head(f, Ph)
    assert (z[:, ::16] - torch.from_numpy(G["head.h32.out_sub"])).abs().max().item() < 1e-5
    for V in (2, 8):
        center = torch.zeros(1, 4096)
        temp = O.teacher_temp(1, 0.04, 0.07, 3)
        assert abs(temp - float(G[f"loss.V{V}.temp_epoch1"])) < 1e-12
        for call in range(2):
            s = torch.from_numpy(det.det_uniform((V * 5, 4096), 31 + call, 1.0)).requires_grad_()
            t = torch.from_numpy(det.det_uniform((10, 4096), 41 + call, 1.0))
            loss, center = O.dino_loss(s, t, center, student_temp=0.1, teacher_temp=temp, num_large_crops=V)
            loss.backward()
            assert abs(loss.item() - float(G[f"loss.V{V}.call{call}.loss"])) < 1e-5
            assert (center[0, ::8] - torch.from_numpy(G[f"loss.V{V}.call{call}.center_sub"])).abs().max().item() < 1e-6
            assert (s.grad[:, ::64] - torch.from_numpy(G[f"loss.V{V}.call{call}.grad_sub"])).abs().max().item() < 1e-7
    tau = O.cosine_tau(0.99, 1.0, 30, 100)
    assert abs(tau - float(G["ema.tau_30_100"])) < 1e-12
    a_w, a_b = det.det_uniform((5, 7), 51), det.det_uniform((5,), 52)
    b_w, b_b = det.det_uniform((5, 7), 53), det.det_uniform((5,), 54)
    nw, nb = O.ema_update([torch.from_numpy(a_w), torch.from_numpy(a_b)], [torch.from_numpy(b_w), torch.from_numpy(b_b)], tau)
    assert np.allclose(nw.numpy(), G["ema.weight"], atol=1e-7) and np.allclose(nb.numpy(), G["ema.bias"], atol=1e-7)


def test_dino_step_matches_reference():
    st = cases()["step"]
    counts, K, sd = st["counts"], st["K"], st["seeds"]
    stu = det_params(O.backbone_shapes(32), sd["stu"]); tea = det_params(O.backbone_shapes(32), sd["tea"])
    sh = det_params(O.head_shapes(32, K), sd["sh"]); th = det_params(O.head_shapes(32, K), sd["th"])
    for d in (stu, sh):
        for v in d.values():
            v.requires_grad_()
    crops = [torch.from_numpy(det.det_pixels(sum(counts), 224, 224, s)) for s in sd["g"]] + \
            [torch.from_numpy(det.det_pixels(sum(counts), 96, 96, s)) for s in sd["l"]]
    loss, center = O.dino_step(crops, [counts] * 4, stu, sh, tea, th, torch.zeros(1, K), nhead=2, final_eps=1e-6,
                               teacher_temp=0.07, run_local_crops=False)
    loss.backward()
    assert abs(loss.item() - float(G["step.loss"])) < 2e-5
    assert (center[0, ::8] - torch.from_numpy(G["step.center_sub"])).abs().max().item() < 1e-6
    for tag, d in (("bb", stu), ("head", sh)):
        for k, v in d.items():
            key = f"step.grad.{tag}.{k}.abs"
            if key in G.files and v.grad is not None:
                ref = float(G[key])
                assert abs(v.grad.double().abs().sum().item() - ref) <= 2e-3 * max(ref, 1e-3), k


def test_head_with_batchnorm_matches_reference():
    """DINOHead(use_bn=True) — the class default (src/methods/dino.py:40,66-73): train-mode forward, input / parameter gradients,
    running statistics after the call, then an eval-mode forward, all against the reference module's golden outputs."""
    P = det_params(O.head_shapes(32, 256, use_bn=True), 12)
    for k in P:
        if k.endswith("num_batches_tracked"):
            P[k] = P[k].to(torch.int64)
        elif not ("running" in k or k.endswith("weight_g")):
            P[k].requires_grad_()
    f = torch.from_numpy(det.det_uniform((12, 32), 22, 1.5)).requires_grad_()
    z = O.dino_head(f, P, bn_training=True)
    assert (z.detach() - torch.from_numpy(G["head.bn.out"])).abs().max().item() < 1e-5
    (z * torch.from_numpy(det.det_uniform(tuple(z.shape), 97, 1.0))).sum().backward()
    assert (f.grad - torch.from_numpy(G["head.bn.grad_in"])).abs().max().item() < 1e-5
    for k in ("mlp.1.weight", "mlp.1.bias", "mlp.4.weight", "mlp.6.bias"):
        ref = torch.from_numpy(G[f"head.bn.grad.{k}"])
        assert (P[k].grad - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item()), k
    for k in ("mlp.1.running_mean", "mlp.1.running_var", "mlp.4.running_var"):
        assert (P[k] - torch.from_numpy(G[f"head.bn.buf.{k}"])).abs().max().item() < 1e-6, k
    assert int(P["mlp.1.num_batches_tracked"]) == int(G["head.bn.buf.mlp.1.num_batches_tracked"]) == 1
    with torch.no_grad():
        ze = O.dino_head(f.detach(), P, bn_training=False)
    assert (ze - torch.from_numpy(G["head.bn.out_eval"])).abs().max().item() < 1e-5
