"""Generate tests/golden/*.npz by running the REFERENCE's own modules (needs /root/reference).

    python tests/golden/make_golden.py

Inputs and weights are regenerated from oracle.det (integer-hash, RNG-free) by the tests, so
the fixtures hold only the reference's *outputs* (small).  Each case dict lists the exact
recipe so tests rebuild identical inputs.  Run in the build container only; the GPU box
never has the reference tree.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import det, ref_loader  # noqa: E402
from oracle import chada_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)

BACKBONE_CASES = {
    # name: dict(D, ctor, counts, hw, all_tokens, max_ch)
    "tiny_224_cls":   dict(D=32, ctor="factory", counts=[1, 3, 5, 10], hw=224, all_tokens=False, max_ch=10, seed=1),
    "tiny_224_all":   dict(D=32, ctor="factory", counts=[2, 1, 4], hw=224, all_tokens=True, max_ch=10, seed=2),
    "tiny_96_cls":    dict(D=32, ctor="factory", counts=[2, 10, 1], hw=96, all_tokens=False, max_ch=10, seed=3),
    "tiny_maxch3":    dict(D=32, ctor="factory", counts=[3, 1], hw=96, all_tokens=False, max_ch=3, seed=4),
    "moyen_224_cls":  dict(D=192, ctor="factory", counts=[3, 1, 6], hw=224, all_tokens=False, max_ch=10, seed=5),
    "moyen_h12_cls":  dict(D=192, ctor="bare", counts=[2, 5], hw=224, all_tokens=False, max_ch=10, seed=6),
}


def load_det(module: torch.nn.Module, salt: int) -> dict:
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = det.det_state_dict(shapes, salt)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def build_backbone(R, c):
    if c["ctor"] == "factory":
        return R.chada_vit(patch_size=16, embed_dim=c["D"], return_all_tokens=c["all_tokens"],
                           max_number_channels=c["max_ch"]), 2, 1e-6
    return R.ChAdaViT(patch_size=16, embed_dim=c["D"], return_all_tokens=c["all_tokens"],
                      max_number_channels=c["max_ch"]), 12, 1e-5


def main():
    R = ref_loader.load()
    out = {}
    meta = {"backbone": BACKBONE_CASES}
    # ---------------- backbone forward (+ grads for one case)
    for name, c in BACKBONE_CASES.items():
        m, nhead, eps = build_backbone(R, c)
        m.train()
        P = load_det(m, c["seed"])
        x = torch.from_numpy(det.det_pixels(sum(c["counts"]), c["hw"], c["hw"], c["seed"]))
        y = m(x, 0, [c["counts"]])
        out[f"bb.{name}.out"] = y.detach().numpy() if y.shape[0] <= 64 else y.detach()[::37].numpy()
        out[f"bb.{name}.out_sum"] = np.float64(y.detach().double().sum().item())
        out[f"bb.{name}.out_shape"] = np.array(y.shape)
        # oracle cross-check while the reference is at hand
        yo = O.backbone_forward(x, 0, [c["counts"]], P, nhead=nhead, final_eps=eps,
                                return_all_tokens=c["all_tokens"], max_channels_model=c["max_ch"])
        print(f"{name}: ref vs oracle max|d| = {(y - yo).abs().max().item():.3e}  shape {tuple(y.shape)}")
        if name in ("tiny_224_cls", "tiny_96_cls", "moyen_224_cls"):
            wgt = torch.from_numpy(det.det_uniform(tuple(y.shape), 99, 1.0))
            (y * wgt).sum().backward()
            for k, p in m.named_parameters():
                if p.grad is None:
                    continue
                g = p.grad.detach()
                out[f"bb.{name}.grad.{k}.sum"] = np.float64(g.double().sum().item())
                out[f"bb.{name}.grad.{k}.abs"] = np.float64(g.double().abs().sum().item())
                if k in ("cls_token", "channel_token", "token_learner.proj.bias", "norm.weight",
                         "blocks.0.norm1.weight", "blocks.11.norm2.bias", "blocks.5.self_attn.in_proj_bias"):
                    out[f"bb.{name}.grad.{k}"] = g.numpy().copy()
                elif k in ("pos_embed", "blocks.3.linear1.weight", "blocks.0.self_attn.in_proj_weight",
                           "blocks.7.self_attn.out_proj.weight", "blocks.11.linear2.weight",
                           "token_learner.proj.weight"):
                    out[f"bb.{name}.grad.{k}.sub"] = g.reshape(-1)[::61].numpy().copy()

    # ---------------- DINO head
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tag, (ind, K) in {"h32": (32, 4096), "h192": (192, 4096)}.items():
            head = R.DINOHead(in_dim=ind, num_prototypes=K, use_bn=False)
            Ph = load_det(head, 11)
            f = torch.from_numpy(det.det_uniform((6, ind), 21, 1.5)).requires_grad_()
            z = head(f)
            zo = O.dino_head(f.detach(), Ph)
            print(f"head {tag}: ref vs oracle {(z - zo).abs().max().item():.3e}")
            out[f"head.{tag}.out_sub"] = z.detach()[:, ::16].numpy().copy()
            out[f"head.{tag}.out_sum"] = np.float64(z.detach().double().sum().item())
            wgt = torch.from_numpy(det.det_uniform(tuple(z.shape), 98, 1.0))
            (z * wgt).sum().backward()
            out[f"head.{tag}.grad_in"] = f.grad.numpy().copy()
            for k, p in head.named_parameters():
                if p.grad is not None:
                    out[f"head.{tag}.grad.{k}.sum"] = np.float64(p.grad.double().sum().item())
                    out[f"head.{tag}.grad.{k}.abs"] = np.float64(p.grad.double().abs().sum().item())
                    out[f"head.{tag}.grad.{k}.sub"] = p.grad.reshape(-1)[::997].numpy().copy()

        # the class default: use_bn=True (BatchNorm1d after the first two Linear layers), train mode, then one eval-mode forward
        head = R.DINOHead(in_dim=32, num_prototypes=256)
        head.train()
        Ph = load_det(head, 12)
        f = torch.from_numpy(det.det_uniform((12, 32), 22, 1.5)).requires_grad_()
        z = head(f)
        zo = O.dino_head(f.detach(), Ph, bn_training=True)
        print(f"head bn: ref vs oracle {(z - zo).abs().max().item():.3e}")
        out["head.bn.out"] = z.detach().numpy().copy()
        wgt = torch.from_numpy(det.det_uniform(tuple(z.shape), 97, 1.0))
        (z * wgt).sum().backward()
        out["head.bn.grad_in"] = f.grad.numpy().copy()
        for k, p in head.named_parameters():
            if p.grad is not None:
                out[f"head.bn.grad.{k}"] = p.grad.numpy().copy() if p.grad.numel() <= 4096 else p.grad.reshape(-1)[::97].numpy().copy()
        for k, b in head.named_buffers():
            out[f"head.bn.buf.{k}"] = b.detach().numpy().copy()
        head.eval()
        with torch.no_grad():
            out["head.bn.out_eval"] = head(f.detach()).numpy().copy()

    # ---------------- DINO loss (+ center, + grad wrt student), V = 2 and V = 8, two consecutive calls
    for V in (2, 8):
        B, K = 5, 4096
        L = R.DINOLoss(num_prototypes=K, warmup_teacher_temp=0.04, teacher_temp=0.07,
                       warmup_teacher_temp_epochs=3, num_epochs=10, num_large_crops=V)
        L.epoch = 1
        for call in range(2):
            s = torch.from_numpy(det.det_uniform((V * B, K), 31 + call, 1.0)).requires_grad_()
            t = torch.from_numpy(det.det_uniform((2 * B, K), 41 + call, 1.0))
            loss = L(s, t)
            loss.backward()
            out[f"loss.V{V}.call{call}.loss"] = np.float64(loss.item())
            out[f"loss.V{V}.call{call}.center_sub"] = L.center[0, ::8].numpy().copy()
            out[f"loss.V{V}.call{call}.grad_sub"] = s.grad[:, ::64].numpy().copy()
            out[f"loss.V{V}.call{call}.grad_abs"] = np.float64(s.grad.double().abs().sum().item())
        out[f"loss.V{V}.temp_epoch1"] = np.float64(L.teacher_temp_schedule[1])

    # ---------------- EMA
    upd = R.MomentumUpdater(0.99, 1.0)
    a = torch.nn.Linear(7, 5); b = torch.nn.Linear(7, 5)
    with torch.no_grad():
        a.weight.copy_(torch.from_numpy(det.det_uniform((5, 7), 51))); a.bias.copy_(torch.from_numpy(det.det_uniform((5,), 52)))
        b.weight.copy_(torch.from_numpy(det.det_uniform((5, 7), 53))); b.bias.copy_(torch.from_numpy(det.det_uniform((5,), 54)))
    upd.update_tau(30, 100)
    out["ema.tau_30_100"] = np.float64(upd.cur_tau)
    upd.update(a, b)
    out["ema.weight"] = b.weight.detach().numpy().copy()
    out["ema.bias"] = b.bias.detach().numpy().copy()

    # ---------------- one DINO step with the reference wiring (tiny): loss, center, grads
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        counts = [1, 3, 2, 5]
        K = 4096
        stu, nhead, eps = build_backbone(R, dict(D=32, ctor="factory", all_tokens=False, max_ch=10))
        tea, _, _ = build_backbone(R, dict(D=32, ctor="factory", all_tokens=False, max_ch=10))
        sh = R.DINOHead(in_dim=32, num_prototypes=K, use_bn=False)
        th = R.DINOHead(in_dim=32, num_prototypes=K, use_bn=False)
        load_det(stu, 61); load_det(tea, 62); load_det(sh, 63); load_det(th, 64)
        L = R.DINOLoss(num_prototypes=K, warmup_teacher_temp=0.04, teacher_temp=0.07,
                       warmup_teacher_temp_epochs=0, num_epochs=10)
        crops = [torch.from_numpy(det.det_pixels(sum(counts), 224, 224, 71 + i)) for i in range(2)] + \
                [torch.from_numpy(det.det_pixels(sum(counts), 96, 96, 81 + i)) for i in range(2)]
        lnc = [counts] * 4
        z = [sh(stu(crops[i], i, lnc)) for i in range(2)]
        for i, xc in enumerate(crops[2:]):
            stu(xc, i, lnc)
        with torch.no_grad():
            mz = [th(tea(crops[i], i, lnc)) for i in range(2)]
        loss = L(torch.cat(z), torch.cat(mz))
        loss.backward()
        out["step.loss"] = np.float64(loss.item())
        out["step.center_sub"] = L.center[0, ::8].numpy().copy()
        for mod, tag in ((stu, "bb"), (sh, "head")):
            for k, p in mod.named_parameters():
                if p.grad is None:
                    continue
                out[f"step.grad.{tag}.{k}.sum"] = np.float64(p.grad.double().sum().item())
                out[f"step.grad.{tag}.{k}.abs"] = np.float64(p.grad.double().abs().sum().item())
        meta["step"] = dict(counts=counts, K=K, seeds=dict(stu=61, tea=62, sh=63, th=64, g=[71, 72], l=[81, 82]))

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1)
    sz = os.path.getsize(os.path.join(HERE, "reference_outputs.npz"))
    print(f"wrote {len(out)} arrays, {sz / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
