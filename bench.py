#!/usr/bin/env python
"""Benchmark of the hot path: one ChAda-ViT-moyen/16 DINO pre-training step (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = student forward on 2 global (224^2) + 6 local (96^2) crops of 64 images per GPU with a ragged 1..10 channel
mix, teacher forward on the global crops, DINO head + fused loss, backward, gradient all-reduce (N > 1), fused AdamW,
teacher EMA and centre update — the reference's wiring (local crops go through the student backbone only).
Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same step through the public
API from pinned HOST buffers (H2D of every crop + D2H of the loss inside the timed region).  `roofline` = the dominant
kernel class timed live with CUDA events; `cpu_baseline` = the oracle port on the host cores on a bounded sample.
`--impl reference` times that CPU port as its own arm (the reference is pure Python/PyTorch; /root/reference does not
exist on the GPU box, so the arm runs oracle/chada_oracle.py, which is pinned to the reference by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_MODEL, N_PROTO, BATCH, N_GLOBAL, N_LOCAL = 192, 4096, 64, 2, 6
METRIC = "DINO pretrain imgs/s at 1/2/4/8 B200; varlen attn TFLOPS vs BF16 peak"
FFN3_TRAFFIC = 100.8e6     # dram bytes of one ffn_fwd3_kernel launch (ncu --set full, profiles/r01_ncu_ffn_fwd3.txt)


def channel_counts(batch: int, seed: int = 1234):
    """C_b ~ U{1..10} (HOW_TO_USE.ipynb cell 16), fixed seed, identical on every rank -> ranks are token-balanced."""
    return np.random.RandomState(seed).randint(1, 11, size=batch).tolist()


def make_crops(counts, seed, device, pin=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    G = sum(counts)
    crops = []
    for i in range(N_GLOBAL + N_LOCAL):
        hw = 224 if i < N_GLOBAL else 96
        x = torch.randn(G, 1, hw, hw, generator=g)
        crops.append(x.pin_memory() if pin else x.to(device))
    return crops


def dino_cfg(multicrop=False, graph=True):
    return {"method": "dino", "backbone": {"kwargs": {"patch_size": 16, "embed_dim": D_MODEL, "return_all_tokens": False}},
            "data": {"max_img_channels": 10, "num_large_crops": N_GLOBAL, "num_small_crops": N_LOCAL},
            "method_kwargs": {"num_prototypes": N_PROTO, "multicrop_loss": multicrop}, "max_epochs": 100, "max_steps": 100000,
            "optimizer": {"lr": 5e-4, "weight_decay": 1e-4}, "momentum": {"base_tau": 0.9995, "final_tau": 1.0},
            "engine": {"cuda_graph": graph}}


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int, enabled: bool = True):
        """enabled = False (ranks > 0 of a multi-GPU run): no polling — eight concurrent nvidia-smi pollers would contend for the
        driver and the host cores inside the timed region; rank 0's GPU stands for the box (the line is printed by rank 0)."""
        self.index, self.samples, self.stop, self.enabled = index, [], threading.Event(), enabled
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        if self.enabled:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.enabled:
            self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_port_step(n_images: int, threads: int, reps: int, warmup: int = 1):
    """The reference's algorithm (oracle port, fp32, padded + key-padding mask) for the same step on `n_images` images."""
    from oracle import chada_oracle as O
    from oracle import det
    torch.set_num_threads(threads)
    counts = channel_counts(BATCH)[:n_images]
    stu = {k: torch.from_numpy(v).requires_grad_() for k, v in det.det_state_dict(O.backbone_shapes(D_MODEL), 1).items()}
    tea = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.backbone_shapes(D_MODEL), 2).items()}
    sh = {k: torch.from_numpy(v).requires_grad_() for k, v in det.det_state_dict(O.head_shapes(D_MODEL, N_PROTO), 3).items()}
    th = {k: torch.from_numpy(v) for k, v in det.det_state_dict(O.head_shapes(D_MODEL, N_PROTO), 4).items()}
    sh["last_layer.weight_g"].requires_grad_(False)
    params = [p for p in list(stu.values()) + list(sh.values()) if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=1e-4)
    crops = make_crops(counts, 99, "cpu")
    center = torch.zeros(1, N_PROTO)
    times = []
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, center = O.dino_step(crops, [counts] * (N_GLOBAL + N_LOCAL), stu, sh, tea, th, center, nhead=2, final_eps=1e-6,
                                   num_large_crops=N_GLOBAL, run_local_crops=True)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for d_on, d_mo in ((stu, tea), (sh, th)):
                for k in d_on:
                    d_mo[k].mul_(0.9995).add_(d_on[k].detach(), alpha=0.0005)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return n_images / float(np.median(times)), float(np.median(times))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_img = 4
    t0 = time.perf_counter()
    ips, sec = cpu_port_step(n_img, cores, reps=max(1, args.steps), warmup=min(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "imgs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ChAda-ViT-moyen/16 DINO pretraining step (2x224^2 + 6x96^2 crops, ragged 1-10 channels), "
                                   f"CPU sample of {n_img} images per step", "global_batch": n_img},
            "cpu_baseline": {"value": ips, "unit": "imgs/s", "cores": cores, "kind": "port",
                             "sample": f"{n_img} images x 8 crops per step, oracle port of the reference modules (fp32, padded+masked), "
                                       f"fwd+bwd+AdamW+EMA, median of {max(1, args.steps)} steps"},
            "e2e": {"value": ips, "unit": "imgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multicrop", action="store_true", help="true multi-crop loss (V=8) instead of the reference wiring")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from chadavit_b200 import _lib, ops
    from chadavit_b200.methods import DINO

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    torch.manual_seed(0)                      # identical random init on every rank (DDP replicas)
    model = DINO(dino_cfg(args.multicrop, graph=not args.no_graph)).to(dev)
    counts = channel_counts(BATCH)            # same multiset on every rank -> token-balanced
    lnc = [counts] * (N_GLOBAL + N_LOCAL)
    crops = make_crops(counts, 1234 + rank, dev)
    batch = (crops, None, lnc)
    tokens_g = sum(1 + c * 196 for c in counts)
    tokens_l = sum(1 + c * 36 for c in counts)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        loss = model.fused_train_step(batch)
    sync()
    # ---- timed region 1: inputs resident in HBM
    launches_eager0 = _lib.launch_count
    model.use_cuda_graph, keep = False, model.use_cuda_graph
    model.fused_train_step(batch)             # one eager step only to COUNT the kernels a step launches (graph replays bypass Python)
    model.use_cuda_graph = keep
    launches_per_step = _lib.launch_count - launches_eager0
    sync()
    with ClockSampler(local_rank, enabled=(rank == 0)) as clk:
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = model.fused_train_step(batch)
        e1.record()
        sync()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    # ---- instrumented eager pass of the same step: CUDA-event duration of every attention / GEMM launch (roofline)
    model.use_cuda_graph, keep = False, model.use_cuda_graph
    ops.PROFILE = {"cb_attn_varlen_fwd": [0, 0.0, [], 0.0], "cb_attn_varlen_bwd": [0, 0.0, [], 0.0], "cb_gemm_bf16": [0, 0.0, [], 0.0],
                   "cb_ffn_fwd": [0, 0.0, [], 0.0], "cb_ffn_fwd:nostore": [0, 0.0, [], 0.0]}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        model.fused_train_step(batch)
    ev1.record()
    sync()
    ms_eager = ev0.elapsed_time(ev1)
    model.use_cuda_graph = keep
    prof = {}
    for name, (n, work, evs, nbytes) in ops.PROFILE.items():
        t = sum(a.elapsed_time(b) for a, b in evs)
        prof[name] = {"launches": n, "ms_total": t, "work": work, "bytes": nbytes}
    ops.PROFILE = None
    loss_val = float(loss.item())

    # ---- timed region 2: end to end through the public API from pinned host buffers
    host_crops = make_crops(counts, 1234 + rank, dev, pin=True)
    h2d = sum(c.numel() * 4 for c in host_crops)
    host_batch = (host_crops, None, lnc)
    for _ in range(2):
        model.fused_train_step(model.stage_batch(host_batch)).item()
    sync()
    # Every step's crops are copied from pinned host memory inside the timed region (K copies for K steps) and every step's
    # loss is read back; DINO.stage_batch issues the copy of batch i+1 on the engine's copy stream right after step i has
    # been launched, so it runs under that step's compute (the first copy is exposed).
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    nxt = model.stage_batch(host_batch)
    for i in range(args.steps):
        l = model.fused_train_step(nxt)        # waits for its H2D on the device, then the step
        if i + 1 < args.steps:
            nxt = model.stage_batch(host_batch)
        _ = l.item()                           # D2H read of the step's loss
    e3.record()
    sync()
    ms_e2e = e2.elapsed_time(e3)
    # the same loop without the prefetch (copy, then step, then read-back: fully serial), for reference
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in range(max(2, args.steps // 4)):
        _ = model.fused_train_step(host_batch).item()
    e5.record()
    sync()
    ms_e2e_serial = e4.elapsed_time(e5) / max(2, args.steps // 4)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        finish(world)
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)   # kernels timed inside a long step -> sustained figure
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs" if peaks else "fallback 1.4 PFLOP/s, 6.65 TB/s (B200_PROFILING.md)"
    roof = {}
    for name, p in prof.items():
        sec = p["ms_total"] * 1e-3
        ach = p["work"] / sec / 1e12 if sec > 0 else 0.0
        gbs = p["bytes"] / sec / 1e9 if sec > 0 else 0.0
        roof[name] = {"launches_per_step": p["launches"] / args.steps, "ms_per_step": p["ms_total"] / args.steps,
                      "share_of_step": p["ms_total"] / ms_eager, "achieved_tflops": ach, "frac": ach / tf_peak,
                      "achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak}
    # The dominant SINGLE kernel of the step is picked live among the single-kernel classes (profiles/r01_launches_bench_step.txt:
    # attention backward, attention forward, the two fused feed-forward kernels; the GEMM class is larger in total but is 14
    # instantiations over a dozen shapes, most of them HBM-bound: see "all" for its aggregate TFLOP/s and GB/s).
    # `traffic` = dram read + write bytes of ONE launch from the committed `ncu --set full` capture of that kernel on the
    # ragged global-crop batch of 64 images (T = 68 664 tokens; the bench launches it on 2x that for the packed global crops).
    NCU = {"cb_attn_varlen_bwd": ("attn_bwd_kernel<96> (cb_attn_varlen_bwd)", 223.1e6, "profiles/r01_ncu_attn_bwd.txt",
                                  "capture: T = 68664, H = 2, d = 96; algorithmic bytes of that launch 237.3e6"),
           "cb_attn_varlen_fwd": ("attn_fwd2_kernel<96> (cb_attn_varlen_fwd)", 87.5e6, "profiles/r01_ncu_attn_fwd.txt",
                                  "capture: T = 68664, H = 2, d = 96"),
           "cb_ffn_fwd": ("ffn_fwd_kernel (cb_ffn_fwd, hidden activations stored: student pass)", 366.2e6, "profiles/r01_ncu_ffn_fwd.txt",
                          "capture: T = 68664; algorithmic bytes of that launch 414.7e6 (the tail of the hidden store is still in L2 "
                          "when the kernel ends)"),
           "cb_ffn_fwd:nostore": ("ffn_fwd3_kernel (cb_ffn_fwd, hidden activations not stored: teacher / local crops)", FFN3_TRAFFIC,
                                  "profiles/r01_ncu_ffn_fwd3.txt", "capture: T = 68664; algorithmic bytes of that launch 133.4e6")}
    top = max(NCU, key=lambda k: prof[k]["ms_total"])
    tp = prof[top]
    roofline = {"kernel": NCU[top][0], "bound": "tensor", "achieved": roof[top]["achieved_tflops"], "peak": tf_peak,
                "unit": "TFLOP/s", "frac": roof[top]["frac"], "traffic": NCU[top][1], "traffic_source": NCU[top][2], "traffic_note": NCU[top][3],
                "peak_source": peak_src, "algorithmic_flops_per_launch": tp["work"] / max(1, tp["launches"]),
                "algorithmic_bytes_per_launch": tp["bytes"] / max(1, tp["launches"]),
                "avg_launch_ms": tp["ms_total"] / max(1, tp["launches"]), "share_of_step": roof[top]["share_of_step"], "all": roof,
                "hbm_note": "a one-directional HBM stream tops out near 3.9 TB/s (write) / 4.3 TB/s (read) on this part; only mixed "
                            "traffic reaches the 6.55 TB/s copy figure (tools/membw.py, profiles/r01_hw_probes.txt)",
                "timing": "CUDA events around every launch of the class in an instrumented eager pass of the same step "
                          f"({ms_eager / args.steps:.1f} ms/step eager vs {ms / args.steps:.1f} ms/step timed)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only
        cores = os.cpu_count() or 1
        n_img = 8
        ips, sec = cpu_port_step(n_img, cores, reps=1, warmup=0)
        cpu = {"value": ips, "unit": "imgs/s", "cores": cores, "kind": "port",
               "sample": f"{n_img} images x 8 crops, one step (fwd+bwd+AdamW+EMA) of the oracle port of the reference modules, fp32, {sec:.1f} s"}

    imgs = BATCH * world
    line = {
        "metric": METRIC, "value": imgs * args.steps / (ms * 1e-3), "unit": "imgs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ChAda-ViT-moyen/16 (D=192, 12 blocks, 2 heads, FFN 2048) DINO pretraining step: 2x224^2 global + "
                               "6x96^2 local crops, batch 64/GPU, ragged U{1..10} channels, K=4096 prototypes, AdamW + teacher EMA; "
                               + ("true multi-crop loss (V=8)" if args.multicrop else "reference wiring (local crops: student backbone only, SURVEY Q11)"),
                   "global_batch": imgs, "per_gpu_batch": BATCH, "sum_channels_per_gpu": int(sum(counts)),
                   "tokens_per_gpu_global_crop": tokens_g, "tokens_per_gpu_local_crop": tokens_l,
                   "parallelism": f"dp{world}", "ranks_token_balanced": True, "cuda_graph": bool(model.use_cuda_graph),
                   "l2_policy": "working set per step (~10 GB of activations) >> 126 MB L2; no explicit flush"},
        "clocks": clk.summary(),
        "e2e": {"value": imgs * args.steps / (ms_e2e * 1e-3), "unit": "imgs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "ms_per_step_serial_copy": ms_e2e_serial,
                "pipeline": "H2D of batch i+1 on a copy stream under the compute of batch i (DINO.stage_batch), 2 device buffer sets"},
        "gpu_launches": launches,
        "roofline": roofline,
        "attn_tflops": {"fwd": roof["cb_attn_varlen_fwd"]["achieved_tflops"], "bwd": roof["cb_attn_varlen_bwd"]["achieved_tflops"],
                        "fwd_frac_of_peak": roof["cb_attn_varlen_fwd"]["frac"], "bwd_frac_of_peak": roof["cb_attn_varlen_bwd"]["frac"],
                        "peak_tflops": tf_peak, "flops_model": "fwd 4*D*sum(S_b^2), bwd 10*D*sum(S_b^2) per layer call, packed real tokens only"},
        "cpu_baseline": cpu,
        "loss": loss_val,
    }
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world: int):
    """Leave without tearing anything down.  With N > 1 the interpreter / NCCL / CUDA-graph teardown after the result line has
    been seen to hang (one rank in destroy_process_group while the other already left); the run is over once rank 0 has
    printed, so every rank meets at one last barrier (bounded by a watchdog) and exits hard."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
    os._exit(0)


if __name__ == "__main__":
    main()
