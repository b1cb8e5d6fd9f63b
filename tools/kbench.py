#!/usr/bin/env python
"""Per-kernel microbenchmarks through the public API (CUDA events on the launching stream).

    python tools/kbench.py mnn [--shapes 64x1024x1024x256,...] [--precisions tf32x3,bf16,fp32]
    python tools/kbench.py stages [--config c2_ec_superpoint] [--batch 64]

Development tool: prints one line per measurement to stdout; bench.py remains the judged benchmark.
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import einx  # noqa: E402

synth = importlib.import_module("ei-nexus_official_b200.synth")
DEV = torch.device("cuda", 0)


def time_ms(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_mnn(args):
    ctx = einx.context_for(DEV)
    for shape in args.shapes.split(","):
        B, N, M, D = (int(v) for v in shape.split("x"))
        g = torch.Generator(device=DEV).manual_seed(1)
        d0 = torch.nn.functional.normalize(torch.randn((B, N, D), device=DEV, generator=g), dim=-1)
        d1 = torch.nn.functional.normalize(torch.randn((B, M, D), device=DEV, generator=g), dim=-1)
        flops = 2.0 * B * N * M * D
        for prec in args.precisions.split(","):
            total = time_ms(lambda: einx.mnn(d0, d1, precision=prec))
            ctx.profile(True)
            ks = []
            for _ in range(5):
                einx.mnn(d0, d1, precision=prec)
                ks.append(ctx.profile_read()[3])
            ctx.profile(False)
            k = float(np.median(ks))
            print(f"mnn {shape} {prec}: entry {total:.4f} ms, similarity kernel {k:.4f} ms = {flops / k / 1e9:.1f} TFLOP/s algorithmic", flush=True)


def bench_next(args):
    """Rows adjacent to the path (SURVEY.md section 8 f) at the sizes of a config."""
    c = synth.CONFIGS[args.config]
    B = args.batch
    H, W, cell = c["H"], c["W"], c["cell"]
    Hp, Wp, _ = synth.padded_size(H, W, cell)
    rng = np.random.default_rng(0)
    evs = [synth.events(rng, c["events"], H, W, c["style"], c["dt"]) for _ in range(B)]
    x, y, t, p, off = (a.to(DEV) for a in einx.pack_events(evs))
    nev = x.numel()
    ms = time_ms(lambda: einx.events_image_device(x, y, off, H, W))
    print(f"events_image {B}x{c['events']} ev -> {H}x{W}: {ms:.4f} ms = {(8 * nev + B * H * W) / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    img = einx.events_image_device(x, y, off, H, W)
    ms = time_ms(lambda: einx.events_mask(img, cell))
    print(f"events_mask -> {Hp}x{Wp}: {ms:.4f} ms = {(B * H * W + B * Hp * Wp) / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    if cell > 1:
        C = cell * cell + 1
        lo = torch.randn((B, C, Hp // cell, Wp // cell), device=DEV)
        ms = time_ms(lambda: einx.logits_to_score(lo, cell))
        print(f"logits_to_score {B}x{C}x{Hp // cell}x{Wp // cell}: {ms:.4f} ms = {(lo.numel() * 4 + B * Hp * Wp * 4) / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    K = c["top_k"]
    sc = torch.randn((B, K + 1, K + 1), device=DEV) - 5
    ms = time_ms(lambda: einx.filter_matches(sc, 0.1))
    print(f"filter_matches {B}x{K + 1}x{K + 1}: {ms:.4f} ms = {sc.numel() * 4 / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    sim = 4 * torch.randn((B, K, K), device=DEV)
    z0, z1 = torch.randn((B, K, 1), device=DEV), torch.randn((B, K, 1), device=DEV)
    for mb in ("", "16", "32", "64") if os.environ.get("KBENCH_LDS_SWEEP") else ("",):
        os.environ["EINX_LDS_CHUNK_MB"] = mb
        ms = time_ms(lambda: einx.sigmoid_log_double_softmax(sim, z0, z1))
        print(f"sigmoid_log_double_softmax {B}x{K}x{K}, {mb + ' MB of similarities per chunk' if mb else 'whole batch per launch'}: "
              f"{ms:.4f} ms = {(2 * sim.numel() * 4 + sc.numel() * 4) / ms / 1e6:.0f} GB/s (sim read twice + matrix written once)", flush=True)
    os.environ.pop("EINX_LDS_CHUNK_MB")
    ms = time_ms(lambda: einx.sigmoid_log_double_softmax(sim, z0, z1, carry_best=False))
    print(f"  without the carried row / column maxima: {ms:.4f} ms", flush=True)
    ms = time_ms(lambda: einx.filter_matches(einx.sigmoid_log_double_softmax(sim, z0, z1), 0.1))
    ms2 = time_ms(lambda: einx.filter_matches(einx.sigmoid_log_double_softmax(sim, z0, z1, carry_best=False), 0.1))
    print(f"  sigmoid_log_double_softmax + filter_matches: {ms:.4f} ms with carried maxima, {ms2:.4f} ms as two passes", flush=True)
    ref = lambda: (torch.log_softmax(sim, 2) + torch.log_softmax(sim, 1) + torch.nn.functional.logsigmoid(z0)
                   + torch.nn.functional.logsigmoid(z1).transpose(1, 2))
    ms = time_ms(ref)
    print(f"  torch (log_softmax over rows + over columns + certainties, interior only, for scale): {ms:.4f} ms", flush=True)


def bench_stages(args):
    ctx = einx.context_for(DEV)
    c = synth.CONFIGS[args.config]
    B = args.batch
    Hp, Wp, _ = synth.padded_size(c["H"], c["W"], c["cell"])
    det, desc, mt = (importlib.import_module(f"ei-nexus_official_b200.{m}") for m in ("detection", "describe", "match"))
    evs, s0, r0, s1, r1 = [], [], [], [], []
    for i in range(B):
        ev, sides = synth.pair_inputs(args.config, i, None)
        evs.append(ev)
        s0.append(sides[0][0]); r0.append(sides[0][1]); s1.append(sides[1][0]); r1.append(sides[1][1])
    ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
    s0, r0, s1, r1 = (torch.from_numpy(np.concatenate(a)).to(DEV) for a in (s0, r0, s1, r1))
    cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"],
                          descriptor_scale=c["scale"], precision=args.precision)
    pipe = einx.ExtractMatchPipeline(cfg)
    mode = desc.BILINEAR if cfg.descriptor_mode == "bilinear" else desc.GATHER
    sc = s0.clone()
    _, kp0, cn0 = det.detect(sc, cfg.detection_threshold, cfg.nms_radius, cfg.remove_borders, cfg.top_k, kcap=cfg.top_k)
    sc1 = s1.clone()
    _, kp1, cn1 = det.detect(sc1, cfg.detection_threshold, cfg.nms_radius, cfg.remove_borders, cfg.top_k, kcap=cfg.top_k)
    d0 = desc.sample(r0, kp0, cn0, mode, (Hp, Wp), cfg.descriptor_scale, True)
    d1 = desc.sample(r1, kp1, cn1, mode, (Hp, Wp), cfg.descriptor_scale, True)
    print(f"{args.config} B={B}: keypoints/side mean {cn0.float().mean().item():.0f}", flush=True)
    stages = {
        "voxel": (0, lambda: pipe.voxelize(*ev)),
        "detect": (1, lambda: det.detect(s0, cfg.detection_threshold, cfg.nms_radius, cfg.remove_borders, cfg.top_k, kcap=cfg.top_k)),
        "sample": (2, lambda: desc.sample(r0, kp0, cn0, mode, (Hp, Wp), cfg.descriptor_scale, True)),
        "mnn": (3, lambda: mt.mnn(d0, d1, cn0, cn1, kp0, kp1, None, None, True, cfg.precision)),
    }
    for name, (slot, fn) in stages.items():
        total = time_ms(fn)
        ctx.profile(True)
        ks = []
        for _ in range(5):
            fn()
            ks.append(ctx.profile_read()[slot])
        ctx.profile(False)
        print(f"{name}: entry point {total:.4f} ms, dominant kernel {float(np.median(ks)):.4f} ms", flush=True)
    import dataclasses
    for conc in (False, True):
        pp = einx.ExtractMatchPipeline(dataclasses.replace(cfg, concurrent=conc))
        total = time_ms(lambda: pp(ev, s0, r0, s1, r1), iters=50)
        print(f"pipeline ({'3 streams' if conc else 'serial'}): {total:.4f} ms/step = {B / total * 1e3:.0f} pairs/s", flush=True)
        step = pp.capture(ev, s0, r0, s1, r1)
        total = time_ms(step.replay, iters=50)
        print(f"pipeline ({'3 streams' if conc else 'serial'}, CUDA graph): {total:.4f} ms/step = {B / total * 1e3:.0f} pairs/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["mnn", "stages", "next"])
    ap.add_argument("--shapes", default="64x1024x1024x256,32x2048x2048x128,1x8192x8192x128")
    ap.add_argument("--precisions", default="tf32x3,bf16")
    ap.add_argument("--config", default="c2_ec_superpoint")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="tf32x3")
    a = ap.parse_args()
    {"mnn": bench_mnn, "stages": bench_stages, "next": bench_next}[a.what](a)
