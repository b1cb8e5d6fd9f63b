"""Whole-path orchestration on the GPU: the three-stream fork/join, per-stream contexts and the
host-streaming driver must give exactly what the serial single-stream path gives."""
import dataclasses
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda", 0)
KEYS = ("voxel_grid", "keypoints0", "keypoints1", "counts0", "counts1", "descriptors0", "descriptors1", "matches0",
        "matches1", "matching_scores0", "num_matches", "matched_kpts0", "matched_kpts1")


@pytest.fixture(scope="module")
def einx():
    import einx as m

    m.context_for(DEV)
    return m


@pytest.fixture(scope="module")
def batch(einx):
    synth = importlib.import_module("ei-nexus_official_b200.synth")
    cfgname, B = "c2_ec_superpoint", 6
    c = synth.CONFIGS[cfgname]
    evs, s0, r0, s1, r1 = [], [], [], [], []
    for i in range(B):
        ev, sides = synth.pair_inputs(cfgname, 100 + i, 20_000 + 1000 * i)  # ragged windows
        evs.append(ev)
        s0.append(sides[0][0]); r0.append(sides[0][1]); s1.append(sides[1][0]); r1.append(sides[1][1])
    maps = [np.concatenate(a) for a in (s0, r0, s1, r1)]
    cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"],
                          descriptor_scale=c["scale"], precision="tf32x3")
    return cfg, evs, maps


def run(einx, cfg, evs, maps):
    pipe = einx.ExtractMatchPipeline(cfg)
    ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
    out = pipe(ev, *(torch.from_numpy(m.copy()).to(DEV) for m in maps))
    torch.cuda.synchronize()
    return {k: out[k].cpu() for k in KEYS}


def valid_rows_equal(a, b, counts):
    return all(torch.equal(a[i, :n], b[i, :n]) for i, n in enumerate(counts.tolist()))


def test_three_streams_match_serial(einx, batch):
    cfg, evs, maps = batch
    serial = run(einx, dataclasses.replace(cfg, concurrent=False), evs, maps)
    for _ in range(3):  # repeated: a cross-stream race would not be deterministic
        conc = run(einx, dataclasses.replace(cfg, concurrent=True), evs, maps)
        # the voxel scatter accumulates with atomics: order-dependent to 1e-5 like the reference itself
        assert torch.allclose(conc["voxel_grid"], serial["voxel_grid"], rtol=1e-5, atol=1e-5)
        assert torch.equal(conc["counts0"], serial["counts0"]) and torch.equal(conc["counts1"], serial["counts1"])
        for side in "01":
            n = serial[f"counts{side}"]
            assert valid_rows_equal(conc[f"keypoints{side}"], serial[f"keypoints{side}"], n)
            assert valid_rows_equal(conc[f"descriptors{side}"], serial[f"descriptors{side}"], n)
        assert valid_rows_equal(conc["matches0"], serial["matches0"], serial["counts0"])
        assert torch.equal(conc["num_matches"], serial["num_matches"])
        assert valid_rows_equal(conc["matched_kpts0"], serial["matched_kpts0"], serial["num_matches"])
        assert valid_rows_equal(conc["matched_kpts1"], serial["matched_kpts1"], serial["num_matches"])


def test_one_context_per_stream(einx):
    main_ctx = einx.context_for(DEV)
    s = torch.cuda.Stream(DEV)
    with torch.cuda.stream(s):
        side_ctx = einx.context_for(DEV)
        assert einx.context_for(DEV) is side_ctx
    assert side_ctx is not main_ctx and einx.context_for(DEV) is main_ctx
    assert {id(main_ctx), id(side_ctx)} <= {id(c) for c in einx.contexts_of(DEV)}
    assert einx.launch_count(DEV) >= main_ctx.launches


@pytest.mark.parametrize("chunks,compact", [(1, True), (4, True), (2, False)])
def test_host_streamer_matches_device_path(einx, batch, chunks, compact):
    """Sub-batch streaming from pinned host memory, with the 13-byte wire format for integer-pixel events
    (x, y uint16, p int8, t fp64) and with the plain 20-byte one, must reproduce the device-resident path."""
    cfg, evs, maps = batch
    ref = run(einx, cfg, evs, maps)
    B, K = len(evs), cfg.top_k
    hb = einx.HostBatch(evs, *maps, chunks=chunks, compact_events=compact)
    assert len(hb.chunks) == chunks and hb.batch == B and hb.compact == compact  # EC-style events are integral
    per_event = 13 if compact else 20
    assert hb.nbytes == sum(per_event * len(e["t"]) for e in evs) + 8 * (B + chunks) + sum(m.nbytes for m in maps)
    sub = dict(evs[0])
    sub["x"] = sub["x"] + 0.25  # sub-pixel coordinates: the compact format would lose them
    assert not einx.HostBatch([sub] + list(evs[1:]), *maps, chunks=1).compact
    out_host = {"matches0": torch.full((B, K), -7, dtype=torch.int64).pin_memory(),
                "num_matches": torch.zeros((B,), dtype=torch.int32).pin_memory(),
                "matched_kpts0": torch.zeros((B, K, 3)).pin_memory(),
                "matched_kpts1": torch.zeros((B, K, 3)).pin_memory()}
    streamer = einx.HostStreamer(einx.ExtractMatchPipeline(cfg), DEV)
    for _ in range(2):  # second pass reuses the staging buffers
        streamer.run(hb, out_host)
    torch.cuda.synchronize()
    assert torch.equal(out_host["num_matches"], ref["num_matches"])
    assert valid_rows_equal(out_host["matches0"], ref["matches0"], ref["counts0"])
    assert valid_rows_equal(out_host["matched_kpts0"], ref["matched_kpts0"], ref["num_matches"])
    assert valid_rows_equal(out_host["matched_kpts1"], ref["matched_kpts1"], ref["num_matches"])


def test_captured_step_replays_exactly(einx, batch):
    cfg, evs, maps = batch
    ref = run(einx, cfg, evs, maps)
    pipe = einx.ExtractMatchPipeline(cfg)
    ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
    dmaps = [torch.from_numpy(m.copy()).to(DEV) for m in maps]
    step = pipe.capture(ev, *dmaps)
    for _ in range(3):
        out = step.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["counts0"].cpu(), ref["counts0"]) and torch.equal(out["num_matches"].cpu(), ref["num_matches"])
    assert valid_rows_equal(out["matches0"].cpu(), ref["matches0"], ref["counts0"])
    assert valid_rows_equal(out["keypoints1"].cpu(), ref["keypoints1"], ref["counts1"])
    assert valid_rows_equal(out["matched_kpts1"].cpu(), ref["matched_kpts1"], ref["num_matches"])
    # refreshed inputs, same buffers: replay follows the data
    dmaps[0].copy_(torch.from_numpy(maps[2]))   # side 0 now sees side 1's score map
    dmaps[1].copy_(torch.from_numpy(maps[3]))
    out = step.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["counts0"].cpu(), ref["counts1"])
    assert valid_rows_equal(out["keypoints0"].cpu(), ref["keypoints1"], ref["counts1"])


def test_torch_ops_trace_and_raise(einx):
    synth = importlib.import_module("ei-nexus_official_b200.synth")
    """torch.ops.einx.*: the registered operators behind the host layer.  They trace under torch.compile (FakeTensor
    through the Meta kernels, mutation of the score map declared in the schema) and turn C-ABI error codes into
    Python exceptions."""
    ops = einx._lib.ops()
    rng = np.random.default_rng(4)
    B, H, W, K, D = 2, 64, 96, 64, 32
    score = torch.from_numpy(synth.score_map(rng, B, H, W)).to(DEV)
    raw = torch.from_numpy(synth.descriptor_map(rng, B, D, H, W)).to(DEV)

    def fn(score, raw):
        kp, cn, _ = torch.ops.einx.detect(score, None, 4, 4, 1.0, K, K, False)
        d = torch.ops.einx.sample(raw, kp, cn, 0, H, W, 1.41, True)
        return kp, cn, d * 2.0

    eager = fn(score.clone(), raw)
    compiled = torch.compile(fn, backend="aot_eager", fullgraph=True)(score.clone(), raw)
    for a, b in zip(eager, compiled):
        assert torch.equal(a, b)
    # the same numbers as the host wrappers (which call these ops)
    det = importlib.import_module("ei-nexus_official_b200.detection")
    _, kp, cn = det.detect(score.clone(), 1.0, 4, 4, K, kcap=K)
    assert torch.equal(kp, eager[0]) and torch.equal(cn, eager[1])
    with pytest.raises(RuntimeError, match="einx_detect"):
        ops.detect(score.clone(), None, 99, 4, 1.0, K, K, False)  # nms_radius out of range -> EINX_ERR_UNSUPPORTED


def test_sampler_split_operands_feed_the_matcher(einx):
    """einx_sample_split writes the matcher's FP16X3 operands (hi = fp16(2^10 d), lo = fp16(2^10 d - hi)) next to the
    fp32 descriptors; einx_mnn_split on them gives exactly what einx_mnn derives itself in its pre-pass."""
    synth = importlib.import_module("ei-nexus_official_b200.synth")
    desc = importlib.import_module("ei-nexus_official_b200.describe")
    det = importlib.import_module("ei-nexus_official_b200.detection")
    rng = np.random.default_rng(12)
    for mode, (B, D, H, W, cell, K, scale) in ((desc.BILINEAR, (3, 256, 96, 128, 8, 300, 1.0)), (desc.GATHER, (2, 128, 72, 88, 1, 257, 1.41))):
        sides = []
        for s in range(2):
            score = torch.from_numpy(synth.score_map(rng, B, H, W)).to(DEV)
            raw = torch.from_numpy(synth.descriptor_map(rng, B, D, H // cell, W // cell)).to(DEV)
            _, kp, cn = det.detect(score, 1.0, 4, 4, K, kcap=K)
            d_plain = desc.sample(raw, kp, cn, mode, (H, W), scale, True)
            d, sp = desc.sample(raw, kp, cn, mode, (H, W), scale, True, split=True)
            assert torch.equal(d, d_plain) and sp.shape == (2, B, K, D) and sp.dtype == torch.float16
            x = d * 1024.0
            hi = x.half()
            lo = (x - hi.float()).half()
            assert torch.equal(sp[0], hi) and torch.equal(sp[1], lo)
            if mode == desc.GATHER:  # channels-last map: same operands
                d2, sp2 = desc.sample(raw.contiguous(memory_format=torch.channels_last), kp, cn, mode, (H, W), scale, True, split=True)
                # (the channels-last kernel sums the norm in another order: 2e-6, like its own parity test)
                assert (d2 - d).abs().max().item() <= 2e-6
                x2 = d2 * 1024.0
                assert torch.equal(sp2[0], x2.half()) and torch.equal(sp2[1], (x2 - x2.half().float()).half())
            sides.append((d, sp, kp, cn))
        (d0, sp0, k0, c0), (d1, sp1, k1, c1) = sides
        a = einx.mnn(d0, d1, c0, c1, k0, k1, precision="fp16x3")
        b = einx.mnn(d0, d1, c0, c1, k0, k1, precision="fp16x3", split0=sp0, split1=sp1)
        for key in ("matches0", "matches1", "matching_scores0", "matching_scores1", "num_matches"):
            assert torch.equal(a[key], b[key]), key
        for i in range(B):  # (rows beyond num_matches are unspecified)
            n = int(a["num_matches"][i])
            assert torch.equal(a["matched_kpts0"][i, :n], b["matched_kpts0"][i, :n])
            assert torch.equal(a["matched_kpts1"][i, :n], b["matched_kpts1"][i, :n])
        ref = einx.mnn(d0, d1, c0, c1, k0, k1, precision="fp32")
        agree = (ref["matches0"] == b["matches0"]).float().mean().item()
        assert agree > 0.999, agree  # (index parity modulo fp32 near-ties is adjudicated in test_gpu_mnn_tc.py)


def test_bench_default_combination_against_oracle(einx):
    """The combination bench.py times -- C2 shapes, precision fp16x3 (pre-split operands written by the sampler),
    three streams, one CUDA-graph replay per step -- against the oracle pipeline pair by pair: voxel grid within
    1e-5 relative, keypoints bit-exact, match indices identical except where the fp32 and fp64 oracles themselves
    disagree (summation-order near-ties), adjudicated like the standalone matcher tests."""
    from oracle import einx_oracle as O

    synth = importlib.import_module("ei-nexus_official_b200.synth")
    name, B = "c2_ec_superpoint", 4
    c = synth.CONFIGS[name]
    evs, maps = [], [[], [], [], []]
    for i in range(B):
        ev, sides = synth.pair_inputs(name, 500 + i, None)
        evs.append(ev)
        for k, m in enumerate((sides[0][0], sides[0][1], sides[1][0], sides[1][1])):
            maps[k].append(m)
    maps = [np.concatenate(m) for m in maps]
    cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"],
                          descriptor_scale=c["scale"], precision="fp16x3", concurrent=True)
    pipe = einx.ExtractMatchPipeline(cfg)
    ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
    dmaps = [torch.from_numpy(m.copy()).to(DEV) for m in maps]
    step = pipe.capture(ev, *dmaps)
    for _ in range(2):
        for d, m in zip(dmaps, maps):  # the step zeroes the border of the score maps in place: refresh, then replay
            d.copy_(torch.from_numpy(m))
        out = step.replay()
    torch.cuda.synchronize()
    for i in range(B):
        grid, p0, p1, m = O.pair_pipeline(evs[i], c["bins"], c["H"], c["W"], maps[0][i:i + 1].copy(), maps[1][i:i + 1],
                                          maps[2][i:i + 1].copy(), maps[3][i:i + 1], "low", c["top_k"], c["scale"])
        g = out["voxel_grid"][i].cpu().numpy()
        assert np.all(np.abs(g - grid) <= 1e-5 * np.maximum(np.abs(grid), 1.0) + 1e-6)
        n0, n1 = int(out["counts0"][i]), int(out["counts1"][i])
        assert np.array_equal(out["keypoints0"][i, :n0].cpu().numpy(), p0) and np.array_equal(out["keypoints1"][i, :n1].cpu().numpy(), p1)
        d0, d1 = out["descriptors0"][i, :n0].cpu().numpy(), out["descriptors1"][i, :n1].cpu().numpy()
        r32 = O.mnn_match(d0, d1)
        r64 = O.mnn_match(d0, d1, sim_dtype=np.float64)
        got = out["matches0"][i, :n0].cpu().numpy()
        stable = r32["matches0"] == r64["matches0"]
        assert stable.mean() >= 0.999 and np.array_equal(got[stable], r32["matches0"][stable])
        assert int(out["num_matches"][i]) == int((got > -1).sum())
