"""CUDA path of the rows adjacent to the hot path (SURVEY.md section 8 f) vs reference goldens and the
oracle, through the C ABI (run with -m gpu on the B200 box)."""
import numpy as np
import pytest
import torch

from oracle import einx_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def einx():
    import einx as m

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.context_for(DEV)
    return m


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype else t).to(DEV)


# ------------------------------------------------------------- events image / mask ---- #
def test_events_image_golden(einx, golden):
    g = golden["next"]
    for ci in range(int(g["img_ncases"])):
        H, W, cell = (int(v) for v in g[f"img{ci}_shape"])
        ev = {"x": g[f"img{ci}_x"], "y": g[f"img{ci}_y"]}
        img = einx.draw_events_accumulation_image(ev, (W, H), DEV)
        assert img.dtype == np.uint8 and np.array_equal(img, g[f"img{ci}_out"])  # bit-exact (integer work)
        mask = einx.events_mask(cuda(img)[None], cell)
        assert mask.dtype == torch.bool and np.array_equal(mask[0, 0].cpu().numpy(), g[f"img{ci}_mask"])


def test_events_image_batch_vs_oracle(einx):
    import importlib

    synth = importlib.import_module("ei-nexus_official_b200.synth")
    rng = np.random.default_rng(5)
    H, W = 180, 240
    evs = [synth.events(rng, n, H, W, style) for n, style in ((60_000, "ec"), (1, "ec"), (20_000, "mvsec"), (333, "mvsec"))]
    x, y, t, p, off = (a.to(DEV) for a in einx.pack_events(evs))
    img = einx.events_image_device(x, y, off, H, W).cpu().numpy()
    for i, ev in enumerate(evs):
        # fp32 coordinates: same truncation as fp64 here (EC integral; MVSEC rounding never crosses an integer
        # for this seed -- checked by comparing with the fp32-rounded oracle)
        ref = O.draw_events_accumulation_image(ev["x"].astype(np.float32), ev["y"].astype(np.float32), H, W)
        assert np.array_equal(img[i], ref), i
    mask = einx.events_mask(torch.from_numpy(img).to(DEV), 8).cpu().numpy()
    assert mask.shape == (4, 1, 184, 240)
    for i in range(4):
        assert np.array_equal(mask[i, 0], O.events_mask(img[i], 8))


def test_event_mask_fused_into_detect(einx):
    """score[~mask] = 0 on load (EventExtractors.py:374-375): detect(score, mask) == detect(score * mask)."""
    rng = np.random.default_rng(9)
    score = rng.random((3, 1, 64, 80)).astype(np.float32)
    img = (rng.random((3, 64, 80)) < 0.02).astype(np.uint8) * 200
    mask = einx.events_mask(cuda(img), 8)
    a = cuda(score)
    _, k_a, c_a = einx.detect(a, 1.0, 4, 4, 50, mask=mask)
    ref_in = score * O.events_mask(img, 8)[:, None]
    for i in range(3):
        nms = O.prob_map_to_points_map(ref_in[i:i + 1].copy(), 1.0, 4, 4, 50)
        pos = O.prob_map_to_positions_with_prob(nms)[0]
        n = int(c_a[i])
        assert np.array_equal(k_a[i, :n].cpu().numpy(), pos)
    assert np.array_equal(a.cpu().numpy()[:, :, 4:-4, 4:-4], ref_in[:, :, 4:-4, 4:-4])  # zeroed in place


# ------------------------------------------------------------------ detector head ---- #
HEAD_RTOL = 2e-6  # fp32 exp / sum order; the reference's own CPU and CUDA softmax differ by as much


def test_detector_head_golden(einx, golden):
    g = golden["next"]
    lo = cuda(g["head_logits65"])
    np.testing.assert_allclose(einx.logits_to_prob(lo).cpu().numpy(), g["head_prob65"], rtol=HEAD_RTOL, atol=1e-9)
    np.testing.assert_allclose(einx.logits_to_score(lo, 8).cpu().numpy(), g["head_score65"], rtol=HEAD_RTOL, atol=1e-9)
    # pure data movement is exact
    assert np.array_equal(einx.depth_to_space(cuda(g["head_prob65"]), 8).cpu().numpy(), g["head_score65"])
    l1 = cuda(g["head_logits1"])
    np.testing.assert_allclose(einx.logits_to_prob(l1).cpu().numpy(), g["head_prob1"], rtol=HEAD_RTOL, atol=1e-9)
    np.testing.assert_allclose(einx.logits_to_score(l1, 1).cpu().numpy(), g["head_score1"], rtol=HEAD_RTOL, atol=1e-9)
    np.testing.assert_allclose(einx.logits_to_score(cuda(g["head_logits17"]), 4).cpu().numpy(), g["head_score17"],
                               rtol=HEAD_RTOL, atol=1e-9)


def test_detector_head_config_size(einx):
    rng = np.random.default_rng(11)
    lo = (3 * rng.standard_normal((8, 65, 23, 30))).astype(np.float32)  # EC SuperPoint head: 184x240 / 8
    score = einx.logits_to_score(cuda(lo), 8)
    assert score.shape == (8, 1, 184, 240)
    ref = O.depth_to_space(O.logits_to_prob(lo), 8)
    np.testing.assert_allclose(score.cpu().numpy(), ref, rtol=HEAD_RTOL, atol=1e-9)
    # partition of unity: every cell's 64 scores + dustbin sum to 1
    prob = einx.logits_to_prob(cuda(lo)).cpu().numpy()
    np.testing.assert_allclose(prob.sum(1), 1.0, atol=3e-6)
    with pytest.raises(einx.EinxError):
        einx.logits_to_score(cuda(lo[:, :64]), 8)  # the reference asserts C == cell^2 + 1


# ---------------------------------------------------------------- filter_matches ---- #
def test_filter_matches_golden(einx, golden):
    g = golden["next"]
    for ci in range(int(g["fm_ncases"])):
        m0, m1, s0, s1 = einx.filter_matches(cuda(g[f"fm{ci}_scores"]), float(g[f"fm{ci}_th"]))
        assert m0.dtype == torch.int64
        assert np.array_equal(m0.cpu().numpy(), g[f"fm{ci}_m0"]) and np.array_equal(m1.cpu().numpy(), g[f"fm{ci}_m1"])
        np.testing.assert_allclose(s0.cpu().numpy(), g[f"fm{ci}_s0"], rtol=2e-6)
        np.testing.assert_allclose(s1.cpu().numpy(), g[f"fm{ci}_s1"], rtol=2e-6)


def test_filter_matches_config_size(einx):
    rng = np.random.default_rng(13)
    B, M, N = 4, 1024, 1000
    s = rng.standard_normal((B, M + 1, N + 1)).astype(np.float32) - 6.0
    perm = rng.permutation(N)[:400]
    s[:, np.arange(400), perm] += 5.5  # planted mutual maxima
    m0, m1, s0, s1 = einx.filter_matches(cuda(s), 0.1)
    r0, r1, rs0, rs1 = O.filter_matches(s, 0.1)
    # indices are exact; scores within exp's ulp; a threshold flip would need exp(max) within 1e-6 of th
    assert np.array_equal(m0.cpu().numpy(), r0) and np.array_equal(m1.cpu().numpy(), r1)
    np.testing.assert_allclose(s0.cpu().numpy(), rs0, rtol=2e-6)
    np.testing.assert_allclose(s1.cpu().numpy(), rs1, rtol=2e-6)
    # mutual consistency
    a = m0.cpu().numpy()
    b = m1.cpu().numpy()
    for k in range(B):
        i = np.nonzero(a[k] >= 0)[0]
        assert np.array_equal(b[k][a[k][i]], i) and (a[k] >= 0).sum() == (b[k] >= 0).sum()
        assert (a[k] >= 0).sum() >= 400


# ------------------------------------------------ event stack / time surface ---- #
def test_event_stack_and_time_surface_golden(einx, golden):
    """Bit-exact against the reference's own outputs (integer sums; fp32-rounded fp64 times), including events
    that sit exactly on bin boundaries and therefore count in two bins."""
    g = golden["repr"]
    for ci in range(int(g["ncases"])):
        bins, H, W = (int(v) for v in g[f"c{ci}_shape"])
        ev = {k: g[f"c{ci}_{k}"].copy() for k in "xytp"}
        stack = einx.events_to_event_stack(dict(ev), (bins, H, W), DEV)
        assert stack.dtype == torch.float32 and np.array_equal(stack.numpy(), g[f"c{ci}_stack"]), ci
        mutated = dict(ev)
        surf = einx.events_to_time_surface(mutated, (bins, H, W), DEV)
        assert np.array_equal(surf.numpy(), g[f"c{ci}_surface"]), ci
        assert mutated["t"][0] == 0.0 and mutated["t"][-1] < 1.0  # the reference normalises the caller's dict


def test_binned_representations_batch_vs_oracle(einx):
    import importlib

    synth = importlib.import_module("ei-nexus_official_b200.synth")
    rng = np.random.default_rng(21)
    H, W, bins = 180, 240, 8
    evs = [synth.events(rng, n, H, W, "ec") for n in (60_000, 1, 5_000, 777)]
    x, y, t, p, off = (a.to(DEV) for a in einx.pack_events(evs))
    stack = einx.event_stack_device(x, y, t, p, off, (bins, H, W)).cpu().numpy()
    surf = einx.time_surface_device(x, y, t, p, off, (bins, H, W)).cpu().numpy()
    for i, ev in enumerate(evs):
        assert np.array_equal(stack[i], O.events_to_event_stack(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W)), i
        assert np.array_equal(surf[i], O.events_to_time_surface(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W)), i
    # every event lands in at least one bin: |stack| sums to >= the event count only when no +1/-1 cancel, so check the
    # polarity balance instead: sum over bins and pixels = sum(2p - 1) + boundary duplicates (none for random times)
    assert stack[0].sum() == (2 * evs[0]["p"].astype(np.int32) - 1).sum()


# ------------------------------------------------------- sigmoid_log_double_softmax ---- #
# of max(1, |ref|): fp32 exp / log / summation order.  Measured on B200 at the C2 size: 4.2e-7 against an fp64
# evaluation, 6e-7 against torch's CPU result (which is itself 5.3e-7 from the fp64 one).
LDS_TOL = 2e-6


def test_log_double_softmax_golden(einx, golden):
    g = golden["lg"]
    for ci in range(int(g["ncases"])):
        ref = g[f"c{ci}_scores"]
        out = einx.sigmoid_log_double_softmax(cuda(g[f"c{ci}_sim"]), cuda(g[f"c{ci}_z0"]), cuda(g[f"c{ci}_z1"]))
        assert out.shape == ref.shape and out.dtype == torch.float32
        out = out.cpu().numpy()
        assert np.all(np.abs(out - ref) <= LDS_TOL * np.maximum(1.0, np.abs(ref))), (ci, np.abs(out - ref).max())
        assert out[:, -1, -1].tolist() == [0.0] * ref.shape[0]
        # chained with filter_matches: the reference's matches come out of our matrix
        m0, m1, s0, s1 = einx.filter_matches(cuda(out), float(g[f"c{ci}_th"]))
        assert np.array_equal(m0.cpu().numpy(), g[f"c{ci}_m0"]) and np.array_equal(m1.cpu().numpy(), g[f"c{ci}_m1"])
        np.testing.assert_allclose(s0.cpu().numpy(), g[f"c{ci}_s0"], rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("chunk_mb", ["", "32"])
def test_log_double_softmax_config_size(einx, monkeypatch, chunk_mb):
    """C2 keypoint count; EINX_LDS_CHUNK_MB=32 walks the batch in chunks of 8 items with a ragged last chunk."""
    monkeypatch.setenv("EINX_LDS_CHUNK_MB", chunk_mb) if chunk_mb else monkeypatch.delenv("EINX_LDS_CHUNK_MB", raising=False)
    rng = np.random.default_rng(21)
    B, M, N = 19, 1024, 1000
    sim = (6.0 * rng.standard_normal((B, M, N))).astype(np.float32)
    z0 = (2.0 * rng.standard_normal((B, M, 1))).astype(np.float32)
    z1 = (2.0 * rng.standard_normal((B, N, 1))).astype(np.float32)
    out = einx.sigmoid_log_double_softmax(cuda(sim), cuda(z0), cuda(z1)).cpu().numpy()
    for b in (0, 7, 8, 18):
        ref = O.sigmoid_log_double_softmax(sim[b:b + 1], z0[b:b + 1], z1[b:b + 1])
        err = np.abs(out[b:b + 1] - ref) / np.maximum(1.0, np.abs(ref))
        print(f"log_double_softmax item {b}: max error {err.max():.3g} of max(1, |ref|)")
        assert err.max() <= LDS_TOL, b
    # size-independent property: exp(row log-softmax) sums to one  <=>  logsumexp_j(scores - col terms) ...
    # checked in the simplest form: scores - certainties = row log-softmax + column log-softmax <= 0
    t = torch.from_numpy(out[:, :-1, :-1])
    cert = torch.nn.functional.logsigmoid(torch.from_numpy(z0)) + torch.nn.functional.logsigmoid(torch.from_numpy(z1)).transpose(1, 2)
    assert float((t - cert).max()) <= 1e-5


@pytest.mark.parametrize("B,M,N", [(3, 70, 90), (2, 1024, 1000), (5, 129, 513), (1, 1, 1), (2, 64, 256), (1, 300, 31)])
def test_filter_matches_on_carried_keys(einx, B, M, N):
    """The maxima reduced while the matrix is written give exactly the matches of a separate pass over the stored
    matrix (ties included: duplicated rows / columns), and the carried keys are dropped once the tensor is modified."""
    rng = np.random.default_rng(B * 1000 + M + N)
    sim = (3.0 * rng.standard_normal((B, M, N))).astype(np.float32)
    if M > 12 and N > 8:
        sim[:, :, 7] = sim[:, :, 3]
        sim[:, 11, :] = sim[:, 2, :]
    z0 = rng.standard_normal((B, M, 1)).astype(np.float32)
    z1 = rng.standard_normal((B, N, 1)).astype(np.float32)
    if M > 12 and N > 8:
        z0[:, 11], z1[:, 7] = z0[:, 2], z1[:, 3]  # exact duplicates: identical scores, first index must win
    carried = einx.sigmoid_log_double_softmax(cuda(sim), cuda(z0), cuda(z1))
    plain = einx.sigmoid_log_double_softmax(cuda(sim), cuda(z0), cuda(z1), carry_best=False)
    assert torch.equal(carried, plain) and hasattr(carried, "_einx_best_keys") and not hasattr(plain, "_einx_best_keys")
    before = einx.launch_count(DEV)
    fused = einx.filter_matches(carried, 0.1)
    assert einx.launch_count(DEV) - before == 1  # only the filter kernel: no pass over the matrix
    ref = einx.filter_matches(plain, 0.1)
    for a, b in zip(fused, ref):
        assert torch.equal(a, b)
    o0, o1, _, _ = O.filter_matches(plain.cpu().numpy(), 0.1)
    assert np.array_equal(fused[0].cpu().numpy(), o0) and np.array_equal(fused[1].cpu().numpy(), o1)
    carried[:, 0, 0] += 100.0  # in-place edit: the carried keys are stale and must not be used
    before = einx.launch_count(DEV)
    edited = einx.filter_matches(carried, 0.1)
    assert einx.launch_count(DEV) - before > 1
    assert int(edited[0][0, 0]) == 0 and int(edited[1][0, 0]) == 0


@pytest.mark.parametrize("M,N", [(0, 5), (7, 0), (0, 0)])
def test_log_double_softmax_empty_side(einx, M, N):
    """No keypoints on one side: only the unmatched row / column exists (the reference's torch ops accept the shape)."""
    rng = np.random.default_rng(M + 10 * N)
    sim = np.zeros((2, M, N), np.float32)
    z0 = rng.standard_normal((2, M, 1)).astype(np.float32)
    z1 = rng.standard_normal((2, N, 1)).astype(np.float32)
    out = einx.sigmoid_log_double_softmax(cuda(sim), cuda(z0), cuda(z1)).cpu().numpy()
    ref = O.sigmoid_log_double_softmax(sim, z0, z1)
    assert out.shape == ref.shape == (2, M + 1, N + 1)
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=1e-7)


def test_log_double_softmax_rejects_cpu(einx):
    with pytest.raises(einx.EinxError):
        einx.sigmoid_log_double_softmax(torch.zeros(1, 2, 2), torch.zeros(1, 2, 1), torch.zeros(1, 2, 1))


# --------------------------------------------------- distance map / metric-side reductions ---- #
def test_distance_map_golden_and_oracle(einx, golden):
    """einx_distance_map vs the reference's cv.distanceTransform output (1e-6 relative, the bar the oracle is pinned
    with) and bit-exact against the oracle's closed form."""
    g = golden["repr"]
    cases = [f"c{ci}" for ci in range(int(g["ncases"]))] + ["sparse"]
    for key in cases:
        bins, H, W = (int(v) for v in g[f"{key}_shape"])
        ev = {k: np.array(g[f"{key}_{k}"], copy=True) for k in "xytp"}
        out = einx.events_to_distance_map(ev, (bins, H, W), DEV).numpy()
        ref = g[f"{key}_distance"]
        empty = ref > 1e30
        assert out.shape == ref.shape and np.array_equal(out > 1e30, empty)
        np.testing.assert_allclose(out[~empty], ref[~empty], rtol=1e-6, atol=0)
        orc = O.events_to_distance_map(*[g[f"{key}_{k}"] for k in "xytp"], bins, H, W)
        assert np.array_equal(out, orc), key  # same closed form, fp64 rounded once
        # the reference normalises events['t'] in place
        assert ev["t"].min() >= 0.0 and ev["t"].max() <= 1.0


def test_distance_map_batch_ragged(einx):
    import importlib

    synth = importlib.import_module("ei-nexus_official_b200.synth")
    rng = np.random.default_rng(11)
    H, W, bins = 65, 97, 4
    evs = [synth.events(rng, n, H, W, style) for n, style in ((5000, "mvsec"), (3, "ec"), (800, "ec"))]
    x, y, t, p, off = (a.to(DEV) for a in einx.pack_events(evs))
    out = einx.distance_map_device(x, y, t, off, (bins, H, W)).cpu().numpy()
    for i, ev in enumerate(evs):
        ref = O.events_to_distance_map(ev["x"].astype(np.float32), ev["y"].astype(np.float32), ev["t"], ev["p"], bins, H, W)
        assert np.array_equal(out[i], ref), i


def test_accumulation_image_array_branch(einx):
    """(N, 4) ndarray branch of draw_events_accumulation_image (visualize.py:41-44): += 2p - 1, signed histogram."""
    rng = np.random.default_rng(3)
    H, W, n = 48, 64, 5000
    ev = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n), np.sort(rng.random(n)), rng.integers(0, 2, n)], 1)
    img = einx.draw_events_accumulation_image(ev, (W, H), DEV)
    acc = np.zeros((H, W))
    for i in range(n):  # the reference's loop, verbatim semantics
        acc[int(ev[i, 1]), int(ev[i, 0])] += 2 * ev[i, 3] - 1
    ref = (acc - acc.min()) / (acc.max() - acc.min()) * 255
    ref[ref > 255] = 255
    assert img.dtype == np.uint8 and np.array_equal(img, ref.astype(np.uint8))
    with pytest.raises(ValueError):
        einx.draw_events_accumulation_image(np.array([[1.0, 1.0, 0.0, 0.3]]), (W, H), DEV)


def test_repeatability_golden(einx, golden):
    g = golden["metrics"]
    for ci in range(int(g["ncases"])):
        H, W, thr, xy = (int(v) for v in g[f"c{ci}_cfg"])
        metric = einx.Repeatability("repeatability", distance_thresh=thr, ordering="xy" if xy else "yx", device=DEV)
        out = metric.update_one(torch.from_numpy(g[f"c{ci}_p1"]), torch.from_numpy(g[f"c{ci}_p2"]), (H, W), (H, W),
                                torch.from_numpy(g[f"c{ci}_hom"]))
        ref = float(g[f"c{ci}_value"])
        if np.isnan(ref):
            assert out == {}
        else:
            assert np.float32(out["repeatability"]) == np.float32(ref), ci
        # the reductions themselves on the oracle's warped points (so the comparison is of the kernel alone): bit-exact
        sel = [0, 1] if xy else [1, 0]
        q1 = O._keep_true_points(g[f"c{ci}_p1"].T[sel], g[f"c{ci}_hom"], (H, W))
        q2 = O._keep_true_points(g[f"c{ci}_p2"].T[sel], np.linalg.inv(g[f"c{ci}_hom"]).astype(np.float32), (H, W))
        wp = O._warp_points(q1, g[f"c{ci}_hom"]).T
        if wp.shape[0] and q2.shape[1]:
            min1, min2 = metric.min_distances(cuda(wp), cuda(q2.T))
            d = wp[:, None, :].astype(np.float32) - q2.T[None, :, :].astype(np.float32)
            norm = np.sqrt((d.astype(np.float64) ** 2).sum(-1)).astype(np.float32)
            assert np.array_equal(min1.cpu().numpy(), norm.min(0)) and np.array_equal(min2.cpu().numpy(), norm.min(1))
            np.testing.assert_allclose(min1.cpu().numpy(), g[f"c{ci}_min_over_1"], rtol=1e-5, atol=1e-4)
            np.testing.assert_allclose(min2.cpu().numpy(), g[f"c{ci}_min_over_2"], rtol=1e-5, atol=1e-4)


def test_pairwise_min_dist_ragged_counts(einx):
    rng = np.random.default_rng(9)
    B, N, M = 3, 300, 517
    a = rng.uniform(0, 200, (B, N, 2)).astype(np.float32)
    b = rng.uniform(0, 200, (B, M, 2)).astype(np.float32)
    na = np.array([300, 0, 129], dtype=np.int32)
    nb = np.array([517, 40, 0], dtype=np.int32)
    rmin, cmin = einx.pairwise_min_dist(cuda(a), cuda(b), cuda(na), cuda(nb))
    rmin, cmin = rmin.cpu().numpy(), cmin.cpu().numpy()
    for i in range(B):
        d = a[i, :na[i], None, :] - b[i, None, :nb[i], :]
        norm = np.sqrt((d.astype(np.float64) ** 2).sum(-1)).astype(np.float32)
        if nb[i]:
            assert np.array_equal(rmin[i, :na[i]], norm.min(1))
        else:
            assert np.all(np.isinf(rmin[i, :na[i]]))
        if na[i]:
            assert np.array_equal(cmin[i, :nb[i]], norm.min(0))
        else:
            assert np.all(np.isinf(cmin[i, :nb[i]]))
        assert np.all(np.isinf(rmin[i, na[i]:])) and np.all(np.isinf(cmin[i, nb[i]:]))


def test_gt_assign_golden_and_oracle(einx, golden):
    """einx_gt_assign vs the reference's gt_matches_from_pose_depth outputs (bit-exact: integer indices)."""
    g = golden["gt_assign"]
    for ci in range(int(g["ncases"])):
        pos_th, neg_th = (float(v) for v in g[f"c{ci}_th"])
        args = [cuda(g[f"c{ci}_{k}"]) for k in ("kp0", "kp1", "kp0_1", "kp1_0", "visible0", "visible1", "valid0", "valid1")]
        a, m0, m1 = einx.gt_assign(*args, pos_th=pos_th, neg_th=neg_th)
        M = g[f"c{ci}_kp1"].shape[1]
        want = np.unpackbits(g[f"c{ci}_assignment"], axis=-1)[..., :M].astype(bool)
        assert np.array_equal(m0.cpu().numpy(), g[f"c{ci}_m0"]) and np.array_equal(m1.cpu().numpy(), g[f"c{ci}_m1"]), ci
        assert np.array_equal(a.cpu().numpy(), want), ci
    # ties and fully invisible rows: integer grid points, duplicated columns
    rng = np.random.default_rng(2)
    B, N, M = 2, 70, 90
    kp0 = rng.integers(0, 12, (B, N, 2)).astype(np.float32)
    kp1 = rng.integers(0, 12, (B, M, 2)).astype(np.float32)
    kp0_1 = kp0 + rng.integers(-1, 2, (B, N, 2)).astype(np.float32)
    kp1_0 = kp1 + rng.integers(-1, 2, (B, M, 2)).astype(np.float32)
    vis0, vis1 = rng.random((B, N)) < 0.8, rng.random((B, M)) < 0.8
    vis0[0, :5] = False
    val0, val1 = vis0 | (rng.random((B, N)) < 0.5), vis1 | (rng.random((B, M)) < 0.5)
    ao, m0o, m1o = O.gt_assign(kp0, kp1, kp0_1, kp1_0, vis0, vis1, val0, val1, 2, 3)
    a, m0, m1 = einx.gt_assign(*[cuda(v) for v in (kp0, kp1, kp0_1, kp1_0, vis0, vis1, val0, val1)], pos_th=2, neg_th=3)
    assert np.array_equal(m0.cpu().numpy(), m0o) and np.array_equal(m1.cpu().numpy(), m1o)
    assert np.array_equal(a.cpu().numpy(), ao)
    # empty side: the reference's early return
    a, m0, m1 = einx.gt_assign(cuda(kp0), cuda(kp1[:, :0]), cuda(kp0_1), cuda(kp1_0[:, :0]), cuda(vis0), cuda(vis1[:, :0]),
                               cuda(val0), cuda(val1[:, :0]))
    assert a.shape == (B, N, 0) and bool((m0 == -1).all()) and m1.shape == (B, 0)
