"""CUDA path vs oracle / reference goldens, through the C ABI (run with -m gpu on the B200 box)."""
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
    m.context_for(DEV)  # fails loudly if libeinx.so is missing or the device is not sm_100a
    return m


@pytest.fixture(scope="module")
def synth(einx):
    import importlib

    return importlib.import_module("ei-nexus_official_b200.synth")


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype else t).to(DEV)


def voxel_ok(got, ref, l1):
    return np.abs(got - ref) <= 1e-5 * np.maximum(np.abs(ref), l1) + 1e-30


# ------------------------------------------------------------------ voxelisation ---- #
def test_voxel_golden(einx, golden):
    g = golden["voxel"]
    for ci in range(int(g["ncases"])):
        bins, H, W = (int(v) for v in g[f"c{ci}_shape"])
        ev = {k: g[f"c{ci}_{k}"].copy() for k in "xytp"}
        _, l1 = O.events_to_voxel_grid(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W, False, return_l1=True)
        raw = einx.voxelize_batch([ev], (bins, H, W), normalize=False, device=DEV)[0].cpu().numpy()
        assert voxel_ok(raw, g[f"c{ci}_raw"], l1).all(), ci
        assert ((raw != 0) == (g[f"c{ci}_raw"] != 0)).all()
        nrm = einx.events_to_voxel_grid(ev, (bins, H, W), normalize=True, device=DEV)
        assert nrm.device.type == "cpu" and nrm.dtype == torch.float32 and tuple(nrm.shape) == (bins, H, W)
        ref = g[f"c{ci}_norm"]
        assert (np.abs(nrm.numpy() - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)).all(), ci
        # reference side effects on the caller's dict (representations.py:72-76, :88-89)
        assert isinstance(ev["t"], torch.Tensor) and ev["t"].dtype == torch.float32
        assert float(ev["t"][0]) == 0.0 and float(ev["t"][-1]) <= 1.0
        assert set(np.unique(ev["p"].numpy())) <= {-1.0, 1.0}


@pytest.mark.parametrize("style,n,bins,H,W", [("mvsec", 200_000, 5, 260, 346), ("ec", 60_000, 16, 180, 240),
                                               ("mvsec", 300_000, 10, 720, 1280)])
def test_voxel_vs_oracle_config_sizes(einx, synth, style, n, bins, H, W):
    rng = np.random.default_rng(7)
    batch = [synth.events(rng, n - 17 * i, H, W, style, clustered=(i == 1)) for i in range(3)]  # ragged
    got = einx.voxelize_batch(batch, (bins, H, W), normalize=False, device=DEV).cpu().numpy()
    gotn = einx.voxelize_batch(batch, (bins, H, W), normalize=True, device=DEV).cpu().numpy()
    for i, ev in enumerate(batch):
        ref, l1 = O.events_to_voxel_grid(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W, False, return_l1=True)
        assert voxel_ok(got[i], ref, l1).all()
        refn = O.normalize_nonzero(ref)
        same_mask = (got[i] != 0) == (ref != 0)
        assert same_mask.mean() > 0.99999
        ok = np.abs(gotn[i] - refn) <= 1e-5 * np.maximum(np.abs(refn), 1.0)
        assert ok[same_mask].all()
    # partition of unity: interior events deposit exactly their polarity
    s = got.reshape(3, -1).astype(np.float64).sum(1)
    for i, ev in enumerate(batch):
        pol = np.where(ev["p"] < 1, -1.0, ev["p"]).sum()
        assert abs(s[i] - pol) < 1e-3 * max(1.0, np.sqrt(len(ev["p"])))


def test_voxel_edge_cases(einx):
    one = {"x": np.array([3.2]), "y": np.array([4.7]), "t": np.array([1.5e9]), "p": np.array([1.0])}
    assert not einx.voxelize_batch([one], (5, 8, 8), device=DEV).any()  # 0/0 time span -> nothing lands
    with pytest.raises(IndexError):
        einx.voxelize_batch([{k: np.zeros(0) for k in "xytp"}], (5, 8, 8), device=DEV)
    # out-of-range coordinates are dropped corner by corner, like the reference's mask
    ev = {"x": np.array([-0.5, 7.5, 3.0, 100.0]), "y": np.array([0.2, 7.9, -3.0, 2.0]),
          "t": np.array([0.0, 0.1, 0.2, 0.3]) + 1.5e9, "p": np.array([1.0, 0.0, 2.0, -1.0])}
    got = einx.voxelize_batch([ev], (3, 8, 8), normalize=False, device=DEV)[0].cpu().numpy()
    ref, l1 = O.events_to_voxel_grid(ev["x"], ev["y"], ev["t"], ev["p"], 3, 8, 8, False, return_l1=True)
    assert voxel_ok(got, ref, l1).all()


@pytest.mark.parametrize("bins,H,W", [(5, 180, 240), (5, 260, 346), (8, 96, 128), (3, 700, 400)])
def test_voxel_unsorted_and_tiny_windows(einx, synth, bins, H, W):
    """The fused cluster path finds each bin's events by searching the time-sorted window; unsorted
    windows (checked on the device) and windows of a handful of events must still match."""
    rng = np.random.default_rng(11)
    batch = [synth.events(rng, 30_000, H, W, "mvsec"), synth.events(rng, 20_000, H, W, "ec")]
    perm = rng.permutation(20_000)
    batch[1] = {k: v[perm] for k, v in batch[1].items()}  # unsorted: t[0] / t[-1] are no longer min / max
    for n in (2, 3, 7, 1500):
        batch.append(synth.events(rng, n, H, W, "mvsec"))
    same_t = synth.events(rng, 50, H, W, "mvsec")
    same_t["t"][10:40] = same_t["t"][10]  # a run of equal timestamps
    batch.append(same_t)
    for normalize in (False, True):
        got = einx.voxelize_batch(batch, (bins, H, W), normalize=normalize, device=DEV).cpu().numpy()
        for i, ev in enumerate(batch):
            ref, l1 = O.events_to_voxel_grid(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W, False, return_l1=True)
            if not normalize:
                assert voxel_ok(got[i], ref, l1).all(), (i, normalize)
            else:
                refn = O.normalize_nonzero(ref)
                mask = (ref != 0) & (np.abs(ref) > 1e-6 * l1)  # cells that cancel to ~0 may flip the != 0 mask
                ok = np.abs(got[i] - refn) <= 2e-5 * np.maximum(np.abs(refn), 1.0)
                assert ok[mask].mean() > 0.9999, (i, normalize)


# ------------------------------------------------------------------ detection ------- #
def _detect_cases(g):
    for tag in g["tags"]:
        tag = str(tag)
        if "_r" in tag and "_b" in tag and "_k" not in tag:
            r, b = tag.split("_r")[1].split("_b")
            yield tag, "uniform", int(r), int(b), 50, 0.0
        else:
            name, rest = tag.rsplit("_k", 1)
            k, thr = rest.split("_t")
            yield tag, name, 4, 4, (None if k == "None" else int(k)), float(thr)


def test_detect_golden_bit_exact(einx, golden):
    g = golden["detect"]
    for tag, name, r, b, k, thr in _detect_cases(g):
        src = cuda(g[f"{name}_in"])
        nms = einx.prob_map_to_points_map(src, prob_thresh=thr, nms_dist=r, border_dist=b, use_fast_nms=True, top_k=k)
        border_key = f"{tag}_border" if f"{tag}_border" in g.files else f"{name}_border"
        assert np.array_equal(src.cpu().numpy(), g[border_key]), tag  # in-place border zeroing
        pos = einx.prob_map_to_positions_with_prob(nms, threshold=0.0, ordering="yx")
        assert len(pos) == src.shape[0]
        for i, p in enumerate(pos):
            assert np.array_equal(p.cpu().numpy(), g[f"{tag}_pos{i}"]), (tag, i)
        # the dense map and the keypoint rows agree
        ref_map = O.prob_map_to_points_map(g[f"{name}_in"].copy(), thr, r, b, k)
        assert np.array_equal(nms.cpu().numpy(), ref_map), tag


def test_detect_reference_property_tests(einx, golden):
    """utils_test.py:17-63 on the CUDA path."""
    g = golden["detect"]
    torch.manual_seed(0)
    inp = torch.rand((32, 60, 80))
    got = einx.prob_map_to_points_map(inp.to(DEV), 0.0, 4, 4, use_fast_nms=True).cpu().numpy()
    nz = np.argwhere(got != 0).astype(np.int32)
    assert np.array_equal(nz, g["parity_nz"])
    assert np.array_equal(got[tuple(nz.T)], g["parity_val"])
    small = torch.rand((1, 1, 8, 8)).to(DEV)
    einx.prob_map_to_points_map(small, 0.0, 4, 4)
    assert not small.any()


@pytest.mark.parametrize("B,Hp,Wp,k,kind", [(3, 184, 240, 1024, "uniform"), (5, 264, 352, 2048, "uniform"),
                                            (2, 260, 346, 2048, "ties"), (2, 184, 240, 300, "ties"),
                                            (1, 720, 1280, 8192, "uniform"), (1, 720, 1280, 512, "ties"),
                                            (1, 720, 1280, 16384, "uniform"),  # top of the C5 keypoint sweep
                                            (150, 64, 96, 100, "uniform"), (2, 37, 1000, 64, "uniform")])
def test_detect_vs_oracle_config_sizes(einx, synth, B, Hp, Wp, k, kind):
    rng = np.random.default_rng(Hp * 7 + k)
    m = synth.score_map(rng, B, Hp, Wp, kind)
    if kind == "uniform" and B > 1:
        m[1, :, :, Wp // 3:] = 0  # sparse image: fewer than k survivors -> threshold 0
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 1.0, 4, 4, k, want_map=True)
    ref_in = m.copy()
    ref = O.prob_map_to_points_map(ref_in, 1.0, 4, 4, k)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    pos = O.prob_map_to_positions_with_prob(ref)
    counts = counts.cpu().numpy()
    kp = kpts.cpu().numpy()
    for i in range(B):
        assert counts[i] == len(pos[i]) and counts[i] <= k
        assert np.array_equal(kp[i, : counts[i]], pos[i])
    # idempotence: the fixpoint map is its own NMS
    again, _, c2 = einx.detect(nms.clone(), 0.0, 4, 0, None, want_map=True)
    assert torch.equal(again, nms) and np.array_equal(c2.cpu().numpy(), counts)
    # keypoints are pairwise more than r apart (Chebyshev)
    p = kp[0, : counts[0], :2]
    if len(p) > 1:
        d = np.abs(p[:, None, :] - p[None, :, :]).max(-1) + np.eye(len(p)) * 100
        assert d.min() > 4


def _smooth_map(rng, B, Hp, Wp, sigma):
    from scipy.ndimage import gaussian_filter

    v = np.stack([gaussian_filter(rng.random((Hp, Wp)), sigma) for _ in range(B)])
    v = (v - v.min()) / (v.max() - v.min())
    return v[:, None].astype(np.float32)


@pytest.mark.parametrize("bands", [0, 1, 2, 3, 4, 8])
@pytest.mark.parametrize("B,Hp,Wp,k,kind", [(2, 184, 240, 1024, "uniform"), (2, 184, 240, 1024, "ties"),
                                            (2, 184, 240, 500, "smooth"), (3, 130, 346, 700, "uniform"),
                                            (2, 100, 70, 64, "smooth")])
def test_detect_every_cluster_size(einx, synth, monkeypatch, bands, B, Hp, Wp, k, kind):
    """The single-CTA kernel and every cluster split (bands exchange their edge rows over DSMEM) give the oracle's
    bits; smooth maps need 10+ rounds, most of them dense, ties exercise the first-occurrence rule."""
    rng = np.random.default_rng(bands * 131 + Hp + k)
    m = _smooth_map(rng, B, Hp, Wp, 5.0) if kind == "smooth" else synth.score_map(rng, B, Hp, Wp, kind)
    if bands:
        monkeypatch.setenv("EINX_DETECT_CLUSTER", str(bands))
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 1.0, 4, 4, k, want_map=True)
    ref_in = m.copy()
    ref = O.prob_map_to_points_map(ref_in, 1.0, 4, 4, k)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    pos = O.prob_map_to_positions_with_prob(ref)
    counts, kp = counts.cpu().numpy(), kpts.cpu().numpy()
    for i in range(B):
        assert counts[i] == len(pos[i])
        assert np.array_equal(kp[i, : counts[i]], pos[i])


@pytest.mark.parametrize("r", [0, 1, 2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("bands", [0, 2])
def test_detect_other_radii(einx, synth, monkeypatch, r, bands):
    rng = np.random.default_rng(50 + r)
    B, Hp, Wp = 2, 120, 150  # width not a multiple of 4: two-pixel global accesses
    m = synth.score_map(rng, B, Hp, Wp, "ties" if r % 2 else "uniform")
    if bands:
        monkeypatch.setenv("EINX_DETECT_CLUSTER", str(bands))
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 0.3, r, 3, None, want_map=True)
    ref_in = m.copy()
    ref = O.prob_map_to_points_map(ref_in, 0.3, r, 3, None)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    pos = O.prob_map_to_positions_with_prob(ref)
    for i in range(B):
        assert int(counts[i]) == len(pos[i])
        assert np.array_equal(kpts[i, : int(counts[i])].cpu().numpy(), pos[i])


def test_detect_full_batch_single_cta(einx, synth):
    """Batch 64 of EC-size maps: one image per CTA (the bench configuration), odd width variant included."""
    rng = np.random.default_rng(64)
    for Hp, Wp, k in ((184, 240, 1024), (91, 133, 200)):
        m = synth.score_map(rng, 64, Hp, Wp)
        m[5] = 0
        m[6, :, : Hp // 2] = 0
        src = cuda(m)
        _, kpts, counts = einx.detect(src, 1.0, 4, 4, k, want_map=False)
        ref = O.prob_map_to_points_map(m.copy(), 1.0, 4, 4, k)
        pos = O.prob_map_to_positions_with_prob(ref)
        counts, kp = counts.cpu().numpy(), kpts.cpu().numpy()
        for i in range(64):
            assert counts[i] == len(pos[i]), (Hp, i)
            assert np.array_equal(kp[i, : counts[i]], pos[i]), (Hp, i)


def test_detect_pair_is_two_detects(einx, synth):
    """einx_detect_pair: both sides of a batch of pairs in one launch, bit-identical to one launch per side."""
    rng = np.random.default_rng(77)
    for B, Hp, Wp, k in ((64, 184, 240, 1024), (3, 260, 346, 2048), (1, 96, 128, 100)):
        a, b = synth.score_map(rng, B, Hp, Wp), synth.score_map(rng, B, Hp, Wp, "ties")
        mb = rng.random((B, 1, Hp, Wp)) > 0.3
        sa, sb = cuda(a), cuda(b)
        (k0, c0), (k1, c1) = einx.detect_pair(sa, sb, 1.0, 4, 4, k, None, cuda(mb))
        ra, rb = cuda(a), cuda(b)
        _, e0, d0 = einx.detect(ra, 1.0, 4, 4, k)
        _, e1, d1 = einx.detect(rb, 1.0, 4, 4, k, mask=cuda(mb))
        assert torch.equal(sa, ra) and torch.equal(sb, rb)  # in-place border / mask zeroing
        assert torch.equal(c0, d0) and torch.equal(c1, d1)
        for i in range(B):
            assert torch.equal(k0[i, : int(c0[i])], e0[i, : int(d0[i])])
            assert torch.equal(k1[i, : int(c1[i])], e1[i, : int(d1[i])])


def test_detect_mask_and_positions_generic(einx, synth):
    rng = np.random.default_rng(3)
    m = synth.score_map(rng, 2, 96, 128)
    mask = rng.random((2, 1, 96, 128)) > 0.4
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 1.0, 4, 4, 200, mask=cuda(mask), want_map=True)
    ref_in = m.copy()
    ref_in[~mask] = 0
    ref = O.prob_map_to_points_map(ref_in, 1.0, 4, 4, 200)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    # positions of an arbitrary map / threshold, 'xy' ordering
    pos = einx.prob_map_to_positions_with_prob(cuda(m), threshold=0.97, ordering="xy")
    refp = O.prob_map_to_positions_with_prob(m, 0.97, "xy")
    for a, b in zip(pos, refp):
        assert np.array_equal(a.cpu().numpy(), b)


# ------------------------------------------------------------------ sampling -------- #
def test_sample_golden(einx, golden):
    g = golden["sample"]
    pos = [cuda(g["pos0"]), cuda(g["pos1"])]
    full = einx.sparsify_full_resolution_descriptors(cuda(g["raw_full"]), pos, torch.tensor(1.41), True)
    low = einx.sparsify_low_resolution_descriptors(cuda(g["raw_low"]), pos, (48, 64), torch.tensor(1.0), True)
    assert isinstance(full, tuple) and isinstance(low, list)
    for i in range(2):
        assert np.abs(full[i].cpu().numpy() - g[f"full{i}"]).max() < 2e-6
        assert np.abs(low[i].cpu().numpy() - g[f"low{i}"]).max() < 2e-6
    empty = einx.sparsify_low_resolution_descriptors(cuda(g["raw_low"][:1]), [torch.zeros((0, 3), device=DEV)], (48, 64))
    assert tuple(empty[0].shape) == (0, 32)


@pytest.mark.parametrize("mode,C,Hp,Wp,cell,scale", [("gather", 128, 260, 346, 1, 1.41), ("bilinear", 256, 184, 240, 8, 1.0),
                                                     ("bilinear", 256, 264, 352, 8, 1.0), ("gather", 96, 40, 56, 1, 1.0)])
def test_sample_vs_oracle_config_sizes(einx, synth, mode, C, Hp, Wp, cell, scale):
    rng = np.random.default_rng(C + Hp)
    B = 2
    raw = synth.descriptor_map(rng, B, C, Hp // cell, Wp // cell)
    nms = O.prob_map_to_points_map(synth.score_map(rng, B, Hp, Wp), 1.0, 4, 4, 1024)
    pos = O.prob_map_to_positions_with_prob(nms)
    if mode == "gather":
        ref = O.sparsify_full_resolution_descriptors(raw, pos, scale, True)
        got = einx.sparsify_full_resolution_descriptors(cuda(raw), [cuda(p) for p in pos], scale, True)
    else:
        ref = O.sparsify_low_resolution_descriptors(raw, pos, (Hp, Wp), scale, True)
        got = einx.sparsify_low_resolution_descriptors(cuda(raw), [cuda(p) for p in pos], (Hp, Wp), scale, True)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert np.abs(a.cpu().numpy() - b).max() < 2e-6
        assert np.abs(np.linalg.norm(a.cpu().numpy().astype(np.float64), axis=1) - scale).max() < 1e-5


@pytest.mark.parametrize("C,Hp,Wp,scale", [(128, 260, 346, 1.41), (256, 64, 80, 1.0), (96, 40, 56, 1.0), (36, 33, 47, 2.0)])
def test_sample_gather_channels_last(einx, synth, C, Hp, Wp, scale):
    """A channels-last descriptor map (the layout cuDNN convolutions produce) is gathered in place -- one contiguous
    read per keypoint -- and gives the same descriptors as the NCHW map (descriptor_util.py:50-71)."""
    rng = np.random.default_rng(C * 3 + Hp)
    B = 3
    raw = synth.descriptor_map(rng, B, C, Hp, Wp)
    nms = O.prob_map_to_points_map(synth.score_map(rng, B, Hp, Wp), 1.0, 4, 4, 700)
    pos = list(O.prob_map_to_positions_with_prob(nms))
    pos[1] = pos[1][:0]  # an image without keypoints
    ref = O.sparsify_full_resolution_descriptors(raw, pos, scale, True)
    nchw = cuda(raw)
    nhwc = nchw.contiguous(memory_format=torch.channels_last)
    assert not nhwc.is_contiguous() and nhwc.shape == nchw.shape
    got_a = einx.sparsify_full_resolution_descriptors(nchw, [cuda(p) for p in pos], scale, True)
    got_b = einx.sparsify_full_resolution_descriptors(nhwc, [cuda(p) for p in pos], scale, True)
    for a, b, r in zip(got_a, got_b, ref):
        assert a.shape == r.shape and b.shape == r.shape
        assert np.abs(b.cpu().numpy() - r).max() < 2e-6 if r.size else True
        assert np.abs(a.cpu().numpy() - b.cpu().numpy()).max() < 2e-6 if r.size else True
    # padded form: rows beyond the count are zero
    kpts, counts = importlib_describe().pack_rows([cuda(p) for p in pos], 3, DEV)
    desc = einx.sample(nhwc, kpts, counts, 0, (Hp, Wp), scale, True)
    for i in range(B):
        assert not desc[i, len(pos[i]):].any()


def importlib_describe():
    import importlib

    return importlib.import_module("ei-nexus_official_b200.describe")


# ------------------------------------------------------------------ MNN ------------- #
def test_mnn_golden_bit_exact(einx, golden):
    g = golden["mnn"]
    for ci in range(int(g["ncases"])):
        ratio, dist = (float(v) or None for v in g[f"c{ci}_cfg"])
        matcher = einx.NearestNeighborMatcher(ratio, dist, True, return_dense=True)
        f0 = {"sparse_descriptors": cuda(g[f"c{ci}_d0"])[None], "sparse_positions": cuda(g[f"c{ci}_k0"])[None]}
        f1 = {"sparse_descriptors": cuda(g[f"c{ci}_d1"])[None], "sparse_positions": cuda(g[f"c{ci}_k1"])[None]}
        out = matcher(f0, f1)
        assert out["matches0"].dtype == torch.int64 and tuple(out["matches0"].shape) == (1, g[f"c{ci}_d0"].shape[0])
        for key in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
            assert np.array_equal(out[key][0].cpu().numpy(), g[f"c{ci}_{key}"]), (ci, key)
        for key in ("matched_kpts0", "matched_kpts1"):
            assert np.array_equal(out[key].cpu().numpy(), g[f"c{ci}_{key}"]), (ci, key)
        assert np.abs(out["log_assignment"][0].cpu().numpy() - g[f"c{ci}_log_assignment"]).max() < 1e-5
        d0, d1 = g[f"c{ci}_d0"], g[f"c{ci}_d1"]
        assert np.abs(out["similarity"][0].cpu().numpy() - d0 @ d1.T).max() < 1e-5


def adjudicate(got0, d0, d1, ratio=None, dist=None):
    """Index parity modulo near-ties: rows where fp32 and fp64 oracles disagree are summation-order
    dependent inside the reference itself; everything else must be bit-equal."""
    r32 = O.mnn_match(d0, d1, ratio_thresh=ratio, distance_thresh=dist)["matches0"]
    r64 = O.mnn_match(d0, d1, ratio_thresh=ratio, distance_thresh=dist, sim_dtype=np.float64)["matches0"]
    stable = r32 == r64
    assert stable.mean() > 0.999
    assert np.array_equal(got0[stable], r32[stable])
    return int((~stable).sum())


@pytest.mark.parametrize("N,M,D,scale", [(1024, 1024, 256, 1.0), (2048, 2048, 128, 1.41), (1000, 777, 128, 1.41),
                                         (130, 4100, 64, 1.0)])
def test_mnn_fp32_vs_oracle_config_sizes(einx, synth, N, M, D, scale):
    rng = np.random.default_rng(N + M)
    pairs = [synth.descriptor_pair(rng, N, M, D, scale, dups=4 if b == 0 else 0) for b in range(2)]
    d0 = cuda(np.stack([p[0] for p in pairs]))
    d1 = cuda(np.stack([p[1] for p in pairs]))
    out = einx.mnn(d0, d1, precision="fp32")
    m0, m1 = out["matches0"].cpu().numpy(), out["matches1"].cpu().numpy()
    for b in range(2):
        adjudicate(m0[b], pairs[b][0], pairs[b][1])
        # mutual consistency + equal counts (MNN.py:95)
        keep = m0[b] > -1
        assert np.array_equal(m1[b][m0[b][keep]], np.nonzero(keep)[0])
        assert keep.sum() == (m1[b] > -1).sum() and keep.sum() >= 1
    assert np.array_equal(out["matching_scores0"].cpu().numpy(), (m0 > -1).astype(np.float32))


def test_mnn_ragged_thresholds_and_kpts(einx, synth):
    rng = np.random.default_rng(11)
    N, M, D = 300, 260, 64
    pairs = [synth.descriptor_pair(rng, N, M, D, 1.0) for _ in range(3)]
    n0 = np.array([300, 123, 1], dtype=np.int32)
    n1 = np.array([260, 260, 77], dtype=np.int32)
    k0 = rng.random((3, N, 3)).astype(np.float32)
    k1 = rng.random((3, M, 3)).astype(np.float32)
    for ratio, dist in [(None, None), (0.9, None), (None, 0.9), (0.95, 1.1)]:
        out = einx.mnn(cuda(np.stack([p[0] for p in pairs])), cuda(np.stack([p[1] for p in pairs])), cuda(n0), cuda(n1),
                       cuda(k0), cuda(k1), ratio, dist, True, "fp32")
        for b in range(3):
            a, c = pairs[b][0][: n0[b]], pairs[b][1][: n1[b]]
            ref = O.mnn_match(a, c, k0[b][: n0[b]], k1[b][: n1[b]], ratio, dist)
            m0 = out["matches0"][b].cpu().numpy()
            assert np.array_equal(m0[: n0[b]], ref["matches0"]) and (m0[n0[b]:] == -1).all()
            m1 = out["matches1"][b].cpu().numpy()
            assert np.array_equal(m1[: n1[b]], ref["matches1"]) and (m1[n1[b]:] == -1).all()
            nm = int(out["num_matches"][b])
            assert nm == len(ref["matched_kpts0"])
            assert np.array_equal(out["matched_kpts0"][b, :nm].cpu().numpy(), ref["matched_kpts0"])
            assert np.array_equal(out["matched_kpts1"][b, :nm].cpu().numpy(), ref["matched_kpts1"])


def test_matcher_module_contract(einx, synth):
    """Matchers.py:180-203 consumes these keys per sample at B=1; B>1 returns lists of matched keypoints."""
    rng = np.random.default_rng(2)
    a, b = synth.descriptor_pair(rng, 50, 60, 32, 1.0)
    k0, k1 = rng.random((50, 3)).astype(np.float32), rng.random((60, 3)).astype(np.float32)
    matcher = einx.NearestNeighborMatcher(False, False, True)
    out = matcher({"sparse_descriptors": cuda(a)[None], "sparse_positions": cuda(k0)[None]},
                  {"sparse_descriptors": cuda(b)[None], "sparse_positions": cuda(k1)[None]})
    for key in ("matches0", "matches1", "matching_scores0", "matching_scores1", "matched_kpts0", "matched_kpts1",
                "log_assignment"):
        assert key in out
    assert out["log_assignment"] is None and out["matched_kpts0"].dim() == 2
    out2 = matcher({"sparse_descriptors": cuda(np.stack([a, a])), "sparse_positions": cuda(np.stack([k0, k0]))},
                   {"sparse_descriptors": cuda(np.stack([b, b])), "sparse_positions": cuda(np.stack([k1, k1]))})
    assert isinstance(out2["matched_kpts0"], list) and len(out2["matched_kpts0"]) == 2
    assert torch.equal(out2["matched_kpts0"][1], out["matched_kpts0"])
    empty = matcher({"sparse_descriptors": torch.zeros((1, 0, 32), device=DEV), "sparse_positions": torch.zeros((1, 0, 3), device=DEV)},
                    {"sparse_descriptors": cuda(b)[None], "sparse_positions": cuda(k1)[None]})
    assert tuple(empty["matches1"].shape) == (1, 60) and (empty["matches1"] == -1).all()


# ------------------------------------------------------------------ whole path ------ #
@pytest.mark.parametrize("cfg_name,n_events", [("c2_ec_superpoint", 60_000), ("c1_mvsec_silk", 50_000)])
def test_pipeline_vs_oracle(einx, synth, cfg_name, n_events):
    c = synth.CONFIGS[cfg_name]
    Hp, Wp, _ = synth.padded_size(c["H"], c["W"], c["cell"])
    B = 3
    data = [synth.pair_inputs(cfg_name, s, n_events) for s in range(B)]
    cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"],
                          descriptor_mode=c["kind"], descriptor_scale=c["scale"], precision="fp32")
    pipe = einx.ExtractMatchPipeline(cfg)
    ev = einx.pack_events([d[0] for d in data])
    score = [cuda(np.concatenate([d[1][s][0] for d in data])) for s in range(2)]
    raw = [cuda(np.concatenate([d[1][s][1] for d in data])) for s in range(2)]
    out = pipe(tuple(t.to(DEV) for t in ev), score[0], raw[0], score[1], raw[1])
    torch.cuda.synchronize()
    for i, (events, sides) in enumerate(data):
        grid, p0, p1, m = O.pair_pipeline(events, c["bins"], c["H"], c["W"], sides[0][0].copy(), sides[0][1],
                                          sides[1][0].copy(), sides[1][1], "full" if c["kind"] == "gather" else "low",
                                          c["top_k"], c["scale"])
        g = out["voxel_grid"][i].cpu().numpy()
        assert (np.abs(g - grid) <= 1e-5 * np.maximum(np.abs(grid), 1.0)).mean() > 0.9999
        n0, n1 = int(out["counts0"][i]), int(out["counts1"][i])
        assert np.array_equal(out["keypoints0"][i, :n0].cpu().numpy(), p0)
        assert np.array_equal(out["keypoints1"][i, :n1].cpu().numpy(), p1)
        d0 = out["descriptors0"][i, :n0].cpu().numpy()
        d1 = out["descriptors1"][i, :n1].cpu().numpy()
        adjudicate(out["matches0"][i, :n0].cpu().numpy(), d0, d1)
        assert (out["matches0"][i, n0:] == -1).all()


def test_detect_tiled_large_map_redo_path(einx, synth):
    """Maps that no cluster holds (1280x720) run as row tiles with aprons; a tile whose NMS needs more rounds than its
    apron covers flags its image and the exact L2-resident kernel redoes that image only.  Image 0 is i.i.d. (6 rounds:
    tiled result stands), image 1 carries a long monotone ramp (one new maximum per round along it: dozens of rounds)."""
    rng = np.random.default_rng(77)
    Hp, Wp, k = 720, 1280, 4096
    m = synth.score_map(rng, 2, Hp, Wp)
    ramp = np.linspace(0.2, 0.9, 900, dtype=np.float32)
    m[1, 0, 300:340, 100:1000] = ramp[None, :] + 1e-3 * rng.random((40, 900), dtype=np.float32)
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 1.0, 4, 4, k, want_map=True)
    ref_in = m.copy()
    ref = O.prob_map_to_points_map(ref_in, 1.0, 4, 4, k)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    pos = O.prob_map_to_positions_with_prob(ref)
    for i in range(2):
        n = int(counts[i])
        assert n == len(pos[i]) and np.array_equal(kpts[i, :n].cpu().numpy(), pos[i])
    # both sides in one launch, keypoints only (the pipeline's call)
    (k0, c0), (k1, c1) = einx.detect_pair(cuda(m[:1]), cuda(m[1:]), 1.0, 4, 4, k)
    assert int(c0[0]) == len(pos[0]) and np.array_equal(k0[0, :len(pos[0])].cpu().numpy(), pos[0])
    assert int(c1[0]) == len(pos[1]) and np.array_equal(k1[0, :len(pos[1])].cpu().numpy(), pos[1])


def test_detect_tiled_large_map_with_mask(einx, synth):
    """The event mask (score[~mask] = 0, applied while loading) in the tiled form: every tile reads its own rows of
    the mask, only the band that owns a row writes the zeros back to `score`."""
    rng = np.random.default_rng(78)
    Hp, Wp, k = 720, 1280, 3000
    m = synth.score_map(rng, 1, Hp, Wp)
    mask = (rng.random((1, Hp, Wp)) < 0.7)
    mask[0, 200:260, :] = False  # a dead stripe across two tiles
    src = cuda(m)
    nms, kpts, counts = einx.detect(src, 1.0, 4, 4, k, mask=cuda(mask), want_map=True)
    ref_in = m * mask[:, None].astype(np.float32)
    ref = O.prob_map_to_points_map(ref_in, 1.0, 4, 4, k)   # (zeroes the border of ref_in in place)
    assert np.array_equal(src.cpu().numpy(), ref_in)
    assert np.array_equal(nms.cpu().numpy(), ref)
    pos = O.prob_map_to_positions_with_prob(ref)[0]
    n = int(counts[0])
    assert n == len(pos) and np.array_equal(kpts[0, :n].cpu().numpy(), pos)
