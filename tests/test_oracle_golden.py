"""The oracle (oracle/einx_oracle.py) against fixtures produced by the real reference."""
import numpy as np
import pytest

from oracle import einx_oracle as O


def voxel_close(got, ref, l1):
    """SURVEY.md section 8 a2 parity rule (pre-normalisation)."""
    return np.abs(got - ref) <= 1e-5 * np.maximum(np.abs(ref), l1) + 1e-30


def parse_tag(tag):
    name, rest = tag.split("_k") if "_k" in tag and "_t" in tag else (None, None)
    return name, rest


def test_voxel_matches_reference(golden):
    g = golden["voxel"]
    for ci in range(int(g["ncases"])):
        bins, H, W = (int(v) for v in g[f"c{ci}_shape"])
        ev = [g[f"c{ci}_{k}"] for k in "xytp"]
        raw, l1 = O.events_to_voxel_grid(*ev, bins, H, W, normalize=False, return_l1=True)
        assert voxel_close(raw, g[f"c{ci}_raw"], l1).all()
        assert ((raw != 0) == (g[f"c{ci}_raw"] != 0)).all()
        nrm = O.events_to_voxel_grid(*ev, bins, H, W, normalize=True)
        ref = g[f"c{ci}_norm"]
        assert (np.abs(nrm - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)).all()


def test_voxel_does_not_touch_inputs(golden):
    g = golden["voxel"]
    ev = [g[f"c0_{k}"].copy() for k in "xytp"]
    keep = [e.copy() for e in ev]
    O.events_to_voxel_grid(*ev, 5, 48, 64)
    for a, b in zip(ev, keep):
        assert np.array_equal(a, b)


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


def test_detect_bit_exact(golden):
    g = golden["detect"]
    for tag, name, r, b, k, thr in _detect_cases(g):
        src = g[f"{name}_in"].copy()
        nms = O.prob_map_to_points_map(src, thr, r, b, k)
        border_key = f"{tag}_border" if f"{tag}_border" in g.files else f"{name}_border"
        assert np.array_equal(src, g[border_key]), tag  # in-place border zeroing
        pos = O.prob_map_to_positions_with_prob(nms)
        for i, p in enumerate(pos):
            assert np.array_equal(p, g[f"{tag}_pos{i}"]), (tag, i)


def test_reference_nms_parity_property(golden):
    """utils_test.py:31-63 -- seed 0, rand(32,60,80): fast_nms result, and greedy == fast."""
    import torch

    g = golden["detect"]
    torch.manual_seed(0)
    inp = torch.rand((32, 60, 80)).numpy()
    got = O.prob_map_to_points_map(inp.copy(), 0.0, 4, 4, None)
    nz = np.argwhere(got != 0).astype(np.int32)
    assert np.array_equal(nz, g["parity_nz"])
    assert np.array_equal(got[tuple(nz.T)], g["parity_val"])
    bz = O.remove_border_points(inp[:2].copy(), 4)
    for i in range(2):
        assert np.array_equal(O.greedy_nms(bz[i], 4), got[i])


def test_reference_remove_border_property():
    """utils_test.py:17-29 -- an 8x8 map with border 4 becomes all zeros."""
    v = np.random.default_rng(0).random((8, 8)).astype(np.float32)
    assert not O.remove_border_points(v, 4).any()


def test_greedy_equals_fixpoint_on_ties():
    rng = np.random.default_rng(5)
    for r in (1, 2, 4):
        v = (np.round(rng.random((3, 40, 44)) * 6) / 6).astype(np.float32)
        fix, rounds = O.fast_nms(v, r, return_rounds=True)
        assert rounds >= 2
        for i in range(3):
            assert np.array_equal(O.greedy_nms(v[i], r), fix[i])


@pytest.mark.parametrize("n,k", [(92928, 2048), (44160, 1024), (921600, 8192), (921600, 512),
                                 (89960, 2048), (5120, 60), (100, 99), (100, 1)])
def test_topk_threshold_is_torch_quantile(n, k):
    import torch

    rng = np.random.default_rng(n + k)
    v = rng.random(n).astype(np.float32)
    v[rng.random(n) < 0.9] = 0  # mostly zeros, like an NMS'd map
    v[:7] = v[7:14]  # some exact ties
    q = (torch.tensor(n) - torch.tensor(k)) / n
    ref = torch.from_numpy(v)[None].quantile(q, dim=1, interpolation="midpoint")[0].item()
    assert O.topk_threshold(v, k) == np.float32(ref)


def test_sampling_matches_reference(golden):
    g = golden["sample"]
    pos = [g["pos0"], g["pos1"]]
    full = O.sparsify_full_resolution_descriptors(g["raw_full"], pos, 1.41, True)
    low = O.sparsify_low_resolution_descriptors(g["raw_low"], pos, (48, 64), 1.0, True)
    for i in range(2):
        assert np.abs(full[i] - g[f"full{i}"]).max() < 2e-6
        assert np.abs(low[i] - g[f"low{i}"]).max() < 2e-6
    empty = O.sparsify_low_resolution_descriptors(g["raw_low"][:1], [np.zeros((0, 3), np.float32)], (48, 64))
    assert empty[0].shape == (0, 32)


def test_mnn_bit_exact(golden):
    g = golden["mnn"]
    for ci in range(int(g["ncases"])):
        ratio, dist = (float(v) or None for v in g[f"c{ci}_cfg"])
        out = O.mnn_match(g[f"c{ci}_d0"], g[f"c{ci}_d1"], g[f"c{ci}_k0"], g[f"c{ci}_k1"],
                          ratio, dist, True, return_dense=True)
        for key in ("matches0", "matches1", "matching_scores0", "matching_scores1",
                    "matched_kpts0", "matched_kpts1"):
            assert np.array_equal(out[key], g[f"c{ci}_{key}"]), (ci, key)
        assert out["matches0"].dtype == np.int64
        assert np.abs(out["log_assignment"] - g[f"c{ci}_log_assignment"]).max() < 1e-5
        assert (out["matches0"] > -1).sum() == (out["matches1"] > -1).sum()


def test_padder_sizes():
    assert O.padder_sizes(260, 346, 8) == (3, 3, 2, 2)
    assert O.padder_sizes(180, 240, 8) == (0, 0, 2, 2)
    assert O.padder_sizes(260, 346, 1) == (0, 0, 0, 0)
