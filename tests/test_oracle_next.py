"""Oracle restatements of the rows adjacent to the hot path (SURVEY.md section 8 f) against the fixtures
made by the reference's own functions (tests/golden/make_golden_next.py).  CPU only."""
import numpy as np

from oracle import einx_oracle as O


def test_events_image_and_mask_match_reference(golden):
    g = golden["next"]
    for ci in range(int(g["img_ncases"])):
        H, W, cell = (int(v) for v in g[f"img{ci}_shape"])
        img = O.draw_events_accumulation_image(g[f"img{ci}_x"], g[f"img{ci}_y"], H, W)
        assert img.dtype == np.uint8 and np.array_equal(img, g[f"img{ci}_out"])
        assert np.array_equal(O.events_mask(img, cell), g[f"img{ci}_mask"])


def test_detector_head_matches_reference(golden):
    g = golden["next"]
    p = O.logits_to_prob(g["head_logits65"])
    np.testing.assert_allclose(p, g["head_prob65"], rtol=2e-6, atol=1e-9)
    # the shuffle is pure data movement: exact on the reference's own probabilities
    assert np.array_equal(O.depth_to_space(g["head_prob65"], 8), g["head_score65"])
    np.testing.assert_allclose(O.logits_to_prob(g["head_logits1"]), g["head_prob1"], rtol=2e-6, atol=1e-9)
    assert np.array_equal(O.depth_to_space(g["head_prob1"], 1), g["head_score1"])
    np.testing.assert_allclose(O.depth_to_space(O.logits_to_prob(g["head_logits17"]), 4), g["head_score17"], rtol=2e-6, atol=1e-9)


def test_filter_matches_matches_reference(golden):
    g = golden["next"]
    for ci in range(int(g["fm_ncases"])):
        m0, m1, s0, s1 = O.filter_matches(g[f"fm{ci}_scores"], float(g[f"fm{ci}_th"]))
        assert np.array_equal(m0, g[f"fm{ci}_m0"]) and np.array_equal(m1, g[f"fm{ci}_m1"])
        np.testing.assert_allclose(s0, g[f"fm{ci}_s0"], rtol=2e-6)
        np.testing.assert_allclose(s1, g[f"fm{ci}_s1"], rtol=2e-6)


def test_event_stack_and_time_surface_match_reference(golden):
    g = golden["repr"]
    for ci in range(int(g["ncases"])):
        bins, H, W = (int(v) for v in g[f"c{ci}_shape"])
        ev = [g[f"c{ci}_{k}"] for k in "xytp"]
        assert np.array_equal(O.events_to_event_stack(*ev, bins, H, W), g[f"c{ci}_stack"])
        assert np.array_equal(O.events_to_time_surface(*ev, bins, H, W), g[f"c{ci}_surface"])


def test_log_double_softmax_matches_reference(golden):
    g = golden["lg"]
    for ci in range(int(g["ncases"])):
        ref = g[f"c{ci}_scores"]
        out = O.sigmoid_log_double_softmax(g[f"c{ci}_sim"], g[f"c{ci}_z0"], g[f"c{ci}_z1"])
        assert out.shape == ref.shape and out.dtype == np.float32
        # fp32 exp / log / summation order: 1e-6 of the magnitude (measured 4e-7)
        assert np.all(np.abs(out - ref) <= 1e-6 * np.maximum(1.0, np.abs(ref))), ci
        # the reference's own filter_matches output on its own matrix is reproduced from the oracle's matrix
        m0, m1, s0, s1 = O.filter_matches(ref, float(g[f"c{ci}_th"]))
        assert np.array_equal(m0, g[f"c{ci}_m0"]) and np.array_equal(m1, g[f"c{ci}_m1"])


def test_distance_map_matches_reference(golden):
    """events_to_distance_map (representations.py:215-248): the oracle's closed-form chamfer distance against the
    reference's cv.distanceTransform output.  The IPP build of opencv 4.13 accumulates the fp32 weights along the path:
    1 ulp on dense windows, 4.2e-7 relative at distances of tens of pixels; bar 1e-6 (the plain C build of OpenCV, with
    16.16 fixed-point weights, is 2e-6 from either)."""
    g = golden["repr"]
    cases = [(f"c{ci}", ) for ci in range(int(g["ncases"]))] + [("sparse",)]
    for (key,) in cases:
        bins, H, W = (int(v) for v in g[f"{key}_shape"])
        out = O.events_to_distance_map(*[g[f"{key}_{k}"] for k in "xytp"], bins, H, W)
        ref = g[f"{key}_distance"]
        empty = ref > 1e30  # bins without events: FLT_MAX everywhere
        assert np.array_equal(out > 1e30, empty)
        np.testing.assert_allclose(out[~empty], ref[~empty], rtol=1e-6, atol=0)
    assert (g["sparse_distance"] > 1e30).any() and g["sparse_distance"][g["sparse_distance"] < 1e30].max() > 20


def test_repeatability_matches_reference(golden):
    """Repeatability.update_one (keypoints_metrics.py:57-128): the value exactly (counts over counts), the two min
    reductions of the N x M distance matrix within fp32 rounding of the homography warp."""
    g = golden["metrics"]
    for ci in range(int(g["ncases"])):
        H, W, thr, xy = (int(v) for v in g[f"c{ci}_cfg"])
        value, min1, min2 = O.repeatability(g[f"c{ci}_p1"], g[f"c{ci}_p2"], (H, W), (H, W), g[f"c{ci}_hom"], thr, "xy" if xy else "yx")
        ref = float(g[f"c{ci}_value"])
        if np.isnan(ref):  # no point on either side: the reference reports nothing
            assert value is None
        else:
            assert np.float32(value) == np.float32(ref), ci  # the reference divides fp32 count tensors
        assert min1.shape == g[f"c{ci}_min_over_1"].shape and min2.shape == g[f"c{ci}_min_over_2"].shape
        np.testing.assert_allclose(min1, g[f"c{ci}_min_over_1"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(min2, g[f"c{ci}_min_over_2"], rtol=1e-5, atol=1e-4)


def test_gt_assign_matches_reference(golden):
    """gt_matches_from_pose_depth (gt_generation.py:96-126) run whole in the reference; the oracle restates the
    N x M block on the inputs the function itself computed (projections, visibility, validity)."""
    g = golden["gt_assign"]
    for ci in range(int(g["ncases"])):
        pos_th, neg_th = (float(v) for v in g[f"c{ci}_th"])
        a, m0, m1 = O.gt_assign(g[f"c{ci}_kp0"], g[f"c{ci}_kp1"], g[f"c{ci}_kp0_1"], g[f"c{ci}_kp1_0"], g[f"c{ci}_visible0"],
                                g[f"c{ci}_visible1"], g[f"c{ci}_valid0"], g[f"c{ci}_valid1"], pos_th, neg_th)
        M = g[f"c{ci}_kp1"].shape[1]
        want = np.unpackbits(g[f"c{ci}_assignment"], axis=-1)[..., :M].astype(bool)
        assert np.array_equal(m0, g[f"c{ci}_m0"]) and np.array_equal(m1, g[f"c{ci}_m1"])
        assert np.array_equal(a, want)
