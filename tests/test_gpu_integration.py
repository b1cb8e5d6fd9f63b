"""The reference's REAL extractor modules (oracle/_ref/core/modules/event_extractors/EventExtractors.py: conv backbone,
heads, Padder.pad / unpad_positions, filter_sparse_feats, mapping_positions -- copied verbatim by oracle/make_ref.py)
run on the GPU twice on the same weights and inputs: once as they are, once after ``einx.patch_reference()`` replaced
logits_to_prob / depth_to_space / prob_map_to_points_map / prob_map_to_positions_with_prob / sparsify_*_descriptors in
the module's globals.  Then the frozen matcher loop of core/modules/Matchers.py:192-203 over the reference's
NearestNeighborMatcher and over the drop-in.  (-m gpu; nothing here reads the reference checkout.)"""
import sys

import numpy as np
import pytest
import torch

from oracle import ref_arm

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref/ not made")]

DEV = torch.device("cuda", 0)


@pytest.fixture(scope="module")
def einx():
    import einx as m

    assert torch.cuda.is_available()
    m.context_for(DEV)
    return m


def _build(ee, kind, D, top_k):
    torch.manual_seed(7)
    cls = ee.VGGExtractor if kind == "sp" else ee.VGGExtractorNP
    model = cls(in_channels=5, feat_channels=32, descriptor_dim=D, nms_radius=4, detection_top_k=top_k, detection_threshold=1.0,
                remove_borders=4, ordering="yx", descriptor_scale_factor=1.0 if kind == "sp" else 1.41)
    return model.to(DEV).eval()


@pytest.mark.parametrize("kind,D,H,W,top_k", [("sp", 256, 90, 122, 150), ("silk", 128, 77, 101, 300)])
def test_patched_extractor_forward_matches_unpatched(einx, kind, D, H, W, top_k):
    ee = ref_arm.load_extractors()
    model = _build(ee, kind, D, top_k)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn((2, 5, H, W), generator=g).to(DEV)
    mask = (torch.rand((2, 1, H, W), generator=g) > 0.2).to(DEV)
    torch.backends.cudnn.deterministic = True
    with torch.no_grad():
        ref = model(x, score_mask=mask)
    det_mod, desc_mod = sys.modules["core.modules.utils.detector_util"], sys.modules["core.modules.utils.descriptor_util"]
    done = einx.patch_reference([ee, det_mod, desc_mod])
    try:
        assert "prob_map_to_points_map" in done[ee.__name__] and "logits_to_prob" in done[ee.__name__]
        with torch.no_grad():
            out = model(x, score_mask=mask)
    finally:
        assert einx.unpatch_reference() > 0
    assert ee.prob_map_to_points_map is det_mod.prob_map_to_points_map  # originals are back
    # same conv outputs (same device, same weights): the post-processing is what differs between the two runs
    assert torch.equal(ref["logits"], out["logits"]) and torch.equal(ref["raw_descriptors"], out["raw_descriptors"])
    torch.testing.assert_close(out["score"], ref["score"], rtol=2e-6, atol=2e-7)
    if not torch.equal(out["score"], ref["score"]):
        pytest.skip("softmax differs in the last ulp on this map: keypoint sets are only comparable on identical scores "
                    "(covered bit-exactly from fixed score maps in test_gpu_parity.py)")
    assert torch.equal(out["nms"], ref["nms"])
    for i in range(2):
        p, q = out["sparse_positions"][i], ref["sparse_positions"][i]
        assert p.shape == q.shape and torch.equal(p, q), i          # bit-exact keypoints through pad / unpad / filter
        assert 0 < p.shape[0] <= top_k
        torch.testing.assert_close(out["sparse_descriptors"][i], ref["sparse_descriptors"][i], rtol=0, atol=2e-6)


def test_frozen_matcher_loop_with_drop_in(einx):
    """core/modules/Matchers.py:192-203: one matcher call per sample on dict slices; the loop reads matches0/1,
    matching_scores0/1, matched_kpts0/1 and log_assignment from the result."""
    ee = ref_arm.load_extractors()
    mnn_mod = sys.modules["core.modules.matchers.MNN"]
    model = _build(ee, "silk", 128, 200)
    g = torch.Generator(device="cpu").manual_seed(5)
    x0 = torch.randn((2, 5, 64, 80), generator=g).to(DEV)
    x1 = x0 + 0.05 * torch.randn((2, 5, 64, 80), generator=g).to(DEV)
    with torch.no_grad():
        f0, f1 = model(x0), model(x1)
    keys = ("sparse_positions", "sparse_descriptors")
    results = {}
    for name, matcher in (("ref", mnn_mod.NearestNeighborMatcher(None, None, True)),
                          ("einx", einx.NearestNeighborMatcher(None, None, True, precision="fp32"))):
        out = {k: [] for k in ("matches0", "matches1", "matching_scores0", "matching_scores1", "matched_kpts0", "matched_kpts1",
                               "log_assignment")}
        for i in range(len(f0["sparse_positions"])):
            a = {k: f0[k][i][None, ...] for k in keys}
            b = {k: f1[k][i][None, ...] for k in keys}
            with torch.no_grad():
                o = matcher(a, b)
            out = {k: out[k] + [o[k]] for k in out}
        results[name] = out
    for i in range(2):
        r, e = results["ref"], results["einx"]
        assert torch.equal(r["matches0"][i], e["matches0"][i]) and torch.equal(r["matches1"][i], e["matches1"][i])
        torch.testing.assert_close(e["matching_scores0"][i], r["matching_scores0"][i], rtol=0, atol=2e-6)
        assert torch.equal(r["matched_kpts0"][i], e["matched_kpts0"][i]) and torch.equal(r["matched_kpts1"][i], e["matched_kpts1"][i])
        assert int((r["matches0"][i] > -1).sum()) > 10
