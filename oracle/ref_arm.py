"""The reference's own CPU implementation of the path, loaded from oracle/_ref/ (see oracle/make_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline` leg and `--impl reference` arm and by the tests
that pin the numpy port (oracle/einx_oracle.py) to it.  The functions are called in the reference's own order and
with its own semantics, including the per-sample matcher loop of core/modules/Matchers.py:192-203 and the
per-keypoint Python loop of core/modules/matchers/MNN.py:119-127.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "core/modules/matchers/MNN.py"))


def load():
    """(detector_util, descriptor_util, util, MNN, representations) of the reference; cached."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref/ is missing: run `python oracle/make_ref.py` where the reference checkout exists")

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules.setdefault(name, m)

    stub("core", f"{REF}/core")
    stub("core.modules", f"{REF}/core/modules")
    stub("core.modules.utils", f"{REF}/core/modules/utils")
    stub("core.modules.matchers", f"{REF}/core/modules/matchers")

    def ld(name, path):
        if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
            return sys.modules[name]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    det = ld("core.modules.utils.detector_util", f"{REF}/core/modules/utils/detector_util.py")
    desc = ld("core.modules.utils.descriptor_util", f"{REF}/core/modules/utils/descriptor_util.py")
    util = ld("core.modules.utils.util", f"{REF}/core/modules/utils/util.py")
    mnn = ld("core.modules.matchers.MNN", f"{REF}/core/modules/matchers/MNN.py")
    rep = ld("einx_ref_representations", f"{REF}/datasets/representations.py")
    _mods = (det, desc, util, mnn, rep)
    return _mods


_extractors = None


def load_extractors():
    """core/modules/event_extractors/EventExtractors.py of the reference (VGGExtractor: SuperPoint type, cell 8;
    VGGExtractorNP: SiLK type, cell 1) with its net/ modules; `kornia` (imported, not used on this path) is stubbed."""
    global _extractors
    if _extractors is not None:
        return _extractors
    load()
    if not os.path.isfile(f"{REF}/core/modules/event_extractors/EventExtractors.py"):
        raise RuntimeError("oracle/_ref/ has no extractor modules: re-run `python oracle/make_ref.py`")
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))
    for name in ("core.modules.net", "core.modules.event_extractors"):
        m = types.ModuleType(name)
        m.__path__ = [f"{REF}/{name.replace('.', '/')}"]
        sys.modules.setdefault(name, m)

    def ld(name, path):
        if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
            return sys.modules[name]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for leaf in ("vgg", "conv", "pointnet", "backbone", "detector_head", "descriptor_head"):
        ld(f"core.modules.net.{leaf}", f"{REF}/core/modules/net/{leaf}.py")
    _extractors = ld("core.modules.event_extractors.EventExtractors", f"{REF}/core/modules/event_extractors/EventExtractors.py")
    return _extractors


def pair_pipeline(ev, bins, H, W, score0, raw0, score1, raw1, kind, top_k, scale, nms_dist=4, border=4, prob_thresh=1.0):
    """One event-image pair through the reference functions (same arguments as einx_oracle.pair_pipeline).

    voxel grid: datasets/representations.py:66-124; per side prob_map_to_points_map + prob_map_to_positions_with_prob
    (detector_util.py:80-135, :451-484) and sparsify_*_descriptors (descriptor_util.py:50-128); then the frozen
    matcher's per-sample call (Matchers.py:192-203) of NearestNeighborMatcher.forward (MNN.py:43-140)."""
    import contextlib
    import io

    import torch

    det, desc, _, mnn, rep = load()
    with torch.no_grad():
        events = {k: np.array(v, copy=True) for k, v in ev.items()}  # the reference mutates its argument
        grid = rep.events_to_voxel_grid(events, (bins, H, W))
        feats = []
        for score, raw in ((score0, raw0), (score1, raw1)):
            s = torch.from_numpy(np.array(score, copy=True))  # borders are zeroed in place
            r = torch.from_numpy(np.ascontiguousarray(raw))
            nms = det.prob_map_to_points_map(s, prob_thresh=prob_thresh, nms_dist=nms_dist, border_dist=border,
                                             use_fast_nms=True, top_k=top_k)
            pos = det.prob_map_to_positions_with_prob(nms, threshold=0.0, ordering="yx")
            if kind == "full":
                d = desc.sparsify_full_resolution_descriptors(r, pos, scale_factor=torch.tensor(scale), normalize=True)
            else:
                d = desc.sparsify_low_resolution_descriptors(r, pos, tuple(s.shape[-2:]), scale_factor=torch.tensor(scale),
                                                             normalize=True)
            feats.append({"sparse_positions": pos, "sparse_descriptors": d})
        matcher = mnn.NearestNeighborMatcher(ratio_thresh=None, distance_thresh=None, mutual_check=True)
        out = None
        for i in range(len(feats[0]["sparse_positions"])):  # Matchers.py:192-203: one call per sample, B = 1
            f0 = {k: feats[0][k][i][None, ...] for k in feats[0]}
            f1 = {k: feats[1][k][i][None, ...] for k in feats[1]}
            with contextlib.redirect_stdout(io.StringIO()):  # ("No keypoints" is printed, not raised)
                out = matcher(f0, f1)
    m = {"matches0": out["matches0"][0].numpy(), "matches1": out["matches1"][0].numpy(),
         "matching_scores0": out["matching_scores0"][0].numpy()}
    return grid.numpy(), feats[0]["sparse_positions"][0].numpy(), feats[1]["sparse_positions"][0].numpy(), m
