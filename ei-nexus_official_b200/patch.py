"""Swap the CUDA path into an imported copy of the reference.

    import einx
    einx.patch_reference()        # after `import core.modules...` / `import datasets.representations`

Replaces, by name, the functions SURVEY.md section 8 b lists in whichever of the reference's modules are
already in ``sys.modules`` (the extractors import them with ``from ..utils.detector_util import ...``, so
the extractor modules' own globals are patched too).

The kernels are forward-only (inference).  Two guards keep a patched checkout trainable and its data loaders
working:

* the replacements that sit on differentiable tensors (``logits_to_prob``, ``depth_to_space``, the two
  ``sparsify_*_descriptors``, ``sigmoid_log_double_softmax``) hand over to the reference's own function whenever
  autograd is recording and an input requires a gradient (``train_extractor.py`` reads ``pred_feats['score']``
  in core/loss/extractor_loss.py);
* the dataset-side functions (``events_to_voxel_grid`` and the other representations,
  ``draw_events_accumulation_image``) run inside ``Dataset.__getitem__`` -- in forked DataLoader workers under
  the reference's ``num_workers: 8`` -- where CUDA must not be touched: they are patched only on request
  (``datasets=True``) and even then fall back to the reference's function inside a worker process.
"""
from __future__ import annotations

import functools
import sys

import torch

from . import describe, detection as detect, match, metrics, voxel

_FUNCS = {
    "prob_map_to_points_map": detect.prob_map_to_points_map,
    "prob_map_to_positions_with_prob": detect.prob_map_to_positions_with_prob,
    "sparsify_full_resolution_descriptors": describe.sparsify_full_resolution_descriptors,
    "sparsify_low_resolution_descriptors": describe.sparsify_low_resolution_descriptors,
    "NearestNeighborMatcher": match.NearestNeighborMatcher,
    # adjacent rows (SURVEY.md section 8 f)
    "logits_to_prob": detect.logits_to_prob,
    "depth_to_space": detect.depth_to_space,
    "filter_matches": match.filter_matches,
    "sigmoid_log_double_softmax": match.sigmoid_log_double_softmax,
    "Repeatability": metrics.Repeatability,   # core/metrics/keypoints_metrics.py:52 (train_extractor.py:189, val_extractor.py:75)
}
# called from Dataset.__getitem__ (datasets/MVSEC.py:850-860, datasets/EC.py:300-306)
_DATASET_FUNCS = {
    "events_to_voxel_grid": voxel.events_to_voxel_grid,
    "draw_events_accumulation_image": voxel.draw_events_accumulation_image,
    "events_to_event_stack": voxel.events_to_event_stack,
    "events_to_time_surface": voxel.events_to_time_surface,
    "events_to_distance_map": voxel.events_to_distance_map,
}
# outputs the reference differentiates through when it trains the extractor / matcher
_GRAD_SENSITIVE = ("logits_to_prob", "depth_to_space", "sparsify_full_resolution_descriptors",
                   "sparsify_low_resolution_descriptors", "sigmoid_log_double_softmax")

_MODULE_HINTS = ("representations", "detector_util", "descriptor_util", "MNN", "EventExtractors",
                 "superpoint_extractor", "silk_extractor", "Matchers", "MVSEC", "EC", "visualize", "lightglue",
                 "keypoints_metrics", "train_extractor", "val_extractor")


def _needs_grad(args, kwargs) -> bool:
    def walk(v):
        if isinstance(v, torch.Tensor):
            return v.requires_grad
        if isinstance(v, (list, tuple)):
            return any(walk(e) for e in v)
        return False

    return torch.is_grad_enabled() and (any(walk(a) for a in args) or any(walk(a) for a in kwargs.values()))


def _in_loader_worker() -> bool:
    try:
        return torch.utils.data.get_worker_info() is not None
    except Exception:  # pragma: no cover
        return False


def _guarded(name, ours, original):
    """`ours`, except where only the reference's own function is correct (see the module docstring)."""
    if original is None or getattr(original, "_einx_guard", False):
        return ours
    if name in _GRAD_SENSITIVE:
        @functools.wraps(ours)
        def fn(*args, **kwargs):
            if _needs_grad(args, kwargs):
                return original(*args, **kwargs)
            return ours(*args, **kwargs)
    elif name in _DATASET_FUNCS:
        @functools.wraps(ours)
        def fn(*args, **kwargs):
            if _in_loader_worker():
                return original(*args, **kwargs)
            return ours(*args, **kwargs)
    else:
        return ours
    fn._einx_guard = True
    fn._einx_original = original
    return fn


_PATCHED = []  # (module, attribute, original) of everything patch_reference() replaced, for unpatch_reference()


def unpatch_reference():
    """Put the reference's own functions back (everything patch_reference() replaced in this process)."""
    n = 0
    while _PATCHED:
        m, attr, original = _PATCHED.pop()
        setattr(m, attr, original)
        n += 1
    return n


def patch_reference(modules=None, datasets: bool = False):
    """Returns {module name: [patched attribute, ...]}.

    ``datasets=True`` also replaces the dataset-side representation builders (use it with ``num_workers=0`` or a
    spawn start method: the replacements run on the GPU; inside a DataLoader worker they defer to the reference)."""
    done = {}
    funcs = dict(_FUNCS)
    if datasets:
        funcs.update(_DATASET_FUNCS)
    mods = modules if modules is not None else [
        m for name, m in list(sys.modules.items())
        if m is not None and name.rsplit(".", 1)[-1] in _MODULE_HINTS and not name.startswith(__package__)]
    for m in mods:
        for attr, fn in funcs.items():
            cur = getattr(m, attr, None)
            if cur is None or cur is fn or getattr(cur, "_einx_guard", False):
                continue
            if getattr(cur, "__module__", "").startswith(__package__):
                continue
            setattr(m, attr, _guarded(attr, fn, cur))
            _PATCHED.append((m, attr, cur))
            done.setdefault(m.__name__, []).append(attr)
    return done
