"""Swap the CUDA path into an imported copy of the reference.

    import einx
    einx.patch_reference()        # after `import core.modules...` / `import datasets.representations`

Replaces, by name, the functions SURVEY.md section 8 b lists in whichever of the reference's modules are
already in ``sys.modules`` (the extractors import them with ``from ..utils.detector_util import ...``, so
the extractor modules' own globals are patched too).
"""
from __future__ import annotations

import sys

from . import describe, detection as detect, match, voxel

_FUNCS = {
    "events_to_voxel_grid": voxel.events_to_voxel_grid,
    "prob_map_to_points_map": detect.prob_map_to_points_map,
    "prob_map_to_positions_with_prob": detect.prob_map_to_positions_with_prob,
    "sparsify_full_resolution_descriptors": describe.sparsify_full_resolution_descriptors,
    "sparsify_low_resolution_descriptors": describe.sparsify_low_resolution_descriptors,
    "NearestNeighborMatcher": match.NearestNeighborMatcher,
    # adjacent rows (SURVEY.md section 8 f)
    "draw_events_accumulation_image": voxel.draw_events_accumulation_image,
    "events_to_event_stack": voxel.events_to_event_stack,
    "events_to_time_surface": voxel.events_to_time_surface,
    "logits_to_prob": detect.logits_to_prob,
    "depth_to_space": detect.depth_to_space,
    "filter_matches": match.filter_matches,
    "sigmoid_log_double_softmax": match.sigmoid_log_double_softmax,
}

_MODULE_HINTS = ("representations", "detector_util", "descriptor_util", "MNN", "EventExtractors",
                 "superpoint_extractor", "silk_extractor", "Matchers", "MVSEC", "EC", "visualize", "lightglue")


def patch_reference(modules=None):
    """Returns {module name: [patched attribute, ...]}."""
    done = {}
    mods = modules if modules is not None else [
        m for name, m in list(sys.modules.items())
        if m is not None and name.rsplit(".", 1)[-1] in _MODULE_HINTS and not name.startswith(__package__)]
    for m in mods:
        for attr, fn in _FUNCS.items():
            if hasattr(m, attr) and getattr(m, attr) is not fn:
                setattr(m, attr, fn)
                done.setdefault(m.__name__, []).append(attr)
    return done
