"""einx -- B200-native (sm_100a) extraction-and-matching hot path of EI-Nexus.

Event voxelisation, detection post-processing (NMS + top-k), descriptor sampling and MNN matching as
hand-written CUDA behind the C ABI of ``include/einx.h``; this package is the thin host layer that
keeps the reference's Python call surface.  No CPU fallback: everything raises without libeinx.so or
without an sm_100a device.
"""
from . import _lib
from ._lib import EinxError, context_for, contexts_of, launch_count
from .describe import sample, sparsify_full_resolution_descriptors, sparsify_low_resolution_descriptors
from .detection import (depth_to_space, detect, detect_pair, events_mask, logits_to_prob, logits_to_score, prob_map_to_points_map,
                        prob_map_to_positions_with_prob)
from .dist import gather_matches, pack_matches, shard_range
from .match import NearestNeighborMatcher, filter_matches, mnn, mnn_dense, sigmoid_log_double_softmax
from .metrics import Repeatability, gt_assign, pairwise_min_dist
from .patch import patch_reference, unpatch_reference
from .pipeline import CapturedStep, ExtractMatchPipeline, HostBatch, HostStreamer, PathConfig
from .voxel import (distance_map_device, draw_events_accumulation_image, event_stack_device, events_image_signed_device,
                    events_to_distance_map, events_image_device, events_to_event_stack,
                    events_to_time_surface, events_to_voxel_grid, pack_events, time_normalization, time_surface_device,
                    voxelize_batch, voxelize_device)

__all__ = [
    "EinxError", "context_for", "contexts_of", "launch_count", "events_to_voxel_grid", "time_normalization", "pack_events", "voxelize_batch",
    "voxelize_device", "detect", "detect_pair", "prob_map_to_points_map", "prob_map_to_positions_with_prob", "sample",
    "sparsify_full_resolution_descriptors", "sparsify_low_resolution_descriptors", "NearestNeighborMatcher",
    "mnn", "mnn_dense", "ExtractMatchPipeline", "CapturedStep", "HostBatch", "HostStreamer", "PathConfig", "patch_reference", "unpatch_reference", "shard_range", "pack_matches",
    "gather_matches", "logits_to_prob", "depth_to_space", "logits_to_score", "events_mask",
    "draw_events_accumulation_image", "events_image_device", "filter_matches", "events_to_event_stack",
    "events_to_time_surface", "event_stack_device", "time_surface_device", "sigmoid_log_double_softmax",
    "events_to_distance_map", "distance_map_device", "events_image_signed_device", "Repeatability", "gt_assign", "pairwise_min_dist",
]
