"""Detection post-processing -- host side of einx_detect.

Keeps the reference surface of ``core/modules/utils/detector_util.py``:
``prob_map_to_points_map`` (:80-135) and ``prob_map_to_positions_with_prob`` (:451-484).
One kernel launch does border removal, the NMS fixpoint, the top-k threshold and the ordered
keypoint compaction; the dense map the first function returns carries the keypoints so the second
call costs no second pass.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def max_keypoints(Hp: int, Wp: int, nms_dist: int) -> int:
    """Upper bound on NMS survivors: one per (r+1)x(r+1) cell."""
    if nms_dist <= 0:
        return Hp * Wp
    return ((Hp + nms_dist) // (nms_dist + 1)) * ((Wp + nms_dist) // (nms_dist + 1))


@torch.no_grad()
def detect(score: torch.Tensor, prob_thresh: float, nms_dist: int, border_dist: int, top_k: Optional[int],
           mask: Optional[torch.Tensor] = None, want_map: bool = False, kcap: Optional[int] = None):
    """Fused detection on a (B, 1, H, W) / (B, H, W) fp32 CUDA score map.

    ``score`` gets its border frame (and ``mask == 0`` pixels) zeroed in place, as the reference
    does.  Returns ``(nms_map | None, kpts (B, kcap, 3), counts (B,) int32)`` -- rows are
    ``(y + .5, x + .5, prob)`` in raster order, rows >= counts[b] are unspecified.
    """
    if score.dtype != torch.float32 or not score.is_cuda:
        raise _lib.EinxError("detect: score must be a float32 CUDA tensor (there is no CPU fallback)")
    if not score.is_contiguous():
        raise ValueError("detect: score must be contiguous (it is modified in place)")
    if score.dim() == 4:
        if score.shape[1] != 1:
            raise ValueError("detect: expected (B, 1, H, W)")
        B, _, Hp, Wp = score.shape
    elif score.dim() == 3:
        B, Hp, Wp = score.shape
    else:
        raise ValueError("detect: expected (B, 1, H, W) or (B, H, W)")
    dev = score.device
    ctx = _lib.context_for(dev)
    k = int(top_k) if top_k else 0
    if kcap is None:
        # the top-k threshold only bounds the count by k when prob_thresh cannot undercut it
        # (thr = min(thr_k, prob_thresh)); probabilities never exceed 1
        bound = max_keypoints(Hp, Wp, nms_dist)
        kcap = min(k, bound) if (k > 0 and prob_thresh >= 1.0) else bound
    kcap = max(int(kcap), 1)
    kpts = torch.empty((B, kcap, 3), dtype=torch.float32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    nms_map = torch.empty((B, Hp, Wp), dtype=torch.float32, device=dev) if want_map else None
    mptr = None
    if mask is not None:
        m8 = mask.reshape(B, Hp, Wp).to(torch.uint8).contiguous()
        mptr = _lib.ptr(m8)
    rc = ctx.lib.einx_detect(ctx.handle, _lib.ptr(score), mptr, B, Hp, Wp, int(nms_dist), int(border_dist),
                             float(prob_thresh), k, _lib.ptr(nms_map), _lib.ptr(kpts), kcap, _lib.ptr(counts),
                             _lib.stream_of(dev))
    ctx.check(rc, "einx_detect")
    return nms_map, kpts, counts


class _PointsMap(torch.Tensor):
    """The dense ``nms`` tensor, remembering the keypoint rows found in the same launch."""

    @staticmethod
    def wrap(t, kpts, counts):
        out = t.as_subclass(_PointsMap)
        out._einx_kpts = (kpts, counts)
        return out


def prob_map_to_points_map(prob_map: torch.Tensor, prob_thresh: float = 0.015, nms_dist: int = 4,
                           border_dist: int = 4, use_fast_nms: bool = True, top_k: int = None):
    """Drop-in for ``detector_util.py:80-135``: returns the (B, H, W) NMS'd, thresholded map.

    ``use_fast_nms`` is accepted for signature parity; both reference NMS variants define the same
    fixpoint (utils_test.py:31-63) and this is the one kernel for it.
    """
    if isinstance(prob_thresh, torch.Tensor):
        prob_thresh = float(prob_thresh)
    view = prob_map if prob_map.is_contiguous() else None
    work = prob_map if view is not None else prob_map.contiguous()
    nms, kpts, counts = detect(work, prob_thresh, nms_dist, border_dist, top_k, want_map=True)
    if view is None:  # keep the in-place border zeroing observable on a strided caller tensor
        prob_map.copy_(work)
    return _PointsMap.wrap(nms, kpts, counts)


def unpack_rows(rows: torch.Tensor, counts: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """(B, cap, C) padded rows + counts -> tuple of (N_i, C) tensors (one host sync)."""
    n = counts.tolist()
    return tuple(rows[i, : min(c, rows.shape[1])].clone() for i, c in enumerate(n))


def prob_map_to_positions_with_prob(prob_map: torch.Tensor, threshold: float = 0.0, ordering: str = "yx"):
    """Drop-in for ``detector_util.py:451-484``: tuple of (N_i, 3) rows (y+.5, x+.5, prob)."""
    cached = getattr(prob_map, "_einx_kpts", None)
    if cached is not None and threshold == 0.0:
        kpts, counts = cached
    else:
        # stand-alone compaction of `prob_map > threshold` (no NMS, no border): same kernel, r = 0
        pm = prob_map.as_subclass(torch.Tensor).clone().contiguous()
        if pm.dim() == 4:
            pm = pm.squeeze(1)
        _, kpts, counts = detect(pm, float(threshold), 0, 0, None)
    out = unpack_rows(kpts, counts)
    if ordering == "xy":
        out = tuple(torch.cat((p[:, [1, 0]], p[:, 2:]), dim=1) for p in out)
    return out
