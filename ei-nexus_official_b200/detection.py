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
    k = int(top_k) if top_k else 0
    if kcap is None:
        # the top-k threshold only bounds the count by k when prob_thresh cannot undercut it
        # (thr = min(thr_k, prob_thresh)); probabilities never exceed 1
        bound = max_keypoints(Hp, Wp, nms_dist)
        kcap = min(k, bound) if (k > 0 and prob_thresh >= 1.0) else bound
    kcap = max(int(kcap), 1)
    m8 = None if mask is None else mask.reshape(B, Hp, Wp).to(torch.uint8).contiguous()
    # registered PyTorch op over einx_detect (csrc/torch/einx_torch.cpp): current stream, caching allocator
    kpts, counts, nms_map = _lib.ops().detect(score.view(B, Hp, Wp), m8, int(nms_dist), int(border_dist), float(prob_thresh),
                                              k, kcap, bool(want_map))
    if not want_map:
        nms_map = None
    return nms_map, kpts, counts


@torch.no_grad()
def detect_pair(score0: torch.Tensor, score1: torch.Tensor, prob_thresh: float, nms_dist: int, border_dist: int,
                top_k: Optional[int], mask0: Optional[torch.Tensor] = None, mask1: Optional[torch.Tensor] = None,
                kcap: Optional[int] = None):
    """:func:`detect` for the two sides of a batch of pairs in one launch (einx_detect_pair): two (B, 1, H, W) maps of
    one shape -> ``((kpts0, counts0), (kpts1, counts1))``.  Both maps are border-zeroed in place."""
    for s in (score0, score1):
        if s.dtype != torch.float32 or not s.is_cuda:
            raise _lib.EinxError("detect_pair: score maps must be float32 CUDA tensors (there is no CPU fallback)")
        if not s.is_contiguous():
            raise ValueError("detect_pair: score maps must be contiguous (they are modified in place)")
    if score0.shape != score1.shape or score0.device != score1.device:
        raise ValueError("detect_pair: the two sides must have one shape and one device")
    if score0.dim() == 4 and score0.shape[1] != 1:
        raise ValueError("detect_pair: expected (B, 1, H, W)")
    B, Hp, Wp = score0.shape[0], score0.shape[-2], score0.shape[-1]
    k = int(top_k) if top_k else 0
    if kcap is None:
        bound = max_keypoints(Hp, Wp, nms_dist)
        kcap = min(k, bound) if (k > 0 and prob_thresh >= 1.0) else bound
    kcap = max(int(kcap), 1)
    m8 = [None if m is None else m.reshape(B, Hp, Wp).to(torch.uint8).contiguous() for m in (mask0, mask1)]
    k0, c0, k1, c1 = _lib.ops().detect_pair(score0, score1, m8[0], m8[1], int(nms_dist), int(border_dist), float(prob_thresh),
                                            k, kcap)
    kp, cn = (k0, k1), (c0, c1)
    return (kp[0], cn[0]), (kp[1], cn[1])


class _PointsMap(torch.Tensor):
    """The dense ``nms`` tensor, remembering the keypoint rows found in the same launch."""

    @staticmethod
    def wrap(t, kpts, counts):
        out = t.as_subclass(_PointsMap)
        out._einx_kpts = (kpts, counts, t._version)  # the rows describe this version of the map only
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
    if cached is not None and threshold == 0.0 and cached[2] == prob_map._version:  # unmodified since the launch
        kpts, counts = cached[0], cached[1]
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


# --------------------------------------------------------------------------------------------- #
# adjacent rows (SURVEY.md section 8 f): detector head post-processing and the event mask
# --------------------------------------------------------------------------------------------- #
HEAD_SCORE, HEAD_PROB, HEAD_SHUFFLE = 0, 1, 2


def _head(x: torch.Tensor, cell: int, mode: int) -> torch.Tensor:
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.EinxError("detector head: expected a float32 CUDA tensor (there is no CPU fallback)")
    if x.dim() != 4:
        raise ValueError("detector head: expected (B, C, Hc, Wc)")
    x = x.contiguous()
    B, C, Hc, Wc = x.shape
    dev = x.device
    ctx = _lib.context_for(dev)
    shape = (B, C, Hc, Wc) if mode == HEAD_PROB else (B, 1, Hc * cell, Wc * cell)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    rc = ctx.lib.einx_logits_to_score(ctx.handle, _lib.ptr(x), B, C, Hc, Wc, int(cell), mode, _lib.ptr(out),
                                      ctx.stream)
    ctx.check(rc, "einx_logits_to_score")
    return out


@torch.no_grad()
def logits_to_score(logits: torch.Tensor, cell_size: int = 8) -> torch.Tensor:
    """``depth_to_space(logits_to_prob(logits), cell_size)`` in one kernel (detector_util.py:18-77):
    (B, cell^2+1, Hc, Wc) logits -> (B, 1, Hc*cell, Wc*cell) scores; the probability tensor is never
    written."""
    return _head(logits, cell_size, HEAD_SCORE)


@torch.no_grad()
def logits_to_prob(logits: torch.Tensor, channel_dim: int = 1) -> torch.Tensor:
    """Drop-in for ``detector_util.py:18-39`` (softmax over channels, or 1/(1+exp(-x)) for one channel)."""
    if channel_dim not in (1, -3):
        raise ValueError("logits_to_prob: channel_dim must be 1 (the only layout the reference uses)")
    return _head(logits, 1, HEAD_PROB)


@torch.no_grad()
def depth_to_space(prob: torch.Tensor, cell_size: int = 8, channel_dim: int = 1) -> torch.Tensor:
    """Drop-in for ``detector_util.py:42-77`` (drop the dustbin, pixel-shuffle by ``cell_size``)."""
    if channel_dim not in (1, -3):
        raise ValueError("depth_to_space: channel_dim must be 1 (the only layout the reference uses)")
    if cell_size > 1:
        assert prob.shape[1] == cell_size * cell_size + 1
    else:
        assert prob.shape[1] == 1
        return prob
    return _head(prob, cell_size, HEAD_SHUFFLE)


@torch.no_grad()
def events_mask(events_image: torch.Tensor, cell_size: int = 1) -> torch.Tensor:
    """(B, H, W) / (B, 1, H, W) event accumulation image -> (B, 1, Hp, Wp) bool score mask.

    ``events_image > 0`` (train_extractor.py:225), ``Padder.pad`` with constant zeros
    (core/modules/utils/util.py:17-32) and the 3x3 box filter + ``> 0`` of
    core/modules/event_extractors/EventExtractors.py:357-363, in one kernel.  Pass the result as
    ``mask`` to :func:`detect`, which applies ``score[~mask] = 0`` (:374-375) while loading."""
    if not events_image.is_cuda:
        raise _lib.EinxError("events_mask: expected a CUDA tensor (there is no CPU fallback)")
    img = events_image
    if img.dim() == 4:
        if img.shape[1] != 1:
            raise ValueError("events_mask: expected (B, 1, H, W)")
        img = img[:, 0]
    if img.dtype != torch.uint8:
        img = (img > 0).to(torch.uint8)
    img = img.contiguous()
    B, H, W = img.shape
    hp = (((H // cell_size) + 1) * cell_size - H) % cell_size
    wp = (((W // cell_size) + 1) * cell_size - W) % cell_size
    Hp, Wp = H + hp, W + wp
    dev = img.device
    ctx = _lib.context_for(dev)
    mask = torch.empty((B, 1, Hp, Wp), dtype=torch.uint8, device=dev)
    rc = ctx.lib.einx_mask_dilate(ctx.handle, _lib.ptr(img), B, H, W, hp // 2, wp // 2, Hp, Wp, _lib.ptr(mask),
                                  ctx.stream)
    ctx.check(rc, "einx_mask_dilate")
    return mask.view(torch.bool)
