"""CPU oracle for the EI-Nexus extraction-and-matching hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference
algorithm (ZhonghuaYi/EI-Nexus_official); it is the *checker* for the CUDA path,
never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
package ``ei-nexus_official_b200`` must never import anything from ``oracle/``.

Parity pinning: every function below is checked in ``tests/test_oracle_golden.py``
against fixtures under ``tests/golden/`` that were produced by importing the
reference's own Python functions in the build container
(``tests/golden/make_golden.py``) and against the two property tests the
reference carries for this path
(``core/modules/image_extractors/silk/backbones/superpoint/utils_test.py:17-63``).

Reference citations are relative to the reference repository root.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# a1 / a2  event representation  (datasets/representations.py:8-22, 66-124)
# --------------------------------------------------------------------------- #
def time_normalization(t):
    """t <- (t - t[0]) / (t[-1] - t[0] + 1e-8) in fp64 (representations.py:19-20)."""
    t = np.asarray(t, dtype=np.float64)
    t = t - t[0]
    return t / (t[-1] + 1e-8)


def voxel_event_terms(x, y, t, p, bins):
    """fp32 per-event quantities exactly as representations.py:73-89 builds them.

    Returns (xf, yf, tn, pol): fp32 x, fp32 y, fp32 normalised time in
    [0, bins-1], polarity with ``p < 1 -> -1``.
    """
    t64 = time_normalization(t)
    xf = np.asarray(x).astype(F32)
    yf = np.asarray(y).astype(F32)
    pf = np.asarray(p).astype(F32)
    tf = t64.astype(F32)
    # (bins-1) * (t - t0) / (tN - t0), evaluated left to right in fp32 (:80-81)
    tn = (F32(bins - 1) * (tf - tf[0])) / (tf[-1] - tf[0])
    pol = np.where(pf < F32(1), F32(-1), pf).astype(F32)  # :88-89
    return xf, yf, tn.astype(F32), pol


def events_to_voxel_grid(x, y, t, p, bins, H, W, normalize=True, return_l1=False):
    """Trilinear event splat + non-zero mean/std normalisation.

    Follows datasets/representations.py:66-124.  The 8 corner contributions are
    accumulated in fp64 and rounded once to fp32 (the reference accumulates in
    fp32 with a thread-dependent order), so this is the
    order-free value the parity rule ``|d| <= 1e-5 * max(|ref|, sum|w|)`` is
    stated against.  ``return_l1`` also returns sum|w| per cell.
    """
    xf, yf, tn, pol = voxel_event_terms(x, y, t, p, bins)
    x0 = np.trunc(xf).astype(np.int32)  # .int() truncates toward zero (:83-85)
    y0 = np.trunc(yf).astype(np.int32)
    t0 = np.trunc(tn).astype(np.int32)
    ncell = bins * H * W
    acc = np.zeros(ncell, dtype=np.float64)
    l1 = np.zeros(ncell, dtype=np.float64)
    one = F32(1)
    for xl in (x0, x0 + 1):  # corner order of :91-93
        for yl in (y0, y0 + 1):
            for tl in (t0, t0 + 1):
                ok = (xl < W) & (xl >= 0) & (yl < H) & (yl >= 0) & (tl >= 0) & (tl < bins)
                w = pol * (one - np.abs(xl.astype(F32) - xf))
                w = w * (one - np.abs(yl.astype(F32) - yf))
                w = (w * (one - np.abs(tl.astype(F32) - tn))).astype(F32)
                idx = (H * W) * tl.astype(np.int64) + W * yl.astype(np.int64) + xl.astype(np.int64)
                acc += np.bincount(idx[ok], weights=w[ok].astype(np.float64), minlength=ncell)
                if return_l1:
                    l1 += np.bincount(idx[ok], weights=np.abs(w[ok]).astype(np.float64), minlength=ncell)
    grid = acc.astype(F32).reshape(bins, H, W)
    if normalize:
        grid = normalize_nonzero(grid)
    if return_l1:
        return grid, l1.astype(F32).reshape(bins, H, W)
    return grid


def normalize_nonzero(grid):
    """Mean / unbiased std over cells != 0, applied to those cells (:114-122)."""
    grid = np.array(grid, dtype=F32, copy=True)
    m = grid != 0
    n = int(m.sum())
    if n > 0:
        vals = grid[m].astype(np.float64)
        mean = F32(vals.mean())
        # torch.std is unbiased; one element gives nan, which fails `std > 0`
        std = F32(vals.std(ddof=1)) if n > 1 else F32(np.nan)
        if std > 0:
            grid[m] = (grid[m] - mean) / std
        else:
            grid[m] = grid[m] - mean
    return grid


# --------------------------------------------------------------------------- #
# a3 / a4  border removal + iterative NMS  (detector_util.py:138-164, 243-337)
# --------------------------------------------------------------------------- #
def remove_border_points(v, border):
    """Zero a ``border``-wide frame in place on (..., H, W) (detector_util.py:151-162)."""
    if border > 0:
        v[..., :, :border] = 0
        v[..., :, -border:] = 0
        v[..., :border, :] = 0
        v[..., -border:, :] = 0
    return v


def _window_maxes(v, r):
    """Max over the raster-earlier and raster-later halves of the (2r+1)^2 window.

    v: (B, H, W) fp32, zero padding outside the image (F.unfold padding, :289-295).
    """
    B, H, W = v.shape
    P = np.zeros((B, H + 2 * r, W + 2 * r), dtype=v.dtype)
    P[:, r:r + H, r:r + W] = v
    row_full = P[:, :, 0:W].copy()
    for dx in range(1, 2 * r + 1):
        np.maximum(row_full, P[:, :, dx:dx + W], out=row_full)
    earlier = np.zeros_like(v)
    later = np.zeros_like(v)
    for dy in range(0, r):  # rows above the centre
        np.maximum(earlier, row_full[:, dy:dy + H], out=earlier)
    for dy in range(r + 1, 2 * r + 1):  # rows below
        np.maximum(later, row_full[:, dy:dy + H], out=later)
    for dx in range(0, r):  # same row, left
        np.maximum(earlier, P[:, r:r + H, dx:dx + W], out=earlier)
    for dx in range(r + 1, 2 * r + 1):  # same row, right
        np.maximum(later, P[:, r:r + H, dx:dx + W], out=later)
    return earlier, later


def local_maxima(v, r):
    """Centre == first-occurrence argmax of its zero-padded window (:298-299).

    argmax returns the first maximal slot in window raster order, so the centre
    wins iff it is strictly greater than every earlier slot and >= every later
    one; a zero centre never wins (slot 0 of the window ties or beats it).
    """
    earlier, later = _window_maxes(v, r)
    return (v > 0) & (v > earlier) & (v >= later)


def _dilate(mask, r):
    B, H, W = mask.shape
    P = np.zeros((B, H + 2 * r, W + 2 * r), dtype=bool)
    P[:, r:r + H, r:r + W] = mask
    rows = P[:, :, 0:W].copy()
    for dx in range(1, 2 * r + 1):
        rows |= P[:, :, dx:dx + W]
    out = rows[:, 0:H].copy()
    for dy in range(1, 2 * r + 1):
        out |= rows[:, dy:dy + H]
    return out


def fast_nms(v, r, return_rounds=False):
    """Iterative NMS to the fixpoint, with the reference's batch-wide stop rule.

    detector_util.py:286-335: find local maxima, stop when their count over the
    whole batch is unchanged, otherwise zero every pixel that has a local
    maximum in its window (the maximum itself excepted).  v: (B, H, W) >= 0.
    """
    v = np.array(v, dtype=F32, copy=True)
    if r == 0:
        return (v, 0) if return_rounds else v
    count = None
    rounds = 0
    while True:
        lm = local_maxima(v, r)
        c = int(lm.sum())
        if c == count:
            break
        count = c
        v[_dilate(lm, r) & ~lm] = 0
        rounds += 1
    return (v, rounds) if return_rounds else v


def greedy_nms(v, r):
    """Greedy NMS in (value desc, raster asc) order; zeros never selected.

    Small-case cross-check only (pure Python loop).  SURVEY.md section 8 a4: the
    iterative fixpoint equals this ordering of detector_util.py:167-240.
    """
    v = np.array(v, dtype=F32, copy=True)
    H, W = v.shape
    flat = v.ravel()
    order = np.lexsort((np.arange(flat.size), -flat.astype(np.float64)))
    out = np.zeros_like(v)
    alive = v > 0
    for idx in order:
        yy, xx = divmod(int(idx), W)
        if not alive[yy, xx]:
            continue
        out[yy, xx] = v[yy, xx]
        alive[max(0, yy - r):yy + r + 1, max(0, xx - r):xx + r + 1] = False
    return out


# --------------------------------------------------------------------------- #
# a5  top-k threshold  (detector_util.py:108-133)
# --------------------------------------------------------------------------- #
def topk_ranks(n, k):
    """fp32 emulation of ``q = (n-k)/n`` and ``rank = q*(n-1)`` (:113-124).

    torch divides an int64 tensor by a Python int in fp32 and ``quantile``
    multiplies q by (n-1) in the input dtype, so both steps round to fp32.
    """
    q = F32(n - k) / F32(n)
    rank = F32(q * F32(n - 1))
    return int(np.floor(rank)), int(np.ceil(rank))


def topk_threshold(flat, k):
    """'midpoint' quantile of one image's n values, zeros included (:108-124)."""
    flat = np.asarray(flat, dtype=F32).ravel()
    n = flat.size
    if k >= n:
        return F32(0)
    lo, hi = topk_ranks(n, k)
    s = np.sort(flat)
    a, b = s[lo], s[hi]
    # torch.lerp(a, b, 0.5) takes the |w| >= 0.5 branch: b - (b - a) * (1 - w)
    return F32(b - F32(F32(b - a) * F32(0.5)))


def prob_map_to_points_map(prob_map, prob_thresh=0.015, nms_dist=4, border_dist=4, top_k=None):
    """Border removal -> fast_nms -> top-k / prob threshold (detector_util.py:80-135).

    prob_map: (B, 1, H, W) or (B, H, W) fp32; border zeroing happens IN PLACE on
    the caller's array, like the reference.  Returns the (B, H, W) nms map.
    """
    remove_border_points(prob_map, border_dist)
    v = prob_map.reshape(prob_map.shape[0], prob_map.shape[-2], prob_map.shape[-1])
    v = fast_nms(v, nms_dist)
    B = v.shape[0]
    thr = np.full((B,), F32(prob_thresh), dtype=F32)
    if top_k:
        for i in range(B):
            thr[i] = min(topk_threshold(v[i], int(top_k)), F32(prob_thresh))
    return np.where(v > thr[:, None, None], v, F32(0)).astype(F32)


# --------------------------------------------------------------------------- #
# a6  positions  (detector_util.py:451-484)
# --------------------------------------------------------------------------- #
def prob_map_to_positions_with_prob(nms, threshold=0.0, ordering="yx"):
    """Raster-order (y+.5, x+.5, prob) rows per image (:470-484)."""
    nms = nms.reshape(nms.shape[0], nms.shape[-2], nms.shape[-1])
    out = []
    for i in range(nms.shape[0]):
        yy, xx = np.nonzero(nms[i] > threshold)
        pos = np.stack([yy, xx], axis=1).astype(F32) + F32(0.5)
        if ordering == "xy":
            pos = pos[:, ::-1]
        out.append(np.concatenate([pos, nms[i][yy, xx][:, None]], axis=1).astype(F32))
    return tuple(out)


def padder_sizes(h, w, p):
    """(w0, w1, h0, h1) of core/modules/utils/util.py:9-15."""
    hp = (((h // p) + 1) * p - h) % p
    wp = (((w // p) + 1) * p - w) % p
    return (wp // 2, wp - wp // 2, hp // 2, hp - hp // 2)


# --------------------------------------------------------------------------- #
# a7  descriptor sampling  (descriptor_util.py:21-28, 50-128)
# --------------------------------------------------------------------------- #
def normalize_descriptors(d, scale=1.0, normalize=True):
    """scale * d / max(||d||_2, 1e-12) over dim 1 (descriptor_util.py:21-28)."""
    d = np.asarray(d, dtype=F32)
    if not normalize:
        return (F32(scale) * d).astype(F32)
    nrm = np.sqrt((d.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(F32)
    return (F32(scale) * (d / np.maximum(nrm, F32(1e-12)))).astype(F32)


def sparsify_full_resolution_descriptors(raw, positions, scale=1.0, normalize=True):
    """Integer gather raw[i, :, floor(y), floor(x)] + L2 (descriptor_util.py:50-71)."""
    out = []
    for i, pos in enumerate(positions):
        yy = np.floor(pos[:, 0]).astype(np.int64)
        xx = np.floor(pos[:, 1]).astype(np.int64)
        d = raw[i][:, yy, xx].T
        out.append(normalize_descriptors(d, scale, normalize))
    return tuple(out)


def sparsify_low_resolution_descriptors(raw, positions, image_size, scale=1.0, normalize=True):
    """Bilinear sampling of the coarse map + L2 (descriptor_util.py:74-128).

    Restates F.grid_sample(bilinear, zeros padding, align_corners=False) on the
    grid ``2*((pos-0.5)/(size-1)) - 1``: the un-normalisation is
    ``((g+1)*size_in - 1)/2``; taps outside the coarse map contribute 0.
    """
    Hp, Wp = F32(image_size[0]), F32(image_size[1])
    out = []
    for i, pos in enumerate(positions):
        C, Hc, Wc = raw[i].shape
        n = pos.shape[0]
        if n == 0:
            out.append(np.zeros((0, C), dtype=F32))
            continue
        py = pos[:, 0].astype(F32) - F32(0.5)
        px = pos[:, 1].astype(F32) - F32(0.5)
        gy = F32(2) * (py / (Hp - F32(1))) - F32(1)
        gx = F32(2) * (px / (Wp - F32(1))) - F32(1)
        iy = ((gy + F32(1)) * F32(Hc) - F32(1)) / F32(2)
        ix = ((gx + F32(1)) * F32(Wc) - F32(1)) / F32(2)
        y0 = np.floor(iy)
        x0 = np.floor(ix)
        wy1 = (iy - y0).astype(F32)
        wx1 = (ix - x0).astype(F32)
        wy0 = ((y0 + F32(1)) - iy).astype(F32)
        wx0 = ((x0 + F32(1)) - ix).astype(F32)
        y0 = y0.astype(np.int64)
        x0 = x0.astype(np.int64)
        d = np.zeros((n, C), dtype=F32)
        # tap order nw, ne, sw, se as in ATen's grid_sampler_2d
        for (yy, xx, w) in ((y0, x0, wx0 * wy0), (y0, x0 + 1, wx1 * wy0),
                            (y0 + 1, x0, wx0 * wy1), (y0 + 1, x0 + 1, wx1 * wy1)):
            ok = (yy >= 0) & (yy < Hc) & (xx >= 0) & (xx < Wc)
            yc = np.clip(yy, 0, Hc - 1)
            xc = np.clip(xx, 0, Wc - 1)
            tap = raw[i][:, yc, xc].T * w[:, None].astype(F32)
            d += np.where(ok[:, None], tap, F32(0)).astype(F32)
        out.append(normalize_descriptors(d, scale, normalize))
    return out


# --------------------------------------------------------------------------- #
# a8  mutual nearest neighbour matching  (core/modules/matchers/MNN.py:11-140)
# --------------------------------------------------------------------------- #
def find_nn(sim, ratio_thresh=None, distance_thresh=None):
    """First-index argmax per row with optional ratio / distance tests (MNN.py:11-22)."""
    n, m = sim.shape
    ind = np.argmax(sim, axis=1).astype(np.int64)  # first occurrence, like topk(1)
    best = sim[np.arange(n), ind]
    ok = np.ones(n, dtype=bool)
    d0 = F32(2) * (F32(1) - best)
    if ratio_thresh:
        tmp = sim.copy()
        tmp[np.arange(n), ind] = -np.inf
        second = tmp.max(axis=1)
        d1 = F32(2) * (F32(1) - second)
        ok &= d0 <= F32(ratio_thresh ** 2) * d1
    if distance_thresh:
        ok &= d0 <= F32(distance_thresh ** 2)
    return np.where(ok, ind, -1).astype(np.int64)


def mutual_check(m0, m1):
    """Keep i<->j only when each is the other's nearest neighbour (MNN.py:25-32)."""
    i0 = np.arange(m0.size)
    i1 = np.arange(m1.size)
    loop0 = m1[np.where(m0 > -1, m0, 0)] if m1.size else np.zeros_like(m0)
    loop1 = m0[np.where(m1 > -1, m1, 0)] if m0.size else np.zeros_like(m1)
    m0n = np.where((m0 > -1) & (i0 == loop0), m0, -1)
    m1n = np.where((m1 > -1) & (i1 == loop1), m1, -1)
    return m0n.astype(np.int64), m1n.astype(np.int64)


def mnn_match(desc0, desc1, kpts0=None, kpts1=None, ratio_thresh=None, distance_thresh=None,
              mutual=True, return_dense=False, sim_dtype=np.float32):
    """One pair of NearestNeighborMatcher.forward (MNN.py:88-129).

    desc0 (N, D), desc1 (M, D).  ``sim_dtype=np.float64`` gives the tie-free
    adjudicator used for near-tie rows (MNN index parity).
    """
    d0 = np.asarray(desc0, dtype=sim_dtype)
    d1 = np.asarray(desc1, dtype=sim_dtype)
    sim = d0 @ d1.T
    m0 = find_nn(sim, ratio_thresh, distance_thresh)
    m1 = find_nn(sim.T, ratio_thresh, distance_thresh)
    if mutual:
        m0, m1 = mutual_check(m0, m1)
    out = {
        "matches0": m0, "matches1": m1,
        "matching_scores0": (m0 > -1).astype(F32), "matching_scores1": (m1 > -1).astype(F32),
    }
    if kpts0 is not None:
        keep = m0 > -1
        out["matched_kpts0"] = np.asarray(kpts0)[keep]
        out["matched_kpts1"] = np.asarray(kpts1)[m0[keep]]
    if return_dense:
        out["similarity"] = sim
        n, m = sim.shape
        la = np.zeros((n + 1, m + 1), dtype=sim.dtype)
        s64 = sim.astype(np.float64)
        lr = s64.max(1, keepdims=True) + np.log(np.exp(s64 - s64.max(1, keepdims=True)).sum(1, keepdims=True))
        lc = s64.max(0, keepdims=True) + np.log(np.exp(s64 - s64.max(0, keepdims=True)).sum(0, keepdims=True))
        la[:n, :m] = (2 * s64 - lr - lc).astype(sim.dtype)
        out["log_assignment"] = la
    return out


# --------------------------------------------------------------------------- #
# whole path for one event-image pair (the unit bench.py counts)
# --------------------------------------------------------------------------- #
def extract_side(score, raw, kind, nms_dist, border, top_k, prob_thresh, scale):
    """detect -> positions -> sample for one side; score (1,H,W) is border-zeroed in place."""
    nms = prob_map_to_points_map(score, prob_thresh, nms_dist, border, top_k)
    pos = prob_map_to_positions_with_prob(nms)
    if kind == "full":
        desc = sparsify_full_resolution_descriptors(raw, pos, scale, True)
    else:
        desc = sparsify_low_resolution_descriptors(raw, pos, score.shape[-2:], scale, True)
    return pos, desc


def pair_pipeline(ev, bins, H, W, score0, raw0, score1, raw1, kind, top_k, scale,
                  nms_dist=4, border=4, prob_thresh=1.0):
    """voxelise -> [detect -> sample] x2 -> MNN for ONE pair (SURVEY.md section 8 d)."""
    grid = events_to_voxel_grid(ev["x"], ev["y"], ev["t"], ev["p"], bins, H, W, True)
    p0, d0 = extract_side(score0, raw0, kind, nms_dist, border, top_k, prob_thresh, scale)
    p1, d1 = extract_side(score1, raw1, kind, nms_dist, border, top_k, prob_thresh, scale)
    m = mnn_match(d0[0], d1[0], p0[0], p1[0])
    return grid, p0[0], p1[0], m


# --------------------------------------------------------------------------- #
# adjacent rows (SURVEY.md section 8 f)
# --------------------------------------------------------------------------- #
def draw_events_accumulation_image(x, y, H, W):
    """Event count image, min-max scaled to uint8 (datasets/visualize.py:23-49, dict branch).

    One count per event at (int(y), int(x)) -- Python ``int()`` truncates toward zero (:35) -- then
    ``(c - min) / (max - min) * 255`` in fp64 (:45), clipped (:46), ``astype(uint8)`` (:48).
    In-range events only (the reference indexes a numpy array, so negative indices would wrap)."""
    img = np.zeros((H, W), dtype=np.float64)
    iy = np.trunc(np.asarray(y, dtype=np.float64)).astype(np.int64)
    ix = np.trunc(np.asarray(x, dtype=np.float64)).astype(np.int64)
    np.add.at(img, (iy, ix), 1.0)
    mn, mx = img.min(), img.max()
    if mx == mn:  # 0/0 -> NaN -> uint8 cast (0 on x86); the CUDA path defines this case as zeros
        return np.zeros((H, W), dtype=np.uint8)
    img = (img - mn) / (mx - mn) * 255
    img[img > 255] = 255
    return img.astype(np.uint8)


def events_mask(events_image, cell):
    """``events_image > 0`` (train_extractor.py:225) -> constant-zero padding to a multiple of ``cell``
    (core/modules/utils/util.py:9-32, bool tensors) -> 3x3 box filter, ``> 0``
    (core/modules/event_extractors/EventExtractors.py:357-363) = 3x3 dilation of the padded mask."""
    m = np.asarray(events_image) > 0
    H, W = m.shape[-2:]
    w0, w1, h0, h1 = padder_sizes(H, W, cell)
    Hp, Wp = H + h0 + h1, W + w0 + w1
    pad = [(0, 0)] * (m.ndim - 2) + [(h0, h1), (w0, w1)]
    m = np.pad(m, pad, mode="constant")
    z = np.pad(m, [(0, 0)] * (m.ndim - 2) + [(1, 1), (1, 1)], mode="constant")
    out = np.zeros_like(m)
    for dy in range(3):
        for dx in range(3):
            out |= z[..., dy:dy + Hp, dx:dx + Wp]
    return out


def logits_to_prob(logits):
    """Channel softmax, or 1 / (1 + exp(-x)) for a single channel (detector_util.py:18-39); fp32."""
    x = np.asarray(logits, dtype=F32)
    if x.shape[1] == 1:
        return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)
    e = np.exp(x - x.max(axis=1, keepdims=True)).astype(F32)
    return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


def depth_to_space(prob, cell):
    """Drop the dustbin channel and pixel-shuffle by ``cell`` (detector_util.py:42-77):
    out[b, 0, h*cell + i, w*cell + j] = prob[b, i*cell + j, h, w]."""
    p = np.asarray(prob)
    if cell == 1:
        assert p.shape[1] == 1
        return p
    B, C, Hc, Wc = p.shape
    assert C == cell * cell + 1
    p = p[:, :cell * cell].reshape(B, cell, cell, Hc, Wc)
    return np.ascontiguousarray(p.transpose(0, 3, 1, 4, 2).reshape(B, 1, Hc * cell, Wc * cell))


def filter_matches(scores, th):
    """Matches from a LightGlue log-assignment matrix (core/modules/matchers/lightglue.py:402-418):
    row / column argmax of scores[:, :-1, :-1] (first index on ties), mutual check, exp, threshold."""
    s = np.asarray(scores, dtype=F32)[:, :-1, :-1]
    m0, m1 = s.argmax(2), s.argmax(1)
    max0 = s.max(2)
    B = s.shape[0]
    bi = np.arange(B)[:, None]
    mutual0 = np.arange(s.shape[1])[None] == m1[bi, m0]
    mutual1 = np.arange(s.shape[2])[None] == m0[bi, m1]
    ms0 = np.where(mutual0, np.exp(max0), F32(0)).astype(F32)
    ms1 = np.where(mutual1, ms0[bi, m1], F32(0)).astype(F32)
    valid0 = mutual0 & (ms0 > F32(th))
    valid1 = mutual1 & valid0[bi, m1]
    return np.where(valid0, m0, -1), np.where(valid1, m1, -1), ms0, ms1


def _log_sigmoid(x):
    """torch's logsigmoid: min(x, 0) - log1p(exp(-|x|)); fp32."""
    x = np.asarray(x, dtype=F32)
    return (np.minimum(x, F32(0)) - np.log1p(np.exp(-np.abs(x)))).astype(F32)


def _log_softmax(x, axis):
    """(x - max) - log(sum exp(x - max)) in fp32; the sum of the fp32 exponentials is accumulated in fp64 and rounded
    once (torch's vectorised fp32 sum is within an ulp of that; numpy's fp32 sum along a strided axis is sequential and
    drifts by 3e-6 over 1000 terms)."""
    x = np.asarray(x, dtype=F32)
    s = x - x.max(axis=axis, keepdims=True)
    return (s - np.log(np.exp(s).sum(axis=axis, keepdims=True, dtype=np.float64).astype(F32))).astype(F32)


def sigmoid_log_double_softmax(sim, z0, z1):
    """LightGlue log-assignment matrix (core/modules/matchers/lightglue.py:365-377): row + column log-softmax of the
    similarities plus the matchability certainties; unmatched row / column logsigmoid(-z); corner 0.  fp32."""
    sim = np.asarray(sim, dtype=F32)
    b, m, n = sim.shape
    z0 = np.asarray(z0, dtype=F32).reshape(b, m)
    z1 = np.asarray(z1, dtype=F32).reshape(b, n)
    cert = _log_sigmoid(z0)[:, :, None] + _log_sigmoid(z1)[:, None, :]
    out = np.zeros((b, m + 1, n + 1), dtype=F32)
    if m and n:  # torch's log_softmax over an empty dimension returns an empty tensor; numpy's max raises
        out[:, :m, :n] = (_log_softmax(sim, 2) + _log_softmax(sim, 1)) + cert
    out[:, :-1, -1] = _log_sigmoid(-z0)
    out[:, -1, :-1] = _log_sigmoid(-z1)
    return out


def _bin_slices(tn, nbins):
    """np.searchsorted bounds of the reference's per-bin loops (representations.py:44-47, 196-199):
    bin i owns events with i*dt <= t <= i*dt + dt, both ends inclusive."""
    dt = 1.0 / nbins
    for i in range(nbins):
        t0 = i * dt
        t1 = t0 + dt
        yield i, np.searchsorted(tn, t0, side="left"), np.searchsorted(tn, t1, side="right")


def events_to_event_stack(x, y, t, p, bins, H, W):
    """Per-bin sum of 2*int(p) - 1 at (int(y), int(x)) (datasets/representations.py:177-214); fp32 (bins, H, W)."""
    tn = time_normalization(t)
    x0, y0 = np.asarray(x).astype(np.int32), np.asarray(y).astype(np.int32)
    p0 = 2 * np.asarray(p).astype(np.int32) - 1
    out = np.zeros((bins, H, W), dtype=np.float32)
    for i, a, b in _bin_slices(tn, bins):
        xs, ys, ps = x0[a:b], y0[a:b], p0[a:b]
        ok = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)  # :209
        np.add.at(out[i], (ys[ok], xs[ok]), ps[ok].astype(np.float32))
    return out


def events_to_time_surface(x, y, t, p, bins, H, W):
    """Latest normalised time per (2*bin + int(p), int(y), int(x)) over bins // 2 time bins
    (datasets/representations.py:25-63; the fancy assignment at :57 keeps the last, i.e. latest, event).
    In-range events with polarity 0/1 only (anything else indexes another channel in the reference)."""
    nb = bins // 2
    tn = time_normalization(t)
    x0, y0, p0 = (np.asarray(v).astype(np.int32) for v in (x, y, p))
    out = np.zeros((bins, H, W), dtype=np.float32)
    for i, a, b in _bin_slices(tn, nb):
        out[2 * i + p0[a:b], y0[a:b], x0[a:b]] = tn[a:b]
    return out


# OpenCV's 3x3 chamfer weights for DIST_L2 (axial, diagonal), as fp32
_CHAMFER_A, _CHAMFER_B = np.float32(0.955), np.float32(1.3693)


def chamfer_3x3(event_map):
    """cv.distanceTransform(1 - event_map, cv.DIST_L2, 3): the 3x3 chamfer distance to the nearest event pixel.

    The two-pass raster algorithm computes the shortest 8-connected path with weights (a, b); on an unobstructed grid
    that is min over event pixels of  b * min(|dx|, |dy|) + a * (max(|dx|, |dy|) - min(|dx|, |dy|))  (b < 2a).  Evaluated
    here in fp64 and rounded once.  OpenCV's result depends on its build: the IPP path accumulates fp32 weights along the
    path (1 ulp on dense maps, 4.2e-7 relative at distances of tens of pixels: measured on opencv 4.13), the plain C path uses 16.16 fixed-point weights (2e-6
    relative off).  A map without events yields FLT_MAX everywhere (opencv 4.13)."""
    m = np.asarray(event_map) > 0
    H, W = m.shape
    if not m.any():
        return np.full((H, W), np.finfo(np.float32).max, dtype=F32)
    ys, xs = np.nonzero(m)
    out = np.empty((H, W), dtype=F32)
    a, b = float(_CHAMFER_A), float(_CHAMFER_B)
    for y in range(H):  # row by row keeps the (W, n_events) temporaries small
        dy = np.abs(y - ys)[None, :]
        dx = np.abs(np.arange(W)[:, None] - xs[None, :])
        mn, mx = np.minimum(dx, dy), np.maximum(dx, dy)
        out[y] = (b * mn + a * (mx - mn)).min(axis=1).astype(F32)
    return out


def events_to_distance_map(x, y, t, p, bins, H, W):
    """Per time bin, the chamfer distance of every pixel to the nearest event of the bin
    (datasets/representations.py:215-248).  Bin i owns events with i/bins <= t <= (i + 1)/bins, both ends inclusive
    (searchsorted 'left' ... 'right'); coordinates truncate like astype(int32); polarity is not used."""
    tn = time_normalization(t)
    x0, y0 = np.asarray(x).astype(np.int32), np.asarray(y).astype(np.int32)
    ct = 1 / bins
    out = np.zeros((bins, H, W), dtype=F32)
    for i in range(bins):
        a_, b_ = np.searchsorted(tn, i * ct, side="left"), np.searchsorted(tn, (i + 1) * ct, side="right")
        em = np.zeros((H, W), dtype=np.uint8)
        em[y0[a_:b_], x0[a_:b_]] = 1
        out[i] = chamfer_3x3(em)
    return out


def _warp_points(pts, hom):
    """core/metrics/util.py:5-39: (2, n) points through a 3 x 3 homography, fp32."""
    h = np.asarray(hom, dtype=F32)
    p = np.vstack([np.asarray(pts, dtype=F32)[:2], np.ones((1, pts.shape[1]), dtype=F32)])
    q = (h @ p).astype(F32)
    return np.vstack([q[0] / q[2], q[1] / q[2]]).astype(F32)


def _keep_true_points(pts, hom, shape):
    """core/metrics/util.py:43-104: keep the points whose warp lands inside (H, W)."""
    w = _warp_points(pts, hom)
    mask = (w[0] >= 0) & (w[0] < shape[1]) & (w[1] >= 0) & (w[1] < shape[0])
    return pts[:, mask]


def repeatability(points1, points2, img1_shape, img2_shape, homography, distance_thresh=3, ordering="xy"):
    """Repeatability.update_one (core/metrics/keypoints_metrics.py:57-128).  points (N, >=2) rows in `ordering`; returns
    (value or None, min over side 1 per side-2 point, min over side 2 per warped side-1 point): the N x M Euclidean
    distance matrix of :110-113 reduced along both axes, counts of minima <= threshold over the number of points."""
    sel = [0, 1] if ordering == "xy" else [1, 0]
    q1 = np.asarray(points1, dtype=F32).T[sel]
    q2 = np.asarray(points2, dtype=F32).T[sel]
    h = np.asarray(homography, dtype=F32)
    q2 = _keep_true_points(q2, np.linalg.inv(h).astype(F32), img1_shape)
    q1 = _keep_true_points(q1, h, img2_shape)
    wp = _warp_points(q1, h).T
    q2 = q2.T
    n, m = wp.shape[0], q2.shape[0]
    d = wp[:, None, :] - q2[None, :, :]
    norm = np.sqrt((d.astype(np.float64) ** 2).sum(-1)).astype(F32)  # torch's CPU norm accumulates in fp64
    min1 = norm.min(0) if n else np.zeros(0, F32)
    min2 = norm.min(1) if m else np.zeros(0, F32)
    c1 = int((min1 <= distance_thresh).sum()) if n else 0
    c2 = int((min2 <= distance_thresh).sum()) if m else 0
    value = (c1 + c2) / (n + m) if n + m > 0 else None
    return value, min1, min2


def gt_assign(kp0, kp1, kp0_1, kp1_0, visible0, visible1, valid0, valid1, pos_th=3, neg_th=5):
    """core/geometry/gt_generation.py:96-126 (gt_matches_from_pose_depth between `project` and the epipolar pass),
    one batch item at a time.  kp* (B, N|M, 2) fp32 in the order the function indexes them; returns
    (assignment (B, N, M) bool, m0 (B, N) int64, m1 (B, M) int64) with -1 = unmatched, -2 = ignore."""
    kp0, kp1, kp0_1, kp1_0 = (np.asarray(a, dtype=F32) for a in (kp0, kp1, kp0_1, kp1_0))
    B, N, M = kp0.shape[0], kp0.shape[1], kp1.shape[1]
    assignment = np.zeros((B, N, M), dtype=bool)
    m0 = np.full((B, N), -1, dtype=np.int64)
    m1 = np.full((B, M), -1, dtype=np.int64)
    if N == 0 or M == 0:  # :63-71
        return assignment, m0, m1
    pos2, neg2 = pos_th ** 2, neg_th ** 2
    with np.errstate(invalid="ignore"):
        for b in range(B):
            def sq(p, q):  # torch.sum((p[:, None] - q[None]) ** 2, -1): fp32 differences, squares, one fp32 add
                d = (p[:, None, :] - q[None, :, :]).astype(F32)
                return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(F32)
            dist0 = sq(kp0_1[b], kp1[b])                      # :100
            dist1 = sq(kp0[b], kp1_0[b])                      # :101
            dist = np.maximum(dist0, dist1)                   # :102 (NaN propagates like torch.max)
            vis = np.asarray(visible0[b], bool)[:, None] & np.asarray(visible1[b], bool)[None, :]
            dist = np.where(vis, dist, np.float32(np.inf))    # :103-104
            min0 = dist.argmin(-1)                            # :106 first index on ties
            min1 = dist.argmin(-2)                            # :107
            ismin0 = np.zeros_like(vis)
            ismin1 = np.zeros_like(vis)
            ismin0[np.arange(N), min0] = True
            ismin1[min1, np.arange(M)] = True
            positive = ismin0 & ismin1 & (dist < pos2)        # :113
            # torch.min propagates NaN (np.min too); NaN > x is False
            neg0 = (dist0.min(-1) > neg2) & np.asarray(valid0[b], bool)   # :115
            neg1 = (dist1.min(-2) > neg2) & np.asarray(valid1[b], bool)   # :116
            a0 = np.where(positive.any(-1), min0, -2)         # :123
            a1 = np.where(positive.any(-2), min1, -2)
            m0[b] = np.where(neg0, -1, a0)                    # :125
            m1[b] = np.where(neg1, -1, a1)
            assignment[b] = positive
    return assignment, m0, m1
