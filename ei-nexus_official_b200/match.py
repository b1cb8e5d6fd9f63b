"""Mutual-nearest-neighbour matcher -- host side of einx_mnn.

``NearestNeighborMatcher`` keeps the constructor and the forward contract of the reference's
``core/modules/matchers/MNN.py:35-140``.  ``similarity`` / ``log_assignment`` are only
materialised when ``return_dense=True`` (no live consumer in the reference);
otherwise the keys are present with value ``None`` so ``core/modules/Matchers.py:180-203`` keeps working.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

FP32, TF32X3, BF16, FP16X3 = 0, 1, 2, 3
_PRECISION = {"fp32": FP32, "tf32x3": TF32X3, "bf16": BF16, "fp16x3": FP16X3}


@torch.no_grad()
def mnn(desc0: torch.Tensor, desc1: torch.Tensor, n0: Optional[torch.Tensor] = None,
        n1: Optional[torch.Tensor] = None, kpts0: Optional[torch.Tensor] = None,
        kpts1: Optional[torch.Tensor] = None, ratio_thresh=None, distance_thresh=None, mutual: bool = True,
        precision="fp32", split0: Optional[torch.Tensor] = None, split1: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """einx_mnn on padded batches: desc0 (B, N, D), desc1 (B, M, D) fp32 CUDA; n0/n1 valid-row counts.

    Returns matches0/1 (int64, -1 = none), matching_scores0/1 and, when keypoints are given, the
    compacted ``matched_kpts0/1`` (B, N, 3) with ``num_matches`` (B,) int32.  ``split0`` / ``split1``: the fp16
    operand pairs ``describe.sample(..., split=True)`` wrote for these descriptors (``fp16x3`` only; without them the
    library derives the operands in a pre-pass -- same results).
    """
    if desc0.dtype != torch.float32 or not desc0.is_cuda or desc1.dtype != torch.float32 or not desc1.is_cuda:
        raise _lib.EinxError("mnn: descriptors must be float32 CUDA tensors (there is no CPU fallback)")
    desc0, desc1 = desc0.contiguous(), desc1.contiguous()
    B, N, D = desc0.shape
    M = desc1.shape[1]
    if desc1.shape[0] != B or desc1.shape[2] != D:
        raise ValueError("mnn: desc0 / desc1 batch or feature size mismatch")
    prec = _PRECISION[precision] if isinstance(precision, str) else int(precision)
    if kpts0 is not None:
        if kpts0.shape[-1] != 3 or kpts1.shape[-1] != 3:
            raise ValueError("mnn: keypoint rows must be (y, x, prob) -- three columns")
        kpts0, kpts1 = kpts0.contiguous(), kpts1.contiguous()
    res = _lib.ops().mnn(desc0, desc1, n0, n1, kpts0, kpts1, float(ratio_thresh or 0.0), float(distance_thresh or 0.0),
                         bool(mutual), prec, split0, split1)
    m0, m1, s0, s1 = res[:4]
    if kpts0 is not None:
        mk0, mk1, nm = res[4:]
    out = {"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1}
    if kpts0 is not None:
        out.update(matched_kpts0=mk0, matched_kpts1=mk1, num_matches=nm)
    return out


def _rows3(kpts: torch.Tensor) -> torch.Tensor:
    """(B, N, 2 | 3 | more) keypoints -> (B, N, 3) fp32: the kernels read rows of three floats (y, x, prob);
    two-column positions get a zero third column, extra columns are dropped."""
    k = kpts[..., :3].float()
    if k.shape[-1] < 3:
        k = torch.cat((k, k.new_zeros(k.shape[:-1] + (3 - k.shape[-1],))), dim=-1)
    return k


@torch.no_grad()
def mnn_dense(desc0: torch.Tensor, desc1: torch.Tensor):
    """Opt-in ``similarity`` (B, N, M) and ``log_assignment`` (B, N+1, M+1) of MNN.py:88, :96-98."""
    desc0, desc1 = desc0.contiguous(), desc1.contiguous()
    B, N, D = desc0.shape
    M = desc1.shape[1]
    dev = desc0.device
    ctx = _lib.context_for(dev)
    sim = torch.empty((B, N, M), dtype=torch.float32, device=dev)
    la = torch.empty((B, N + 1, M + 1), dtype=torch.float32, device=dev)
    rc = ctx.lib.einx_mnn_dense(ctx.handle, _lib.ptr(desc0), _lib.ptr(desc1), B, N, M, D, _lib.ptr(sim),
                                _lib.ptr(la), ctx.stream)
    ctx.check(rc, "einx_mnn_dense")
    return sim, la


class NearestNeighborMatcher(nn.Module):
    """Drop-in for ``core/modules/matchers/MNN.py:35-140``."""

    def __init__(self, ratio_thresh=None, distance_thresh=None, mutual_check=True, precision="fp32",
                 return_dense=False):
        super().__init__()
        self.ratio_thresh = ratio_thresh
        self.distance_thresh = distance_thresh
        self.mutual_check = mutual_check
        self.precision = precision
        self.return_dense = return_dense

    @torch.no_grad()
    def forward(self, feats0: Dict[str, torch.Tensor], feats1: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        desc0, desc1 = feats0["sparse_descriptors"], feats1["sparse_descriptors"]
        kpts0, kpts1 = feats0["sparse_positions"], feats1["sparse_positions"]
        b, n, m = desc0.shape[0], desc0.shape[1], desc1.shape[1]
        dense = self.return_dense
        if kpts0.numel() == 0 or kpts1.numel() == 0:  # MNN.py:63-86
            print("No keypoints found in either image")
            empty0 = [kpts0.new_zeros((kpts0.shape[1], 3))] * b if b > 1 else kpts0.new_zeros((0, 3))
            empty1 = [kpts1.new_zeros((kpts1.shape[1], 3))] * b if b > 1 else kpts1.new_zeros((0, 3))
            return {
                "matches0": desc0.new_full((b, n), -1), "matches1": desc1.new_full((b, m), -1),
                "matching_scores0": desc0.new_zeros((b, n)), "matching_scores1": desc1.new_zeros((b, m)),
                "matched_kpts0": empty0, "matched_kpts1": empty1,
                "similarity": desc0.new_zeros((b, n, m)) if dense else None,
                "log_assignment": desc0.new_zeros((b, n + 1, m + 1)) if dense else None,
            }
        out = mnn(desc0, desc1, None, None, _rows3(kpts0), _rows3(kpts1), self.ratio_thresh,
                  self.distance_thresh, self.mutual_check, self.precision)
        if self.mutual_check:
            assert (out["matches0"] > -1).sum() == (out["matches1"] > -1).sum()  # MNN.py:95
        nm = out.pop("num_matches").tolist()
        mk0, mk1 = out["matched_kpts0"], out["matched_kpts1"]
        if b > 1:
            out["matched_kpts0"] = [mk0[i, : nm[i]] for i in range(b)]
            out["matched_kpts1"] = [mk1[i, : nm[i]] for i in range(b)]
        else:
            out["matched_kpts0"], out["matched_kpts1"] = mk0[0, : nm[0]], mk1[0, : nm[0]]
        out["similarity"], out["log_assignment"] = mnn_dense(desc0, desc1) if dense else (None, None)
        return out


@torch.no_grad()
def filter_matches(scores: torch.Tensor, th: float):
    """Drop-in for ``core/modules/matchers/lightglue.py:402-418``: matches from a log-assignment matrix
    (B, M+1, N+1) -> ``(m0, m1, mscores0, mscores1)``; one pass over the matrix (einx_filter_matches)."""
    if scores.dtype != torch.float32 or not scores.is_cuda:
        raise _lib.EinxError("filter_matches: expected a float32 CUDA tensor (there is no CPU fallback)")
    if scores.dim() != 3:
        raise ValueError("filter_matches: expected (B, M+1, N+1)")
    carried = getattr(scores, "_einx_best_keys", None)  # left by sigmoid_log_double_softmax on this very tensor
    scores = scores.contiguous()
    B, M, N = scores.shape[0], scores.shape[1] - 1, scores.shape[2] - 1
    dev = scores.device
    ctx = _lib.context_for(dev)
    m0 = torch.empty((B, M), dtype=torch.int64, device=dev)
    m1 = torch.empty((B, N), dtype=torch.int64, device=dev)
    s0 = torch.empty((B, M), dtype=torch.float32, device=dev)
    s1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    if carried is not None and carried[1] == scores._version and carried[0].numel() == B * (M + N) and M > 0 and N > 0:
        # the row / column maxima were reduced while the matrix was written and it has not been modified since
        rc = ctx.lib.einx_filter_matches_keys(ctx.handle, _lib.ptr(carried[0]), B, M, N, float(th), _lib.ptr(m0), _lib.ptr(m1),
                                              _lib.ptr(s0), _lib.ptr(s1), ctx.stream)
        ctx.check(rc, "einx_filter_matches_keys")
        return m0, m1, s0, s1
    rc = ctx.lib.einx_filter_matches(ctx.handle, _lib.ptr(scores), B, M, N, float(th), _lib.ptr(m0), _lib.ptr(m1),
                                     _lib.ptr(s0), _lib.ptr(s1), ctx.stream)
    ctx.check(rc, "einx_filter_matches")
    return m0, m1, s0, s1


def sigmoid_log_double_softmax(sim: torch.Tensor, z0: torch.Tensor, z1: torch.Tensor, carry_best: bool = True) -> torch.Tensor:
    """Drop-in for ``core/modules/matchers/lightglue.py:365-377``: the (B, M+1, N+1) log-assignment matrix from
    similarities (B, M, N) and matchability logits z0 (B, M, 1), z1 (B, N, 1) (einx_log_double_softmax: row and
    column log-sum-exp in one pass, the matrix written in a second pass).  Forward only: raises when a
    gradient is required (LightGlue training keeps the reference's function).

    ``carry_best``: the write pass also reduces the row / column maxima of the values it stores and the returned
    tensor carries them (``_einx_best_keys``); ``filter_matches`` called on that same, unmodified tensor -- what
    LightGlue.forward does next (lightglue.py:393 -> :402) -- then skips its pass over the matrix."""
    if torch.is_grad_enabled() and (sim.requires_grad or z0.requires_grad or z1.requires_grad):
        raise _lib.EinxError("sigmoid_log_double_softmax: forward only -- call it under torch.no_grad(), or keep the "
                             "reference's function when training the matcher")
    for name, v in (("sim", sim), ("z0", z0), ("z1", z1)):
        if v.dtype != torch.float32 or not v.is_cuda:
            raise _lib.EinxError(f"sigmoid_log_double_softmax: {name} must be a float32 CUDA tensor (there is no CPU fallback)")
    if sim.dim() != 3:
        raise ValueError("sigmoid_log_double_softmax: expected sim of shape (B, M, N)")
    B, M, N = sim.shape
    if z0.numel() != B * M or z1.numel() != B * N:
        raise ValueError("sigmoid_log_double_softmax: z0 / z1 must be (B, M, 1) / (B, N, 1)")
    dev = sim.device
    scores = torch.empty((B, M + 1, N + 1), dtype=torch.float32, device=dev)
    if M == 0 or N == 0:  # nothing to normalise: only the unmatched row / column (torch accepts this shape)
        scores.zero_()
        if M:
            scores[:, :-1, -1] = torch.nn.functional.logsigmoid(-z0.reshape(B, M))
        if N:
            scores[:, -1, :-1] = torch.nn.functional.logsigmoid(-z1.reshape(B, N))
        return scores
    sim, z0, z1 = sim.contiguous(), z0.reshape(B, M).contiguous(), z1.reshape(B, N).contiguous()
    ctx = _lib.context_for(dev)
    keys = torch.empty((B * (M + N),), dtype=torch.int64, device=dev) if carry_best else None
    rc = ctx.lib.einx_log_double_softmax(ctx.handle, _lib.ptr(sim), _lib.ptr(z0), _lib.ptr(z1), B, M, N, _lib.ptr(scores),
                                         _lib.ptr(keys), ctx.stream)
    ctx.check(rc, "einx_log_double_softmax")
    if carry_best:
        scores._einx_best_keys = (keys, scores._version)
    return scores
