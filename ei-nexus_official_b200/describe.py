"""Descriptor sampling + L2 normalisation -- host side of einx_sample.

Keeps the reference surface of ``core/modules/utils/descriptor_util.py``:
``sparsify_full_resolution_descriptors`` (:50-71) and ``sparsify_low_resolution_descriptors``
(:74-128).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib

GATHER, BILINEAR, GATHER_NHWC = 0, 1, 2


def pack_rows(rows: Sequence[torch.Tensor], width: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ragged (N_i, width) tensors -> zero-padded (B, cap, width) + int32 counts."""
    B = len(rows)
    cap = max([int(r.shape[0]) for r in rows] + [1])
    out = torch.zeros((B, cap, width), dtype=torch.float32, device=device)
    for i, r in enumerate(rows):
        if r.shape[0]:
            out[i, : r.shape[0]] = r[:, :width]
    counts = torch.tensor([int(r.shape[0]) for r in rows], dtype=torch.int32, device=device)
    return out, counts


@torch.no_grad()
def sample(raw: torch.Tensor, kpts: torch.Tensor, counts: torch.Tensor, mode: int, image_size=(0, 0),
           scale_factor=1.0, normalize: bool = True, split: bool = False):
    """einx_sample on padded keypoints: (B, C, Hd, Wd) map -> (B, kcap, C) descriptors (rows >= count zero).

    ``split=True`` (einx_sample_split) also returns the (2, B, kcap, C) fp16 [hi | lo] operands of the matcher's
    ``fp16x3`` mode, written by the same kernel: pass them to :func:`match.mnn` as ``split0`` / ``split1``."""
    if raw.dtype != torch.float32 or not raw.is_cuda:
        raise _lib.EinxError("sample: raw descriptors must be a float32 CUDA tensor (there is no CPU fallback)")
    C = raw.shape[1]
    if not (mode == GATHER and C % 4 == 0 and C > 1 and not raw.is_contiguous()
            and raw.is_contiguous(memory_format=torch.channels_last)):
        # (a channels-last map -- what cuDNN convolutions produce on Blackwell -- is read in place by the op: its
        # memory is (B, Hd, Wd, C), so a keypoint's descriptor is one contiguous read instead of C strided sectors)
        raw = raw.contiguous()
    op = _lib.ops().sample_split if split else _lib.ops().sample
    return op(raw, kpts.contiguous(), counts, int(mode), int(image_size[0]), int(image_size[1]), float(scale_factor),
              bool(normalize))


def _sparsify(raw, positions, mode, image_size, scale_factor, normalize):
    kpts, counts = pack_rows(positions, 3 if positions[0].shape[-1] >= 3 else 2, raw.device)
    if kpts.shape[-1] == 2:
        kpts = torch.cat((kpts, torch.zeros_like(kpts[..., :1])), dim=-1)
    desc = sample(raw, kpts, counts, mode, image_size, float(scale_factor), normalize)
    return [desc[i, : positions[i].shape[0]].clone() for i in range(len(positions))]


def sparsify_full_resolution_descriptors(raw_descriptors, positions, scale_factor: float = 1.0,
                                         normalize: bool = True):
    """Drop-in for ``descriptor_util.py:50-71`` (positions in 'yx' order, like every shipped config)."""
    return tuple(_sparsify(raw_descriptors, positions, GATHER, (0, 0), scale_factor, normalize))


def sparsify_low_resolution_descriptors(raw_descriptors, positions, image_size, scale_factor: float = 1.0,
                                        normalize: bool = True):
    """Drop-in for ``descriptor_util.py:74-128``; ``image_size`` is the padded (H, W) the positions live in."""
    return _sparsify(raw_descriptors, positions, BILINEAR, image_size, scale_factor, normalize)
