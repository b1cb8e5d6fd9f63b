"""Event representation builder -- host side of einx_voxelize.

Keeps the call surface of the reference's ``datasets/representations.py``:
``events_to_voxel_grid(events, input_size, normalize=True)`` (:66-124) takes a dict of four
equal-length numpy arrays and returns a CPU fp32 ``(bins, H, W)`` tensor, with the same side
effects on ``events``.  ``voxelize_batch`` is the batched, device-resident form the pipeline uses.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def time_normalization(events: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Reference ``time_normalization`` (representations.py:8-22), same in-place dict update."""
    events["t"] = events["t"] - events["t"][0]
    events["t"] = events["t"] / (events["t"][-1] + 1e-8)
    return events


def pack_events(batch: Sequence[Dict[str, np.ndarray]], pin: bool = False):
    """Concatenate event windows into the SoA layout of the C ABI.

    Returns CPU tensors (x, y, t, p, offsets): x/y/p fp32 (the cast of representations.py:73-75),
    t fp64 (left untouched for the on-device fp64 offset subtraction), offsets int64 (B+1).
    """
    counts = [len(ev["t"]) for ev in batch]
    for ev, n in zip(batch, counts):
        if n == 0:
            raise IndexError("index 0 is out of bounds: empty event window (reference raises here too)")
        if not (len(ev["x"]) == len(ev["y"]) == len(ev["p"]) == n):
            raise ValueError("events x, y, t, p must have equal length")
    off = np.zeros(len(batch) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    total = int(off[-1])

    def alloc(dtype):
        t = torch.empty(total, dtype=dtype)
        return t.pin_memory() if pin else t

    x, y, p, t = alloc(torch.float32), alloc(torch.float32), alloc(torch.float32), alloc(torch.float64)
    xn, yn, pn, tn = x.numpy(), y.numpy(), p.numpy(), t.numpy()
    for ev, a, b in zip(batch, off[:-1], off[1:]):
        xn[a:b] = ev["x"]  # numpy casts to float32 exactly like .astype("float32")
        yn[a:b] = ev["y"]
        pn[a:b] = ev["p"]
        tn[a:b] = ev["t"]
    return x, y, t, p, torch.from_numpy(off)


def voxelize_device(x, y, t, p, offsets, input_size: Tuple[int, int, int], normalize: bool = True,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """einx_voxelize on device-resident SoA events; returns (B, bins, H, W) fp32 on the same device."""
    bins, H, W = (int(v) for v in input_size)
    dev = x.device
    ctx = _lib.context_for(dev)
    for name, ten, dt in (("x", x, torch.float32), ("y", y, torch.float32), ("p", p, torch.float32),
                          ("t", t, torch.float64), ("offsets", offsets, torch.int64)):
        if ten.dtype != dt or not ten.is_contiguous() or ten.device != dev:
            raise ValueError(f"{name}: expected contiguous {dt} on {dev}")
    B = offsets.numel() - 1
    if out is None:  # registered PyTorch op over einx_voxelize
        return _lib.ops().voxelize(x, y, t, p, offsets, bins, H, W, bool(normalize))
    rc = ctx.lib.einx_voxelize(ctx.handle, _lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), _lib.ptr(offsets),
                               B, bins, H, W, int(bool(normalize)), _lib.ptr(out), ctx.stream)
    ctx.check(rc, "einx_voxelize")
    return out


def voxelize_batch(batch: Sequence[Dict[str, np.ndarray]], input_size, normalize: bool = True,
                   device="cuda") -> torch.Tensor:
    """Ragged batch of event dicts -> (B, bins, H, W) voxel grids on ``device``."""
    x, y, t, p, off = pack_events(batch)
    dev = torch.device(device)
    return voxelize_device(x.to(dev), y.to(dev), t.to(dev), p.to(dev), off.to(dev), input_size, normalize)


@torch.no_grad()
def events_to_voxel_grid(events: Dict, input_size: Tuple, normalize: bool = True, device="cuda") -> torch.Tensor:
    """Drop-in for ``datasets/representations.py:66-124``.

    Same inputs, same CPU fp32 result, same observable side effects: afterwards ``events`` holds
    fp32 torch tensors, ``t`` normalised to [0, 1) and ``p`` with every value < 1 set to -1.
    """
    grid = voxelize_batch([events], input_size, normalize, device)[0].cpu()
    # side effects of :72-76 and :88-89 (cheap host work; the kernel recomputes them on device)
    events = time_normalization(events)
    events["x"] = torch.from_numpy(events["x"].astype("float32"))
    events["y"] = torch.from_numpy(events["y"].astype("float32"))
    events["p"] = torch.from_numpy(events["p"].astype("float32"))
    events["t"] = torch.from_numpy(events["t"].astype("float32"))
    events["p"][events["p"] < 1] = -1
    return grid


# --------------------------------------------------------------------------------------------- #
# adjacent row (SURVEY.md section 8 f): event accumulation image
# --------------------------------------------------------------------------------------------- #
@torch.no_grad()
def events_image_device(x: torch.Tensor, y: torch.Tensor, offsets: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """einx_events_image on device-resident coordinates (fp32 or fp64) -> (B, H, W) uint8."""
    dev = x.device
    if not x.is_cuda:
        raise _lib.EinxError("events_image: expected CUDA tensors (there is no CPU fallback)")
    if x.dtype != y.dtype or x.dtype not in (torch.float32, torch.float64):
        raise ValueError("events_image: x and y must both be float32 or both float64")
    if offsets.dtype != torch.int64:
        raise ValueError("events_image: offsets must be int64")
    ctx = _lib.context_for(dev)
    B = offsets.numel() - 1
    out = torch.empty((B, int(height), int(width)), dtype=torch.uint8, device=dev)
    rc = ctx.lib.einx_events_image(ctx.handle, _lib.ptr(x.contiguous()), _lib.ptr(y.contiguous()),
                                   int(x.dtype == torch.float64), _lib.ptr(offsets.contiguous()), B, int(height),
                                   int(width), _lib.ptr(out), ctx.stream)
    ctx.check(rc, "einx_events_image")
    return out


@torch.no_grad()
def events_image_signed_device(x: torch.Tensor, y: torch.Tensor, p: torch.Tensor, offsets: torch.Tensor, height: int,
                               width: int) -> torch.Tensor:
    """einx_events_image_signed (every event adds 2 * p - 1) on device-resident arrays -> (B, H, W) uint8."""
    dev = x.device
    if not x.is_cuda:
        raise _lib.EinxError("events_image: expected CUDA tensors (there is no CPU fallback)")
    if not (x.dtype == y.dtype == p.dtype) or x.dtype not in (torch.float32, torch.float64):
        raise ValueError("events_image: x, y and p must all be float32 or all float64")
    if offsets.dtype != torch.int64:
        raise ValueError("events_image: offsets must be int64")
    ctx = _lib.context_for(dev)
    B = offsets.numel() - 1
    out = torch.empty((B, int(height), int(width)), dtype=torch.uint8, device=dev)
    rc = ctx.lib.einx_events_image_signed(ctx.handle, _lib.ptr(x.contiguous()), _lib.ptr(y.contiguous()),
                                          _lib.ptr(p.contiguous()), int(x.dtype == torch.float64),
                                          _lib.ptr(offsets.contiguous()), B, int(height), int(width), _lib.ptr(out), ctx.stream)
    ctx.check(rc, "einx_events_image_signed")
    return out


def draw_events_accumulation_image(events, image_shape, device="cuda") -> np.ndarray:
    """Drop-in for ``datasets/visualize.py:23-49``: (H, W) uint8 numpy image.

    ``image_shape`` is (W, H) like the reference's ``RESOLUTION`` tuples.  The coordinates cross to the
    device as fp64, so ``int()`` truncation matches the reference's per-event Python loop exactly.  A dict counts
    events per pixel (:37-40); an (N, 4) array of (x, y, t, p) rows adds ``2 * p - 1`` per event (:41-44) -- the sums
    stay exact for integer-valued ``2 * p - 1`` (polarities 0/1 or -1/1), anything else is refused."""
    dev = torch.device(device)
    if isinstance(events, dict):
        x = torch.from_numpy(np.ascontiguousarray(events["x"], dtype=np.float64)).to(dev)
        y = torch.from_numpy(np.ascontiguousarray(events["y"], dtype=np.float64)).to(dev)
        off = torch.tensor([0, x.numel()], dtype=torch.int64, device=dev)
        return events_image_device(x, y, off, image_shape[1], image_shape[0])[0].cpu().numpy()
    if isinstance(events, np.ndarray):
        if events.ndim != 2 or events.shape[1] < 4:
            raise ValueError("events array must have shape [N, 4]")
        w = 2.0 * events[:, 3].astype(np.float64) - 1.0
        if not np.all(w == np.round(w)):
            raise ValueError("draw_events_accumulation_image: 2 * p - 1 must be integer-valued (polarities 0/1 or -1/1)")
        cols = [torch.from_numpy(np.ascontiguousarray(events[:, k], dtype=np.float64)).to(dev) for k in (0, 1, 3)]
        off = torch.tensor([0, events.shape[0]], dtype=torch.int64, device=dev)
        return events_image_signed_device(cols[0], cols[1], cols[2], off, image_shape[1], image_shape[0])[0].cpu().numpy()
    raise ValueError("events must be a dictionary or numpy array.")


# --------------------------------------------------------------------------------------------- #
# adjacent row (SURVEY.md section 8 f): the reference's other scatter representations
# --------------------------------------------------------------------------------------------- #
def _binned(entry: str, x, y, t, p, offsets, input_size) -> torch.Tensor:
    bins, H, W = (int(v) for v in input_size)
    dev = x.device
    if not x.is_cuda:
        raise _lib.EinxError(f"{entry}: expected CUDA tensors (there is no CPU fallback)")
    for name, ten, dt in (("x", x, torch.float32), ("y", y, torch.float32), ("p", p, torch.float32),
                          ("t", t, torch.float64), ("offsets", offsets, torch.int64)):
        if ten.dtype != dt or not ten.is_contiguous() or ten.device != dev:
            raise ValueError(f"{name}: expected contiguous {dt} on {dev}")
    ctx = _lib.context_for(dev)
    B = offsets.numel() - 1
    out = torch.empty((B, bins, H, W), dtype=torch.float32, device=dev)
    rc = getattr(ctx.lib, entry)(ctx.handle, _lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), _lib.ptr(offsets),
                                 B, bins, H, W, _lib.ptr(out), ctx.stream)
    ctx.check(rc, entry)
    return out


def event_stack_device(x, y, t, p, offsets, input_size) -> torch.Tensor:
    """einx_event_stack on device-resident SoA events -> (B, bins, H, W) fp32."""
    return _binned("einx_event_stack", x, y, t, p, offsets, input_size)


def time_surface_device(x, y, t, p, offsets, input_size) -> torch.Tensor:
    """einx_time_surface on device-resident SoA events -> (B, bins, H, W) fp32."""
    return _binned("einx_time_surface", x, y, t, p, offsets, input_size)


def _single(events: Dict, fn, input_size, device) -> torch.Tensor:
    dev = torch.device(device)
    x, y, t, p, off = (a.to(dev) for a in pack_events([events]))
    grid = fn(x, y, t, p, off, input_size)[0].cpu()
    time_normalization(events)  # the reference mutates the caller's dict (:183, :31)
    return grid


@torch.no_grad()
def events_to_event_stack(events: Dict, input_size: Tuple, device="cuda") -> torch.Tensor:
    """Drop-in for ``datasets/representations.py:177-214``: CPU fp32 (bins, H, W); ``events['t']`` is normalised."""
    return _single(events, event_stack_device, input_size, device)


@torch.no_grad()
def events_to_time_surface(events: Dict, input_size: Tuple, device="cuda") -> torch.Tensor:
    """Drop-in for ``datasets/representations.py:25-63``: CPU fp32 (bins, H, W); ``events['t']`` is normalised."""
    return _single(events, time_surface_device, input_size, device)


@torch.no_grad()
def distance_map_device(x, y, t, offsets, input_size) -> torch.Tensor:
    """einx_distance_map on device-resident SoA events -> (B, bins, H, W) fp32 (polarity is not used)."""
    bins, H, W = (int(v) for v in input_size)
    dev = x.device
    if not x.is_cuda:
        raise _lib.EinxError("einx_distance_map: expected CUDA tensors (there is no CPU fallback)")
    for name, ten, dt in (("x", x, torch.float32), ("y", y, torch.float32), ("t", t, torch.float64),
                          ("offsets", offsets, torch.int64)):
        if ten.dtype != dt or not ten.is_contiguous() or ten.device != dev:
            raise ValueError(f"{name}: expected contiguous {dt} on {dev}")
    ctx = _lib.context_for(dev)
    B = offsets.numel() - 1
    out = torch.empty((B, bins, H, W), dtype=torch.float32, device=dev)
    rc = ctx.lib.einx_distance_map(ctx.handle, _lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(offsets), B, bins, H, W,
                                   _lib.ptr(out), ctx.stream)
    ctx.check(rc, "einx_distance_map")
    return out


def events_to_distance_map(events: Dict, input_size: Tuple, device="cuda") -> torch.Tensor:
    """Drop-in for ``datasets/representations.py:215-248``: CPU fp32 (bins, H, W); ``events['t']`` is normalised."""
    return _single(events, lambda x, y, t, p, off, size: distance_map_device(x, y, t, off, size), input_size, device)
