"""ctypes binding of the C ABI in include/einx.h.  There is no fallback: a missing library or a
missing sm_100a device raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libeinx.so")
TORCH_LIB_PATH = os.path.join(_PKG, "libeinx_torch.so")  # TORCH_LIBRARY(einx, ...) over the C ABI (csrc/torch/einx_torch.cpp)

c_ctx = C.c_void_p
_P = C.c_void_p

# name -> (restype, argtypes); mirrors include/einx.h one to one
SIGNATURES = {
    "einx_version": (C.c_int, []),
    "einx_create": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "einx_destroy": (None, [c_ctx]),
    "einx_last_error": (C.c_char_p, [c_ctx]),
    "einx_launch_count": (C.c_int64, [c_ctx]),
    "einx_profile_enable": (C.c_int, [c_ctx, C.c_int]),
    "einx_profile_read": (C.c_int, [c_ctx, C.POINTER(C.c_float)]),
    "einx_voxelize": (C.c_int, [c_ctx, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_detect": (C.c_int, [c_ctx, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                              _P, _P, C.c_int, _P, _P]),
    "einx_detect_pair": (C.c_int, [c_ctx, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                   _P, _P, _P, _P, C.c_int, _P, _P, _P]),
    "einx_sample": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                              C.c_int, C.c_float, C.c_int, _P, _P]),
    "einx_sample_split": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                    C.c_int, C.c_float, C.c_int, _P, _P, _P]),
    "einx_mnn_split": (C.c_int, [c_ctx, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                 C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "einx_mnn": (C.c_int, [c_ctx, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                           C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "einx_mnn_dense": (C.c_int, [c_ctx, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "einx_events_image": (C.c_int, [c_ctx, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_events_image_signed": (C.c_int, [c_ctx, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_distance_map": (C.c_int, [c_ctx, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_pairwise_min_dist": (C.c_int, [c_ctx, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "einx_gt_assign": (C.c_int, [c_ctx, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                 _P, _P, _P, _P, _P]),
    "einx_event_stack": (C.c_int, [c_ctx, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_time_surface": (C.c_int, [c_ctx, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_unpack_events": (C.c_int, [c_ctx, _P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "einx_mask_dilate": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_logits_to_score": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "einx_filter_matches": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P]),
    "einx_log_double_softmax": (C.c_int, [c_ctx, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "einx_filter_matches_keys": (C.c_int, [c_ctx, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P]),
}

_lib = None
_torch_lib = None
_lock = threading.Lock()


class EinxError(RuntimeError):
    pass


def load():
    """dlopen libeinx.so and declare every prototype of include/einx.h."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EinxError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc 12.9). There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def load_torch_ops():
    """Load the PyTorch operator library (``torch.ops.einx.*``): the hot path's entry points as registered ops with
    CUDA and Meta kernels.  Returns its ctypes handle (for ``einx_torch_context``)."""
    global _torch_lib
    load()
    with _lock:
        if _torch_lib is not None:
            return _torch_lib
        if not os.path.exists(TORCH_LIB_PATH):
            raise EinxError(
                f"{TORCH_LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for this path.")
        import torch

        torch.ops.load_library(TORCH_LIB_PATH)
        h = C.CDLL(TORCH_LIB_PATH)
        h.einx_torch_context.restype = c_ctx
        h.einx_torch_context.argtypes = [C.c_int, C.c_void_p]
        _torch_lib = h
        return h


class Context:
    """One einx_ctx per CUDA device (workspace owner).  Not thread-safe, like the C object."""

    def __init__(self, device: int, stream: int = 0):
        self.lib = load()
        self.device = int(device)
        # the operator library owns the contexts (one per device and stream), so torch.ops.einx.* and the ctypes
        # calls of this package share workspaces, launch counters and profiling slots
        h = load_torch_ops().einx_torch_context(self.device, C.c_void_p(stream))
        if not h:
            msg = self.lib.einx_last_error(None)
            raise EinxError(f"einx_create({device}) failed: {msg.decode() if msg else ''}")
        self.handle = C.c_void_p(h)
        self.stream = C.c_void_p(stream)  # the one stream this context serves

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.einx_last_error(self.handle)
            raise EinxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def profile(self, on: bool):
        """Bracket each entry point's dominant kernel with CUDA events on the caller's stream."""
        self.check(self.lib.einx_profile_enable(self.handle, int(on)), "einx_profile_enable")

    def profile_read(self):
        """ms of the last (voxel scatter, detect, sample, MNN similarity) kernel; synchronises."""
        out = (C.c_float * 4)()
        self.check(self.lib.einx_profile_read(self.handle, out), "einx_profile_read")
        return [float(v) for v in out]

    @property
    def launches(self) -> int:
        return int(self.lib.einx_launch_count(self.handle))



_contexts = {}


def _raw_stream(idx: int) -> int:
    import torch

    try:
        return int(torch._C._cuda_getCurrentRawStream(idx))  # one C call; the public path builds a Stream object
    except AttributeError:  # pragma: no cover
        return int(torch.cuda.current_stream(idx).cuda_stream)


def context_for(device) -> Context:
    """Context of a torch device / ordinal and of the CURRENT stream on it (created on first use).

    One einx_ctx owns one stream-ordered workspace, so kernels issued on different streams must not
    share it: every (device, stream) pair gets its own context and concurrent streams never alias.
    ``ctx.stream`` is that stream as the ``einx_stream`` argument of the C ABI."""
    import torch

    if isinstance(device, int):
        idx = device
    else:
        dev = device if isinstance(device, torch.device) else torch.device(device)
        if dev.type != "cuda":
            raise EinxError(f"einx kernels run on CUDA sm_100a only; got device '{dev}'. There is no CPU fallback.")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
    sid = _raw_stream(idx)
    key = (idx, sid)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = _contexts[key] = Context(idx, sid)
    return ctx


def contexts_of(device):
    """Every context created so far on a device (one per stream that has run einx kernels)."""
    import torch

    dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return [c for (d, _), c in _contexts.items() if d == idx]


def launch_count(device) -> int:
    """Kernel launches issued through libeinx on a device, over all of its streams."""
    return sum(c.launches for c in contexts_of(device))


def ops():
    """``torch.ops.einx`` -- the registered PyTorch operators (loads libeinx_torch.so on first use)."""
    import torch

    load_torch_ops()
    return torch.ops.einx


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(device):
    return context_for(device).stream
