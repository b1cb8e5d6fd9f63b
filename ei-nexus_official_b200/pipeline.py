"""The whole hot path for a batch of event-image pairs, device-resident end to end:

    voxelise(events) -> [detect -> sample] for both sides -> MNN

Nothing synchronises with the host between the stages: keypoints stay in padded (B, K, 3) buffers
with per-image counts, exactly the layout the next kernel consumes.  This is the unit
``bench.py`` times ("pairs/sec") and ``__graft_entry__.smoke()`` runs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, describe, detection as _detect, match, voxel


@dataclass
class PathConfig:
    """Parameters pinned by the reference configs."""
    bins: int = 5
    height: int = 260          # sensor H (voxel grid)
    width: int = 346           # sensor W
    nms_radius: int = 4        # configs/model/SiLK_MNN.yaml:15
    remove_borders: int = 4    # :18
    detection_threshold: float = 1.0  # :16  -> min(topk_thr, 1.0) selects the top-k threshold
    top_k: int = 1024          # :17
    descriptor_mode: str = "bilinear"  # "bilinear" = SuperPoint type (cell 8), "gather" = SiLK type (cell 1)
    descriptor_scale: float = 1.0      # 1.0 (SP, D=256) / 1.41 (SiLK, D=128)
    precision: str = "fp32"    # MNN arithmetic: fp32 | tf32x3 | fp16x3 (both fp32-accurate tensor-core splits) | bf16
    normalize_voxels: bool = True
    concurrent: bool = True    # run voxelisation and the two sides' detect -> sample chains on three streams
    pair_detect: bool = True   # both sides' maps in one detect launch (einx_detect_pair) when their shapes agree


class ExtractMatchPipeline:
    """Batched drop-in for the post-backbone part of ``EIM.forward`` (core/modules/EIM.py:89-93)."""

    def __init__(self, cfg: PathConfig):
        if cfg.precision == "fp16x3" and not abs(cfg.descriptor_scale) < 60.0:
            raise ValueError("precision='fp16x3' splits 2^10 * descriptor into fp16 pairs and needs |descriptor| < 63: "
                             "use 'tf32x3' for descriptor_scale >= 60")
        self.cfg = cfg
        self._side_streams = {}
        self._capture_streams = {}

    @torch.no_grad()
    def voxelize(self, x, y, t, p, offsets) -> torch.Tensor:
        c = self.cfg
        return voxel.voxelize_device(x, y, t, p, offsets, (c.bins, c.height, c.width), c.normalize_voxels)

    @torch.no_grad()
    def extract(self, score: torch.Tensor, raw: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """(B,1,Hp,Wp) score + (B,C,Hd,Wd) raw descriptors -> padded keypoints, counts, descriptors."""
        c = self.cfg
        _, kpts, counts = _detect.detect(score, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k,
                                         mask=mask, want_map=False)
        mode = describe.BILINEAR if c.descriptor_mode == "bilinear" else describe.GATHER
        desc = describe.sample(raw, kpts, counts, mode, score.shape[-2:], c.descriptor_scale, True)
        return kpts, counts, desc

    def _streams(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        st = self._side_streams.get(key)
        if st is None:
            st = self._side_streams[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
        return st

    @torch.no_grad()
    def __call__(self, events, score0, raw0, score1, raw1, mask0=None, mask1=None) -> Dict[str, torch.Tensor]:
        """events = (x, y, t, p, offsets) on the device; score/raw maps of both sides on the device.

        The three branches of the step share no data until the matcher: voxelisation (L2-reduction bound)
        and the second side's detect -> sample chain are forked onto two side streams, each with its own
        einx context (workspace); the side chain joins before the MNN kernel, the voxel stream after it -- so the
        latency-bound NMS rounds of one side overlap the other side's sampling and the event scatter instead of
        queueing behind them."""
        c = self.cfg
        mode = describe.BILINEAR if c.descriptor_mode == "bilinear" else describe.GATHER
        # fp16x3: the samplers write the matcher's fp16 hi / lo operands next to the fp32 descriptors (one fused
        # kernel), so the similarity kernel's pipeline is TMA -> tcgen05.mma with nothing to convert
        want_split = c.precision == "fp16x3" and raw0.shape[1] % 8 == 0 and raw1.shape[1] % 8 == 0
        sp0 = sp1 = None

        def sample(raw, k, n, size):
            r = describe.sample(raw, k, n, mode, size, c.descriptor_scale, True, split=want_split)
            return r if want_split else (r, None)

        if not c.concurrent and c.pair_detect and score0.shape == score1.shape:
            grid = self.voxelize(*events)
            (k0, c0), (k1, c1) = _detect.detect_pair(score0, score1, c.detection_threshold, c.nms_radius, c.remove_borders,
                                                    c.top_k, mask0, mask1)
            d0, sp0 = sample(raw0, k0, c0, score0.shape[-2:])
            d1, sp1 = sample(raw1, k1, c1, score1.shape[-2:])
        elif not c.concurrent:
            grid = self.voxelize(*events)
            _, k0, c0 = _detect.detect(score0, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k, mask=mask0)
            d0, sp0 = sample(raw0, k0, c0, score0.shape[-2:])
            _, k1, c1 = _detect.detect(score1, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k, mask=mask1)
            d1, sp1 = sample(raw1, k1, c1, score1.shape[-2:])
        else:
            dev = score0.device
            main = torch.cuda.current_stream(dev)
            s_vox, s_side = self._streams(dev)
            s_vox.wait_stream(main)
            with torch.cuda.stream(s_vox):
                grid = self.voxelize(*events)
            if c.pair_detect and score0.shape == score1.shape:
                # both sides' maps in ONE detect launch (an image is owned by one CTA, so 2B images fill the machine
                # where two launches of B would run back to back); the two sampling kernels then run side by side
                (k0, c0), (k1, c1) = _detect.detect_pair(score0, score1, c.detection_threshold, c.nms_radius,
                                                        c.remove_borders, c.top_k, mask0, mask1)
                s_side.wait_stream(main)
                with torch.cuda.stream(s_side):
                    d1, sp1 = sample(raw1, k1, c1, score1.shape[-2:])
                d0, sp0 = sample(raw0, k0, c0, score0.shape[-2:])
            else:
                s_side.wait_stream(main)
                with torch.cuda.stream(s_side):
                    _, k1, c1 = _detect.detect(score1, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k, mask=mask1)
                    d1, sp1 = sample(raw1, k1, c1, score1.shape[-2:])
                _, k0, c0 = _detect.detect(score0, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k, mask=mask0)
                d0, sp0 = sample(raw0, k0, c0, score0.shape[-2:])
            main.wait_stream(s_side)
            if not torch.cuda.is_current_stream_capturing():
                # caching-allocator bookkeeping: tensors cross streams in both directions (a captured
                # step owns its memory pool for the lifetime of the graph instead)
                for t in (k1, c1, d1, grid, sp1):
                    if t is not None:
                        t.record_stream(main)
                for t in (score1, raw1, mask1, k1, c1):
                    if t is not None:
                        t.record_stream(s_side)
                for t in events:
                    t.record_stream(s_vox)
        out = match.mnn(d0, d1, c0, c1, k0, k1, None, None, True, self.cfg.precision, sp0, sp1)
        if self.cfg.concurrent:
            # the matcher does not read the voxel grid: its stream joins only here, so a late event scatter never
            # holds the MNN kernel back
            torch.cuda.current_stream(score0.device).wait_stream(s_vox)
        out.update(voxel_grid=grid, keypoints0=k0, keypoints1=k1, counts0=c0, counts1=c1,
                   descriptors0=d0, descriptors1=d1)
        return out

    def capture(self, events, score0, raw0, score1, raw1, mask0=None, mask1=None) -> "CapturedStep":
        """Record one step over these (fixed) device buffers into a CUDA graph.

        The dozen launches of a step and their fork/join edges then cost one graph launch; the caller
        refreshes the input buffers in place (or keeps one captured step per resident batch) and calls
        ``replay()``.  Warm-up runs first on the capture stream so that every per-stream context has its
        workspace before capture starts (cudaMalloc is not capturable)."""
        dev = score0.device
        cur = torch.cuda.current_stream(dev)
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        cap = self._capture_streams.get(key)
        if cap is None:
            cap = self._capture_streams[key] = torch.cuda.Stream(dev)
        args = (events, score0, raw0, score1, raw1, mask0, mask1)
        cap.wait_stream(cur)
        with torch.cuda.stream(cap):
            for _ in range(2):
                self(*args)
        cap.synchronize()
        graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (the NCCL watchdog under torchrun, samplers) keep using CUDA
        with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):
            out = self(*args)
        cur.wait_stream(cap)
        return CapturedStep(graph, out, args)


class CapturedStep:
    """A step of the path frozen into a CUDA graph (see ExtractMatchPipeline.capture)."""

    def __init__(self, graph, outputs, inputs):
        self.graph, self.outputs, self.inputs = graph, outputs, inputs  # inputs kept alive: the graph reads them

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.outputs


class HostBatch:
    """Pinned host inputs of one batch of pairs, split into ``len(chunks)`` contiguous sub-batches so
    that the host->device copy of one sub-batch overlaps the kernels of the previous one.

    Every sub-batch is ONE pinned byte buffer holding its nine arrays (x, y, t, p, offsets, score0, raw0,
    score1, raw1) back to back at 256-byte aligned offsets, so it crosses PCIe as a single large copy."""

    ALIGN = 256

    def __init__(self, events: Sequence[dict], score0=None, raw0=None, score1=None, raw1=None, chunks: int = 4,
                 compact_events: bool = True):
        """``score0 .. raw1`` = None: an events-only batch -- the maps are produced on the device (by the conv backbones
        in the real model, core/modules/EIM.py:89-93) and handed to ``HostStreamer.run(..., resident_maps=...)``."""
        import numpy as np

        B = len(events)
        chunks = max(1, min(int(chunks), B))
        bounds = [B * i // chunks for i in range(chunks + 1)]
        self.batch = B
        # integer-pixel events (EC: datasets/rectify_ec.py:66-83; any raw sensor stream) cross PCIe as uint16 x, y and
        # int8 p -- 13 instead of 20 bytes per event with the fp64 timestamp -- when that is lossless for the whole batch
        self.compact = bool(compact_events) and all(self._lossless(ev) for ev in events)
        self.chunks: List[tuple] = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            if self.compact:
                sub = events[a:b]
                off = np.zeros(len(sub) + 1, dtype=np.int64)
                np.cumsum([len(ev["t"]) for ev in sub], out=off[1:])
                ev = (torch.from_numpy(np.concatenate([np.asarray(e["x"]) for e in sub]).astype(np.uint16)),
                      torch.from_numpy(np.concatenate([np.asarray(e["y"]) for e in sub]).astype(np.uint16)),
                      torch.from_numpy(np.concatenate([np.asarray(e["t"], dtype=np.float64) for e in sub])),
                      torch.from_numpy(np.concatenate([np.asarray(e["p"]) for e in sub]).astype(np.int8)),
                      torch.from_numpy(off))
            else:
                ev = voxel.pack_events(events[a:b])
            maps = ([] if score0 is None else
                    [torch.from_numpy(m[a:b]) if not torch.is_tensor(m) else m[a:b] for m in (score0, raw0, score1, raw1)])
            arrays = [t.contiguous() for t in (*ev, *maps)]
            layout, off_b = [], 0
            for t in arrays:
                layout.append((off_b, t.dtype, tuple(t.shape)))
                off_b += (t.numel() * t.element_size() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            buf = torch.empty(off_b, dtype=torch.uint8).pin_memory()
            for t, (o, dt, shape) in zip(arrays, layout):
                n = t.numel() * t.element_size()
                buf[o:o + n].view(dt).view(shape).copy_(t)
            self.chunks.append((a, b, buf, layout, sum(t.numel() * t.element_size() for t in arrays)))

    @staticmethod
    def _lossless(ev) -> bool:
        import numpy as np

        if len(ev["t"]) == 0:
            return False
        for key, lo, hi in (("x", 0, 65535), ("y", 0, 65535), ("p", -128, 127)):
            v = np.asarray(ev[key])
            if not (np.all(v == np.floor(v)) and v.min() >= lo and v.max() <= hi):
                return False
        return True

    @property
    def nbytes(self) -> int:
        """Payload bytes (without alignment padding)."""
        return sum(c[4] for c in self.chunks)

    @staticmethod
    def views(buf: torch.Tensor, layout):
        out = []
        for o, dt, shape in layout:
            n = 1
            for d in shape:
                n *= d
            out.append(buf[o:o + n * torch.empty((), dtype=dt).element_size()].view(dt).view(shape))
        return out


RESULT_KEYS = ("matches0", "matching_scores0", "num_matches", "matched_kpts0", "matched_kpts1")


class HostStreamer:
    """End-to-end driver of the path for inputs that live in pinned HOST memory.

    A copy stream uploads sub-batch i+1 into the second of two device staging sets while the compute
    streams run sub-batch i; results are written back to the caller's pinned host tensors on the
    compute stream.  Everything is asynchronous: the caller synchronises (or records an event) when it
    needs the results."""

    def __init__(self, pipe: ExtractMatchPipeline, device, graphs: bool = True):
        """``graphs``: freeze the device work of a sub-batch (event unpack, the pipeline's launches on its three
        streams, the result copies to the caller's pinned tensors) into a CUDA graph the first time a (staging set,
        sub-batch layout, output buffers) combination is seen and replay it afterwards -- one launch per sub-batch
        instead of ~25 eager calls, which keeps the PCIe-bound path from becoming host bound on a busy host.  A
        sub-batch whose layout differs (ragged event counts) is captured under its own key."""
        self.pipe = pipe
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.use_graphs = bool(graphs)
        self._cap_stream = torch.cuda.Stream(self.dev) if self.use_graphs else None
        self._graphs = {}
        self._stage = {}
        self._unpacked = {}
        self._free = [None, None]  # event: the kernels that read staging set k have finished

    def _staging(self, k, nbytes):
        st = self._stage.get(k)
        if st is None or st.numel() < nbytes:
            st = self._stage[k] = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            self._graphs = {key: g for key, g in self._graphs.items() if key[0] != k}  # captured addresses are gone
        return st

    def _body(self, hb, k, a, b, dbuf, layout, out_host, resident_maps):
        """Device work of one sub-batch on the current stream: unpack, pipeline, result copies."""
        if resident_maps is None:
            x, y, t, p, off, s0, r0, s1, r1 = HostBatch.views(dbuf, layout)
        else:
            x, y, t, p, off = HostBatch.views(dbuf, layout)
            s0, r0, s1, r1 = (m[a:b] for m in resident_maps)
        if hb.compact:  # expand the 13-byte wire format to the fp32 SoA the voxeliser reads
            n = x.numel()
            f = self._unpacked[k]
            ctx = _lib.context_for(self.dev)
            ctx.check(ctx.lib.einx_unpack_events(ctx.handle, _lib.ptr(x), _lib.ptr(y), _lib.ptr(p), n, _lib.ptr(f[0]),
                                                 _lib.ptr(f[1]), _lib.ptr(f[2]), ctx.stream), "einx_unpack_events")
            x, y, p = f[0, :n], f[1, :n], f[2, :n]
        out = self.pipe((x, y, t, p, off), s0, r0, s1, r1)
        for key in RESULT_KEYS:
            if key in out_host:
                out_host[key][a:b].copy_(out[key], non_blocking=True)

    @torch.no_grad()
    def run(self, hb: HostBatch, out_host: Dict[str, torch.Tensor], resident_maps=None) -> None:
        """``resident_maps`` = (score0, raw0, score1, raw1) device tensors of the whole batch, for an events-only
        ``hb``: only the events cross PCIe (the boundary of EIM.forward: maps come from on-device backbones)."""
        main = torch.cuda.current_stream(self.dev)
        # (the copy stream never waits for the compute stream as a whole -- only, through `_free`, for the kernels
        # that last read the staging set it is about to overwrite -- so uploads of the next call start while the
        # last sub-batch of this one still computes)
        for i, (a, b, hbuf, layout, _) in enumerate(hb.chunks):
            k = i & 1
            dbuf = self._staging(k, hbuf.numel())[:hbuf.numel()]
            if hb.compact:
                n = layout[0][2][0]  # events of the sub-batch
                f = self._unpacked.get(k)
                if f is None or f.shape[1] < n:
                    self._unpacked[k] = torch.empty((3, n), dtype=torch.float32, device=self.dev)
                    self._graphs = {key: g for key, g in self._graphs.items() if key[0] != k}
            with torch.cuda.stream(self.copy_stream):
                if self._free[k] is not None:
                    self.copy_stream.wait_event(self._free[k])
                dbuf.copy_(hbuf, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            main.wait_event(ready)
            args = (hb, k, a, b, dbuf, layout, out_host, resident_maps)
            if not self.use_graphs:
                self._body(*args)
            else:
                key = (k, a, b, hbuf.numel(), hb.compact, tuple(layout),
                       tuple((name, t.data_ptr()) for name, t in sorted(out_host.items())),
                       None if resident_maps is None else tuple(m.data_ptr() for m in resident_maps))
                graph = self._graphs.get(key)
                if graph is None:
                    # warm-up on the capture stream (every per-stream context gets its workspace: nothing may allocate
                    # during capture), then capture; the warm-up already produced this sub-batch's results
                    cap = self._cap_stream
                    cap.wait_stream(main)
                    with torch.cuda.stream(cap):
                        self._body(*args)
                    cap.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):
                        self._body(*args)
                    main.wait_stream(cap)
                    self._graphs[key] = graph
                else:
                    graph.replay()
            self._free[k] = torch.cuda.Event()
            self._free[k].record(main)
