"""CPU-side checks: the C-ABI library loads and exports what include/einx.h declares, the product
path refuses to run without a GPU (no fallback), and the host-side helpers behave."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def einx():
    import einx as m

    return m


@pytest.fixture(scope="module")
def built_lib():
    build = importlib.import_module("ei-nexus_official_b200.build")
    return build.build_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "einx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(einx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = header_symbols()
    assert {"einx_create", "einx_destroy", "einx_voxelize", "einx_detect", "einx_sample", "einx_mnn"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/einx.h but not exported"
    lib.einx_version.restype = ctypes.c_int
    assert lib.einx_version() == 100


def test_ctypes_table_matches_header(einx):
    assert sorted(einx._lib.SIGNATURES) == header_symbols()


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "--list-elf", built_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(einx):
    ev = {"x": np.array([1.0, 2.0]), "y": np.array([1.0, 2.0]), "t": np.array([0.0, 1.0]), "p": np.array([1.0, 0.0])}
    with pytest.raises(Exception):
        einx.events_to_voxel_grid(ev, (3, 8, 8))
    with pytest.raises(einx.EinxError):
        einx.detect(torch.rand(1, 1, 16, 16), 1.0, 4, 4, 10)
    with pytest.raises(einx.EinxError):
        einx.mnn(torch.rand(1, 4, 8), torch.rand(1, 4, 8))
    with pytest.raises(einx.EinxError):
        einx.sample(torch.rand(1, 8, 4, 4), torch.zeros(1, 2, 3), torch.zeros(1, dtype=torch.int32), 0)
    with pytest.raises(einx.EinxError):
        einx._lib.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ei-nexus_official_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(^|\n)\s*(from|import)\s+oracle|einx_oracle|oracle/", text), f"{f} uses oracle/"


def test_pack_events_layout(einx):
    rng = np.random.default_rng(0)
    batch = [{"x": rng.random(n) * 10, "y": rng.random(n) * 10, "t": np.sort(rng.random(n)) + 1.5e9,
              "p": rng.integers(0, 2, n).astype(np.float64)} for n in (5, 1, 9)]
    x, y, t, p, off = einx.pack_events(batch)
    assert off.tolist() == [0, 5, 6, 15] and off.dtype == torch.int64
    assert x.dtype == torch.float32 and t.dtype == torch.float64
    assert np.array_equal(x[5:6].numpy(), batch[1]["x"].astype("float32"))
    assert np.array_equal(t[6:].numpy(), batch[2]["t"])
    with pytest.raises(IndexError):
        einx.pack_events([{k: np.zeros(0) for k in "xytp"}])


def test_time_normalization_matches_oracle(einx):
    from oracle import einx_oracle as O

    t = np.sort(np.random.default_rng(1).uniform(1.5e9, 1.5e9 + 0.4, 100))
    ev = einx.time_normalization({"t": t.copy()})
    assert np.array_equal(ev["t"], O.time_normalization(t))


def test_shard_range_partitions(einx):
    for total in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [einx.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_padded_size_and_seed():
    synth = importlib.import_module("ei-nexus_official_b200.synth")
    from oracle import einx_oracle as O

    assert synth.padded_size(260, 346, 8) == (264, 352, (3, 3, 2, 2))
    assert synth.padded_size(180, 240, 8) == (184, 240, (0, 0, 2, 2))
    assert synth.padded_size(180, 240, 8)[2] == O.padder_sizes(180, 240, 8)
    assert synth.seed_for(2, 5) == 1234 + 2000 + 5


def test_max_keypoints_bound():
    det = importlib.import_module("ei-nexus_official_b200.detection")
    from oracle import einx_oracle as O

    rng = np.random.default_rng(4)
    for r in (1, 2, 4):
        v = O.fast_nms(rng.random((2, 40, 56)).astype(np.float32), r)
        assert (v > 0).reshape(2, -1).sum(1).max() <= det.max_keypoints(40, 56, r)


def test_fp16x3_needs_bounded_descriptors(einx):
    with pytest.raises(ValueError):
        einx.ExtractMatchPipeline(einx.PathConfig(precision="fp16x3", descriptor_scale=100.0))
    einx.ExtractMatchPipeline(einx.PathConfig(precision="fp16x3", descriptor_scale=1.41))


def test_patch_reference_swaps_adjacent_rows(einx):
    import types

    import torch

    calls = []

    def ref_fn(name):
        def f(*a, **k):
            calls.append(name)
            return "reference"
        f.__module__ = "fake_ref"
        return f

    fake = types.ModuleType("fake_ref.detector_util")
    for name in ("logits_to_prob", "depth_to_space", "prob_map_to_points_map"):
        setattr(fake, name, ref_fn(name))
    lg = types.ModuleType("fake_ref.lightglue")
    lg.filter_matches = ref_fn("filter_matches")
    lg.sigmoid_log_double_softmax = ref_fn("sigmoid_log_double_softmax")
    vis = types.ModuleType("fake_ref.visualize")
    vis.draw_events_accumulation_image = ref_fn("draw_events_accumulation_image")
    done = einx.patch_reference([fake, lg, vis])
    # inference-side functions are replaced outright
    assert fake.prob_map_to_points_map is einx.prob_map_to_points_map
    assert lg.filter_matches is einx.filter_matches
    # dataset-side functions run in DataLoader workers: only on request
    assert "fake_ref.visualize" not in done and not hasattr(vis.draw_events_accumulation_image, "_einx_guard")
    assert set(done) == {"fake_ref.detector_util", "fake_ref.lightglue"}
    # differentiable outputs: the kernels are forward only, so a tensor that requires grad goes to the reference
    x = torch.zeros(1, 65, 2, 2, requires_grad=True)
    assert fake.logits_to_prob(x) == "reference" and fake.depth_to_space(x, 8) == "reference"
    assert lg.sigmoid_log_double_softmax(x, x, x) == "reference"
    assert calls == ["logits_to_prob", "depth_to_space", "sigmoid_log_double_softmax"]
    with torch.no_grad(), pytest.raises(einx.EinxError):  # inference: ours (which refuses CPU tensors -- no fallback)
        fake.logits_to_prob(x)
    with pytest.raises(einx.EinxError):
        fake.logits_to_prob(x.detach())
    # patching twice does not wrap a wrapper
    assert einx.patch_reference([fake, lg, vis]) == {}
    done = einx.patch_reference([vis], datasets=True)
    assert done == {"fake_ref.visualize": ["draw_events_accumulation_image"]}
    assert vis.draw_events_accumulation_image._einx_original.__module__ == "fake_ref"


def test_dataset_functions_defer_to_reference_in_loader_workers(einx):
    """A patched events_to_voxel_grid inside a (forked) DataLoader worker must not touch CUDA."""
    import types

    import torch

    ds_mod = types.ModuleType("fake_ref.representations")

    def ref_voxel(events, input_size, normalize=True):
        return torch.full(tuple(input_size), 7.0)

    ref_voxel.__module__ = "fake_ref"
    ds_mod.events_to_voxel_grid = ref_voxel
    einx.patch_reference([ds_mod], datasets=True)

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 2

        def __getitem__(self, i):
            ev = {k: np.zeros(4) for k in "xytp"}
            return ds_mod.events_to_voxel_grid(ev, (2, 3, 4))

    out = next(iter(torch.utils.data.DataLoader(DS(), batch_size=2, num_workers=1)))
    assert out.shape == (2, 2, 3, 4) and bool((out == 7.0).all())


def test_points_map_cache_is_versioned(einx):
    import torch

    det = importlib.import_module("ei-nexus_official_b200.detection")
    t = torch.zeros(1, 4, 4)
    wrapped = det._PointsMap.wrap(t, "kpts", "counts")
    assert wrapped._einx_kpts[2] == t._version
    wrapped.mul_(2.0)  # an in-place edit invalidates the carried keypoints
    assert wrapped._einx_kpts[2] != wrapped._version


def test_log_double_softmax_host_checks(einx):
    """No CPU fallback, no silent loss of the autograd graph, shape errors before anything reaches the library."""
    import torch

    sim, z0, z1 = torch.zeros(1, 3, 2), torch.zeros(1, 3, 1), torch.zeros(1, 2, 1)
    with pytest.raises(einx.EinxError, match="CUDA"):
        einx.sigmoid_log_double_softmax(sim, z0, z1)
    with pytest.raises(einx.EinxError, match="forward only"):
        einx.sigmoid_log_double_softmax(sim.clone().requires_grad_(), z0, z1)
    with torch.no_grad():  # under no_grad a leaf that requires grad is fine for the reference too: only the device check fires
        with pytest.raises(einx.EinxError, match="CUDA"):
            einx.sigmoid_log_double_softmax(sim.clone().requires_grad_(), z0, z1)
    with pytest.raises(einx.EinxError):
        einx.filter_matches(torch.zeros(1, 4, 3), 0.1)


def test_torch_ops_registered_with_meta_kernels(einx):
    """The hot path's entry points are registered PyTorch operators (TORCH_LIBRARY(einx) in libeinx_torch.so, built by
    build()): schemas carry the in-place annotation of the score maps, and the Meta kernels give shapes / dtypes
    without a GPU -- what FakeTensor tracing (torch.compile) needs."""
    import torch

    ops = einx._lib.ops()
    for name in ("voxelize", "detect", "detect_pair", "sample", "mnn"):
        assert hasattr(ops, name), name
    assert "Tensor(a!) score" in str(ops.detect.default._schema)
    assert "Tensor(a!) score0, Tensor(b!) score1" in str(ops.detect_pair.default._schema)
    m = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device="meta")
    B, K, D = 3, 128, 64
    g = ops.voxelize(m(1000), m(1000), m(1000, dtype=torch.float64), m(1000), m(B + 1, dtype=torch.int64), 5, 60, 80, True)
    assert g.shape == (B, 5, 60, 80) and g.dtype == torch.float32
    kp, cn, mp = ops.detect(m(B, 1, 64, 96), None, 4, 4, 1.0, K, K, False)
    assert kp.shape == (B, K, 3) and cn.shape == (B,) and cn.dtype == torch.int32 and mp.numel() == 0
    k0, c0, k1, c1 = ops.detect_pair(m(B, 1, 64, 96), m(B, 1, 64, 96), None, None, 4, 4, 1.0, K, K)
    assert k0.shape == k1.shape == (B, K, 3) and c0.dtype == torch.int32
    d = ops.sample(m(B, D, 8, 12), kp, cn, 1, 64, 96, 1.0, True)
    assert d.shape == (B, K, D)
    out = ops.mnn(d, d, cn, cn, kp, kp, 0.0, 0.0, True, 3)
    assert [tuple(t.shape) for t in out] == [(B, K), (B, K), (B, K), (B, K), (B, K, 3), (B, K, 3), (B,)]
    assert out[0].dtype == torch.int64 and out[6].dtype == torch.int32
    assert len(ops.mnn(d, d, None, None, None, None, 0.0, 0.0, True, 0)) == 4
    # no CPU kernels are registered: a CPU tensor is refused by the dispatcher (no fallback)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.sample(torch.zeros(1, 4, 2, 2), torch.zeros(1, 1, 3), torch.zeros(1, dtype=torch.int32), 0, 2, 2, 1.0, True)


def test_torch_ops_error_path_unwinds(einx):
    """A failing C-ABI call inside a registered op must become a c10::Error (-> RuntimeError in Python), not a crash:
    the operator library is linked so that exceptions unwind through it (build.py pins the system g++)."""
    assert einx._lib.load_torch_ops().einx_torch_error_path_selftest() == 1
