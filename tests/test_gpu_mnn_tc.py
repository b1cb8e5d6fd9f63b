"""Tensor-core (tcgen05) MNN paths vs the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import einx_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def einx():
    import einx as m

    m.context_for(DEV)
    return m


@pytest.fixture(scope="module")
def synth(einx):
    import importlib

    return importlib.import_module("ei-nexus_official_b200.synth")


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


def check_against(got0, got1, d0, d1, min_stable=0.999):
    """Index parity modulo near-ties (rows where the fp32 and fp64 oracles disagree are
    summation-order dependent); everything else must be bit-equal."""
    r32 = O.mnn_match(d0, d1)
    r64 = O.mnn_match(d0, d1, sim_dtype=np.float64)
    for got, key in ((got0, "matches0"), (got1, "matches1")):
        stable = r32[key] == r64[key]
        assert stable.mean() >= min_stable
        assert np.array_equal(got[stable], r32[key][stable]), key


SHAPES = [(1024, 1024, 256, 1.0), (2048, 2048, 128, 1.41), (1000, 777, 128, 1.41), (130, 4100, 64, 1.0),
          (384, 512, 32, 1.0), (5, 3, 8, 1.0)]


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
@pytest.mark.parametrize("N,M,D,scale", SHAPES)
def test_split_paths_index_parity_with_fp32(einx, synth, N, M, D, scale, precision):
    """The two fp32-accurate tensor-core paths (3xTF32 and the 3-term fp16 split) against the oracle."""
    rng = np.random.default_rng(N * 3 + M)
    pairs = [synth.descriptor_pair(rng, N, M, D, scale, dups=3 if b == 0 else 0) for b in range(3)]
    d0 = cuda(np.stack([p[0] for p in pairs]))
    d1 = cuda(np.stack([p[1] for p in pairs]))
    out = einx.mnn(d0, d1, precision=precision)
    ref = einx.mnn(d0, d1, precision="fp32")
    m0, m1 = out["matches0"].cpu().numpy(), out["matches1"].cpu().numpy()
    for b in range(3):
        check_against(m0[b], m1[b], pairs[b][0], pairs[b][1])
        keep = m0[b] > -1
        assert np.array_equal(m1[b][m0[b][keep]], np.nonzero(keep)[0])
    # and against our own FFMA path: identical except on near-ties
    agree = (out["matches0"] == ref["matches0"]).float().mean().item()
    assert agree > 0.999, agree


@pytest.mark.parametrize("N,M,D,scale", SHAPES)
def test_bf16_kernel_is_exact_on_bf16_inputs(einx, synth, N, M, D, scale):
    """The bf16 path must equal an exact matcher run on the bf16-rounded descriptors."""
    rng = np.random.default_rng(N + 5 * M)
    pairs = [synth.descriptor_pair(rng, N, M, D, scale, dups=3 if b == 0 else 0) for b in range(2)]
    d0 = np.stack([p[0] for p in pairs])
    d1 = np.stack([p[1] for p in pairs])
    out = einx.mnn(cuda(d0), cuda(d1), precision="bf16")
    m0, m1 = out["matches0"].cpu().numpy(), out["matches1"].cpu().numpy()
    for b in range(2):
        check_against(m0[b], m1[b], bf16_round(d0[b]), bf16_round(d1[b]), min_stable=0.99)


@pytest.mark.parametrize("N,D,scale", [(1024, 256, 1.0), (2048, 128, 1.41)])
def test_bf16_match_set_agreement(einx, synth, N, D, scale):
    """north_star: the bf16 path reports its agreement with the fp32 matcher; >= 99.5 %."""
    rng = np.random.default_rng(N)
    pairs = [synth.descriptor_pair(rng, N, N, D, scale) for _ in range(4)]
    d0 = cuda(np.stack([p[0] for p in pairs]))
    d1 = cuda(np.stack([p[1] for p in pairs]))
    a = einx.mnn(d0, d1, precision="bf16")["matches0"]
    b = einx.mnn(d0, d1, precision="fp32")["matches0"]
    agree = (a == b).float().mean().item()
    print(f"bf16 vs fp32 matches0 agreement N={N} D={D}: {agree:.5f}")
    assert agree >= 0.995


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3", "bf16"])
def test_tc_ragged_counts(einx, synth, precision):
    rng = np.random.default_rng(9)
    N, M, D = 300, 520, 64
    pairs = [synth.descriptor_pair(rng, N, M, D, 1.0) for _ in range(4)]
    n0 = np.array([300, 129, 1, 0], dtype=np.int32)
    n1 = np.array([520, 257, 77, 100], dtype=np.int32)
    k0 = rng.random((4, N, 3)).astype(np.float32)
    k1 = rng.random((4, M, 3)).astype(np.float32)
    d0 = np.stack([p[0] for p in pairs])
    d1 = np.stack([p[1] for p in pairs])
    out = einx.mnn(cuda(d0), cuda(d1), cuda(n0), cuda(n1), cuda(k0), cuda(k1), None, None, True, precision)
    for b in range(4):
        a, c = d0[b][: n0[b]], d1[b][: n1[b]]
        m0 = out["matches0"][b].cpu().numpy()
        m1 = out["matches1"][b].cpu().numpy()
        assert (m0[n0[b]:] == -1).all() and (m1[n1[b]:] == -1).all()
        if n0[b] == 0:
            assert (m1 == -1).all() and int(out["num_matches"][b]) == 0
            continue
        if precision == "bf16":
            a, c = bf16_round(a), bf16_round(c)
        check_against(m0[: n0[b]], m1[: n1[b]], a, c, min_stable=0.99)
        nm = int(out["num_matches"][b])
        keep = m0 > -1
        assert nm == keep.sum()
        assert np.array_equal(out["matched_kpts0"][b, :nm].cpu().numpy(), k0[b][keep])
        assert np.array_equal(out["matched_kpts1"][b, :nm].cpu().numpy(), k1[b][m0[keep]])


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3", "fp32"])
def test_similarity_values_are_fp32_accurate(einx, precision):
    """The 3xTF32 split must reproduce fp32 similarity VALUES (not only the argmax): probe them through
    the distance threshold.  Rows are a_i*e_i against b_i*e_i, so sim(i, i) = a_i*b_i exactly; a
    single-pass tf32 product (or a wrong hi/lo split) is off by ~2e-4 and misclassifies dozens of rows."""
    rng = np.random.default_rng(77)
    B, N, D = 32, 64, 64
    a = (1.0 - 0.01 * rng.random((B, N))).astype(np.float32)
    b = (1.0 - 0.01 * rng.random((B, N))).astype(np.float32)
    sign = np.where(rng.random((B, N)) < 0.5, -1.0, 1.0).astype(np.float32)  # negative operands too
    d0 = np.zeros((B, N, D), np.float32)
    d1 = np.zeros((B, N, D), np.float32)
    idx = np.arange(N)
    d0[:, idx, idx] = a * sign
    d1[:, idx, idx] = b * sign
    p = a.astype(np.float64) * b.astype(np.float64)
    T = float(np.median(p))
    thr = float(np.sqrt(2.0 * (1.0 - T)))  # dist = 2 (1 - sim) <= thr^2  <=>  sim >= T
    out = einx.mnn(cuda(d0), cuda(d1), distance_thresh=thr, precision=precision)
    got = (out["matches0"] > -1).cpu().numpy()
    assert np.array_equal(out["matches0"].cpu().numpy()[got], np.broadcast_to(idx, (B, N))[got])
    must, must_not = p >= T + 1e-6, p <= T - 1e-6
    assert must.sum() > 500 and must_not.sum() > 500
    assert got[must].all(), f"{(~got[must]).sum()} rows above the threshold were dropped"
    assert not got[must_not].any(), f"{got[must_not].sum()} rows below the threshold were kept"


@pytest.mark.parametrize("precision", ["fp16x3", "tf32x3"])
def test_split_paths_at_c4_size(einx, synth, precision):
    """BASELINE configs[3]: 8192 x 8192 keypoints, 128-d.  Index parity with the fp32 oracle (near-ties excluded) and the
    size-independent properties: mutual consistency and equal match counts on both sides."""
    rng = np.random.default_rng(8192)
    d0, d1 = synth.descriptor_pair(rng, 8192, 8192, 128, 1.41, dups=5)
    out = einx.mnn(cuda(d0[None]), cuda(d1[None]), precision=precision)
    m0, m1 = out["matches0"][0].cpu().numpy(), out["matches1"][0].cpu().numpy()
    check_against(m0, m1, d0, d1)
    keep = m0 > -1
    assert keep.sum() == (m1 > -1).sum() > 3000
    assert np.array_equal(m1[m0[keep]], np.nonzero(keep)[0])


def test_c5_keypoint_extreme(einx, synth):
    """BASELINE configs[4], top of the keypoint sweep: one pair with 16384 x 16384 keypoints (fp32-accurate default mode)."""
    rng = np.random.default_rng(16384)
    d0, d1 = synth.descriptor_pair(rng, 16384, 16384, 128, 1.41, dups=3)
    out = einx.mnn(cuda(d0[None]), cuda(d1[None]), precision="fp16x3")
    m0, m1 = out["matches0"][0].cpu().numpy(), out["matches1"][0].cpu().numpy()
    check_against(m0, m1, d0, d1)
    keep = m0 > -1
    assert keep.sum() == (m1 > -1).sum() > 6000
    assert np.array_equal(m1[m0[keep]], np.nonzero(keep)[0])


@pytest.mark.parametrize("precision", ["fp16x3", "tf32x3", "bf16"])
def test_c5_batch_extreme(einx, synth, precision):
    """BASELINE configs[4], top of the batch sweep: 1024 pairs with 512 keypoints each; every 97th pair is checked
    against the oracle, all of them for equal match counts on both sides."""
    rng = np.random.default_rng(1024)
    B, K, D = 1024, 512, 64
    pairs = [synth.descriptor_pair(rng, K, K, D, 1.0) for _ in range(12)]
    if precision == "bf16":  # the bf16 kernel is exact on bf16-representable inputs
        pairs = [(bf16_round(a), bf16_round(b)) for a, b in pairs]
    a = np.stack([pairs[i % 12][0] for i in range(B)])
    b = np.stack([pairs[(i * 5 + i // 12) % 12][1] for i in range(B)])
    out = einx.mnn(cuda(a), cuda(b), precision=precision)
    g0, g1 = out["matches0"].cpu().numpy(), out["matches1"].cpu().numpy()
    for i in range(0, B, 97):
        check_against(g0[i], g1[i], a[i], b[i], min_stable=0.99)
    assert ((g0 > -1).sum(1) == (g1 > -1).sum(1)).all()
