"""The numpy port (oracle/einx_oracle.py) against the reference's OWN functions (oracle/_ref/, oracle/ref_arm.py)
on the inputs bench.py feeds both: the CPU arm bench.py times and the checker the GPU tests use agree."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import einx_oracle as O  # noqa: E402
from oracle import ref_arm  # noqa: E402

synth = importlib.import_module("ei-nexus_official_b200.synth")

pytestmark = pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref/ not made (python oracle/make_ref.py)")


@pytest.mark.parametrize("cfg_name,sample", [("c2_ec_superpoint", 3), ("c3_mvsec_silk_b256", 1)])
def test_port_equals_reference_on_bench_inputs(cfg_name, sample):
    cfg = synth.CONFIGS[cfg_name]
    ev, sides = synth.pair_inputs(cfg_name, sample, None)
    kind = "full" if cfg["kind"] == "gather" else "low"
    args = (cfg["bins"], cfg["H"], cfg["W"])
    g_ref, p0_ref, p1_ref, m_ref = ref_arm.pair_pipeline(ev, *args, sides[0][0].copy(), sides[0][1], sides[1][0].copy(),
                                                         sides[1][1], kind, cfg["top_k"], cfg["scale"])
    g, p0, p1, m = O.pair_pipeline(ev, *args, sides[0][0].copy(), sides[0][1], sides[1][0].copy(), sides[1][1], kind,
                                   cfg["top_k"], cfg["scale"])
    # voxel grid: 1e-5 relative per cell (north_star tolerance; the port accumulates with bincount)
    assert np.all(np.abs(g - g_ref) <= 1e-5 * np.maximum(np.abs(g_ref), 1.0))
    # keypoints: bit-exact; matches: identical indices (both fp32 similarity on the same descriptors, up to the
    # matmul's summation order -- a flip would need a tie within an ulp, adjudicated here rather than tolerated)
    assert np.array_equal(p0, p0_ref) and np.array_equal(p1, p1_ref)
    assert np.array_equal(m["matches0"], m_ref["matches0"])
    assert np.array_equal(m["matches1"], m_ref["matches1"])
    np.testing.assert_allclose(m["matching_scores0"], m_ref["matching_scores0"], rtol=0, atol=2e-6)


def test_ref_arm_is_the_reference_source():
    """oracle/_ref/ holds verbatim copies (made by oracle/make_ref.py); when the checkout is present they match it."""
    from oracle import make_ref
    ref = os.environ.get("EINX_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout here")
    for f in make_ref.FILES:
        with open(os.path.join(ref, f), "rb") as a, open(os.path.join(make_ref.DST, f), "rb") as b:
            assert a.read() == b.read(), f
