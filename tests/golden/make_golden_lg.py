"""Golden fixtures for the LightGlue log-assignment matrix (SURVEY.md section 8 f, row 3), from the REAL reference.

Runs only in the build container (reference mounted at /root/reference); writes tests/golden/lg.npz:
inputs (sim, z0, z1) and the output of core/modules/matchers/lightglue.py:365 sigmoid_log_double_softmax,
plus filter_matches (:402) applied to it.

    python tests/golden/make_golden_lg.py
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, load_reference  # noqa: E402
from make_golden_next import load  # noqa: E402


def main():
    torch.set_num_threads(1)
    load_reference()
    for name in ("matplotlib", "matplotlib.pyplot", "omegaconf"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["omegaconf"].OmegaConf = object
    lg = load("core.modules.matchers.lightglue", f"{REF}/core/modules/matchers/lightglue.py")
    rng = np.random.default_rng(20241019)
    g = {}
    # (B, M, N, similarity scale, threshold): ragged tile edges (128 x 256 tiles in the kernel), a single row,
    # a single column, large-magnitude similarities (exp underflow across a row) and strongly negative logits
    cases = [(2, 70, 90, 4.0, 0.1), (1, 130, 300, 1.0, 0.0), (3, 33, 31, 12.0, 0.2), (1, 1, 5, 2.0, 0.0), (2, 257, 1, 3.0, 0.0),
             (1, 129, 260, 30.0, 0.1)]
    for ci, (B, M, N, scale, th) in enumerate(cases):
        sim = torch.from_numpy((scale * rng.standard_normal((B, M, N))).astype(np.float32))
        z0 = torch.from_numpy((3.0 * rng.standard_normal((B, M, 1))).astype(np.float32))
        z1 = torch.from_numpy((3.0 * rng.standard_normal((B, N, 1))).astype(np.float32))
        if ci == 2:
            z0[0, :4, 0] = torch.tensor([-40.0, 40.0, -100.0, 0.0])
        scores = lg.sigmoid_log_double_softmax(sim, z0, z1)
        m0, m1, s0, s1 = lg.filter_matches(scores, th)
        g[f"c{ci}_sim"], g[f"c{ci}_z0"], g[f"c{ci}_z1"] = sim.numpy(), z0.numpy(), z1.numpy()
        g[f"c{ci}_scores"], g[f"c{ci}_th"] = scores.numpy(), np.array(th)
        g[f"c{ci}_m0"], g[f"c{ci}_m1"], g[f"c{ci}_s0"], g[f"c{ci}_s1"] = m0.numpy(), m1.numpy(), s0.numpy(), s1.numpy()
    g["ncases"] = np.array(len(cases))
    np.savez_compressed(f"{OUT}/lg.npz", **g)
    print("lg", os.path.getsize(f"{OUT}/lg.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
