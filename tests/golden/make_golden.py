"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Runs only in the build container, where the reference checkout is mounted at
/root/reference (it does not exist on the GPU box; nothing at test time reads
it).  The reference's five leaf modules are imported by file path with stub
parent packages, so ``core/modules/__init__.py`` (kornia / hydra / lightning,
not installed) never executes -- SURVEY.md appendix B.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The fixtures are inputs + the reference's own outputs; tests compare both the
oracle (CPU) and the CUDA path (GPU) against them.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EINX_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    stub("core", f"{REF}/core")
    stub("core.modules", f"{REF}/core/modules")
    stub("core.modules.utils", f"{REF}/core/modules/utils")
    stub("core.modules.matchers", f"{REF}/core/modules/matchers")

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    det = load("core.modules.utils.detector_util", f"{REF}/core/modules/utils/detector_util.py")
    desc = load("core.modules.utils.descriptor_util", f"{REF}/core/modules/utils/descriptor_util.py")
    util = load("core.modules.utils.util", f"{REF}/core/modules/utils/util.py")
    mnn = load("core.modules.matchers.MNN", f"{REF}/core/modules/matchers/MNN.py")
    rep = load("einx_ref_representations", f"{REF}/datasets/representations.py")
    return det, desc, util, mnn, rep


def synth_events(rng, n, H, W, style, T0=1.5e9, dt=0.4):
    """SURVEY.md section 8 d: sub-pixel (MVSEC) or integer-pixel (EC) events, epoch-scale t."""
    x = rng.uniform(0, W - 1, n)
    y = rng.uniform(0, H - 1, n)
    if style == "ec":
        x, y = np.floor(x), np.floor(y)
        p = rng.integers(0, 2, n).astype(np.float64)  # 0/1
    else:
        p = rng.integers(0, 2, n).astype(np.float64) * 2 - 1  # -1/+1
    t = np.sort(rng.uniform(T0, T0 + dt, n))
    return {"x": x, "y": y, "t": t, "p": p}


def smooth_map(rng, B, H, W):
    """A smooth positive score map (box-blurred noise through a sigmoid)."""
    z = torch.from_numpy(rng.standard_normal((B, 1, H, W)).astype(np.float32))
    k = torch.ones(1, 1, 5, 5) / 25.0
    z = torch.nn.functional.conv2d(z, k, padding=2)
    return (1 / (1 + torch.exp(-4 * z))).numpy().astype(np.float32)


def main():
    torch.set_num_threads(1)  # deterministic put_ accumulation order
    det, desc, util, mnn, rep = load_reference()
    rng = np.random.default_rng(20241017)

    # ---- voxel grids ---------------------------------------------------- #
    vox = {}
    cases = [("mvsec", 6000, 5, 48, 64), ("ec", 4000, 16, 30, 40), ("mvsec", 9000, 10, 36, 52),
             ("ec", 300, 3, 12, 16)]
    for ci, (style, n, bins, H, W) in enumerate(cases):
        ev = synth_events(rng, n, H, W, style, dt=0.4 if style == "mvsec" else 0.04)
        for k in "xytp":
            vox[f"c{ci}_{k}"] = ev[k]
        vox[f"c{ci}_shape"] = np.array([bins, H, W])
        raw = rep.events_to_voxel_grid({k: v.copy() for k, v in ev.items()}, (bins, H, W), normalize=False)
        nrm = rep.events_to_voxel_grid({k: v.copy() for k, v in ev.items()}, (bins, H, W), normalize=True)
        vox[f"c{ci}_raw"] = raw.numpy()
        vox[f"c{ci}_norm"] = nrm.numpy()
    vox["ncases"] = np.array(len(cases))
    np.savez_compressed(f"{OUT}/voxel.npz", **vox)

    # ---- detection ------------------------------------------------------ #
    dg = {}
    maps = {
        "uniform": rng.random((3, 1, 64, 80)).astype(np.float32),
        "ties": (np.round(rng.random((2, 1, 56, 72)) * 8) / 8).astype(np.float32),
        "smooth": smooth_map(rng, 2, 72, 96),
        "ec_sp": rng.random((1, 1, 184, 240)).astype(np.float32),
    }
    # a masked variant: a third of the map zeroed, like score[~mask] = 0
    masked = smooth_map(rng, 2, 64, 64)
    masked[:, :, :, 40:] = 0
    maps["masked"] = masked
    ks = {"uniform": [None, 20, 60, 10 ** 6], "ties": [None, 30], "smooth": [None, 40, 100],
          "ec_sp": [1024, 256], "masked": [64]}
    names = []
    for name, m in maps.items():
        dg[f"{name}_in"] = m
        for k in ks[name]:
            for thr in ([0.0, 1.0] if name != "ec_sp" else [1.0]):
                src = torch.from_numpy(m.copy())
                nms = det.prob_map_to_points_map(src, prob_thresh=thr, nms_dist=4, border_dist=4,
                                                 use_fast_nms=True, top_k=k)
                pos = det.prob_map_to_positions_with_prob(nms, threshold=0.0, ordering="yx")
                tag = f"{name}_k{k}_t{thr}"
                names.append(tag)
                dg.setdefault(f"{name}_border", src.numpy())  # the in-place border-zeroed input (same for every k)
                for i, p_i in enumerate(pos):
                    dg[f"{tag}_pos{i}"] = p_i.numpy()
    # radius / border variants
    for (r, b) in [(2, 0), (1, 3), (6, 2)]:
        src = torch.from_numpy(maps["uniform"].copy())
        nms = det.prob_map_to_points_map(src, prob_thresh=0.0, nms_dist=r, border_dist=b,
                                         use_fast_nms=True, top_k=50)
        pos = det.prob_map_to_positions_with_prob(nms, threshold=0.0, ordering="yx")
        tag = f"uniform_r{r}_b{b}"
        names.append(tag)
        dg[f"{tag}_border"] = src.numpy()
        for i, p_i in enumerate(pos):
            dg[f"{tag}_pos{i}"] = p_i.numpy()
    dg["tags"] = np.array(names)
    # the reference's own NMS property test (utils_test.py:31-63): seed 0, rand(32,60,80);
    # store the fast_nms result sparsely (greedy == fast is re-checked on a slice in the tests)
    torch.manual_seed(0)
    inp = torch.rand((32, 60, 80))
    fast = det.prob_map_to_points_map(inp.clone(), 0.0, 4, 4, use_fast_nms=True)
    slow = det.prob_map_to_points_map(inp.clone()[:4], 0.0, 4, 4, use_fast_nms=False)
    assert torch.equal(slow, fast[:4])
    nz = torch.nonzero(fast)
    dg["parity_nz"] = nz.numpy().astype(np.int32)
    dg["parity_val"] = fast[tuple(nz.T)].numpy()
    np.savez_compressed(f"{OUT}/detect.npz", **dg)

    # ---- descriptor sampling ------------------------------------------- #
    sg = {}
    score = rng.random((2, 1, 48, 64)).astype(np.float32)
    nms = det.prob_map_to_points_map(torch.from_numpy(score.copy()), 1.0, 4, 4, True, 40)
    pos = det.prob_map_to_positions_with_prob(nms, 0.0, "yx")
    raw_full = rng.standard_normal((2, 16, 48, 64)).astype(np.float32)
    raw_low = rng.standard_normal((2, 32, 6, 8)).astype(np.float32)
    full = desc.sparsify_full_resolution_descriptors(torch.from_numpy(raw_full), pos,
                                                     torch.tensor(1.41), True)
    low = desc.sparsify_low_resolution_descriptors(torch.from_numpy(raw_low), pos, (48, 64),
                                                   torch.tensor(1.0), True)
    sg["raw_full"], sg["raw_low"] = raw_full, raw_low
    for i in range(2):
        sg[f"pos{i}"] = pos[i].numpy()
        sg[f"full{i}"] = full[i].numpy()
        sg[f"low{i}"] = low[i].numpy()
    np.savez_compressed(f"{OUT}/sample.npz", **sg)

    # ---- MNN ------------------------------------------------------------ #
    mg = {}

    def descs(n, m, d, scale):
        a = rng.standard_normal((n, d))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        b = rng.standard_normal((m, d))
        perm = rng.permutation(m)[: min(n, m) // 2]
        b[perm] = a[: perm.size] + 0.2 * rng.standard_normal((perm.size, d))
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        return (scale * a).astype(np.float32), (scale * b).astype(np.float32)

    mcases = [(70, 90, 32, 1.0, None, None), (128, 96, 64, 1.41, None, None),
              (60, 60, 32, 1.0, 0.9, None), (60, 60, 32, 1.0, None, 0.8), (50, 40, 16, 1.0, 0.95, 1.0)]
    for ci, (n, m, d, scale, ratio, dist) in enumerate(mcases):
        a, b = descs(n, m, d, scale)
        if ci == 1:  # exact duplicate rows: first-index tie-breaking
            b[5] = b[17]
            a[9] = a[3]
        k0 = np.concatenate([rng.uniform(0, 100, (n, 2)), rng.random((n, 1))], 1).astype(np.float32)
        k1 = np.concatenate([rng.uniform(0, 100, (m, 2)), rng.random((m, 1))], 1).astype(np.float32)
        matcher = mnn.NearestNeighborMatcher(ratio, dist, True)
        out = matcher({"sparse_descriptors": torch.from_numpy(a)[None], "sparse_positions": torch.from_numpy(k0)[None]},
                      {"sparse_descriptors": torch.from_numpy(b)[None], "sparse_positions": torch.from_numpy(k1)[None]})
        mg[f"c{ci}_d0"], mg[f"c{ci}_d1"], mg[f"c{ci}_k0"], mg[f"c{ci}_k1"] = a, b, k0, k1
        mg[f"c{ci}_cfg"] = np.array([ratio or 0.0, dist or 0.0])
        for key in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
            mg[f"c{ci}_{key}"] = out[key][0].numpy()
        mg[f"c{ci}_matched_kpts0"] = out["matched_kpts0"].numpy()
        mg[f"c{ci}_matched_kpts1"] = out["matched_kpts1"].numpy()
        mg[f"c{ci}_log_assignment"] = out["log_assignment"][0].numpy()
    mg["ncases"] = np.array(len(mcases))
    np.savez_compressed(f"{OUT}/mnn.npz", **mg)
    for f in ("voxel", "detect", "sample", "mnn"):
        print(f, os.path.getsize(f"{OUT}/{f}.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
