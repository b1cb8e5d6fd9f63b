"""Golden fixtures for the ground-truth assignment block (SURVEY.md section 8 f, row 4), from the REAL reference:
gt_matches_from_pose_depth of core/geometry/gt_generation.py (lines 96-126 are the N x M part: the two
reprojection distance matrices, their masked maximum, argmin along both axes, positives and negatives).

The function is run whole -- real Camera / Pose wrappers, real `project` / `sample_depth` -- on synthetic scenes
(a tilted plane seen from two poses); the fixture stores the inputs of the block as the function computed them
(`proj_0to1`, `proj_1to0`, `visible*`, and `valid*` from its own `sample_depth`) and its outputs.
`kornia` (imported by core/geometry/depth.py for a function not on this path) is stubbed with an empty module.

    python tests/golden/make_golden_gt.py        # writes tests/golden/gt_assign.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EINX_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_geometry():
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))
    for name, path in (("core", f"{REF}/core"), ("core.geometry", f"{REF}/core/geometry")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for leaf in ("utils", "wrappers", "homography", "epipolar", "depth"):
        load(f"core.geometry.{leaf}", f"{REF}/core/geometry/{leaf}.py")
    return load("core.geometry.gt_generation", f"{REF}/core/geometry/gt_generation.py")


def scene(rng, B, N, M, H, W, shift):
    """A slanted plane seen by two cameras a small motion apart; keypoints of view 1 are partly reprojections."""
    wr = sys.modules["core.geometry.wrappers"]
    K = torch.tensor([[0.8 * W, 0.0, W / 2], [0.0, 0.8 * W, H / 2], [0.0, 0.0, 1.0]]).expand(B, 3, 3).contiguous()
    cam0 = wr.Camera.from_calibration_matrix(K)
    cam1 = wr.Camera.from_calibration_matrix(K)
    ang = 0.03 * rng.standard_normal(B)
    R = torch.eye(3).repeat(B, 1, 1)
    for b in range(B):
        c, s = np.cos(ang[b]), np.sin(ang[b])
        R[b] = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    t = torch.from_numpy((shift * rng.standard_normal((B, 3))).astype(np.float32))
    T01 = wr.Pose.from_Rt(R, t)
    T10 = T01.inv()
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    depth = []
    for b in range(B):
        d = 4.0 + 0.004 * xs + 0.002 * ys + 0.05 * rng.standard_normal((H, W)).astype(np.float32)
        d[rng.random((H, W)) < 0.05] = 0.0  # holes: invalid depth
        depth.append(d)
    depth0 = torch.from_numpy(np.stack(depth))
    depth1 = torch.from_numpy(np.stack(depth)[:, ::-1, :].copy() * 1.02)
    kp0 = torch.from_numpy(np.stack([rng.uniform(1, H - 2, (B, N)), rng.uniform(1, W - 2, (B, N))], -1).astype(np.float32))  # yx
    kp1 = torch.from_numpy(np.stack([rng.uniform(1, H - 2, (B, M)), rng.uniform(1, W - 2, (B, M))], -1).astype(np.float32))
    return cam0, cam1, depth0, depth1, T01, T10, kp0, kp1


def main():
    torch.set_num_threads(1)
    gt = load_geometry()
    depth_mod = sys.modules["core.geometry.depth"]
    rng = np.random.default_rng(20241101)
    g = {}
    # (B, N, M, H, W, pose shift, pos_th, neg_th, cc_th); the second pass of every case plants true correspondences
    cases = [(2, 300, 280, 120, 160, 0.02, 3, 5, None), (1, 1024, 1000, 260, 346, 0.05, 3, 5, 4.0), (3, 64, 50, 60, 80, 0.01, 2, 4, None)]
    for ci, (B, N, M, H, W, shift, pos_th, neg_th, cc_th) in enumerate(cases):
        cam0, cam1, depth0, depth1, T01, T10, kp0, kp1 = scene(rng, B, N, M, H, W, shift)
        # plant correspondences: the first half of view 1's keypoints are view 0's reprojections plus sub-pixel noise
        first = gt.gt_matches_from_pose_depth(kp0, kp1, cam0, cam1, depth0, depth1, T01, T10, pos_th=pos_th, neg_th=neg_th,
                                              ordering="yx", cc_th=cc_th)
        k = min(N, M) // 2
        proj = first["proj_0to1"][:, :k].clone()           # (x, y) order inside the function
        proj = torch.where(torch.isfinite(proj), proj, torch.full_like(proj, 5.0))
        noise = torch.from_numpy(rng.uniform(-1.5, 1.5, (B, k, 2)).astype(np.float32))
        kp1 = kp1.clone()
        kp1[:, :k] = (proj + noise)[..., [1, 0]].clamp(1.0, min(H, W) - 2.0)  # back to yx
        out = gt.gt_matches_from_pose_depth(kp0, kp1, cam0, cam1, depth0, depth1, T01, T10, pos_th=pos_th, neg_th=neg_th,
                                            ordering="yx", cc_th=cc_th)
        kp0_xy, kp1_xy = kp0[..., [1, 0]], kp1[..., [1, 0]]
        _, valid0 = depth_mod.sample_depth(kp0_xy, depth0)
        _, valid1 = depth_mod.sample_depth(kp1_xy, depth1)
        g[f"c{ci}_kp0"], g[f"c{ci}_kp1"] = kp0_xy.numpy(), kp1_xy.numpy()
        g[f"c{ci}_kp0_1"], g[f"c{ci}_kp1_0"] = out["proj_0to1"].numpy(), out["proj_1to0"].numpy()
        g[f"c{ci}_visible0"], g[f"c{ci}_visible1"] = out["visible0"].numpy(), out["visible1"].numpy()
        g[f"c{ci}_valid0"], g[f"c{ci}_valid1"] = valid0.numpy(), valid1.numpy()
        g[f"c{ci}_th"] = np.array([pos_th, neg_th], dtype=np.float32)
        g[f"c{ci}_assignment"] = np.packbits(out["assignment"].numpy(), axis=-1)
        g[f"c{ci}_m0"], g[f"c{ci}_m1"] = out["matches0"].numpy(), out["matches1"].numpy()
        print(f"case {ci}: positives {int(out['assignment'].sum())}, unmatched {(out['matches0'] == -1).sum().item()}, "
              f"ignored {(out['matches0'] == -2).sum().item()}, visible0 {out['visible0'].sum().item()}")
    g["ncases"] = np.array(len(cases))
    np.savez_compressed(f"{OUT}/gt_assign.npz", **g)
    print("gt_assign", os.path.getsize(f"{OUT}/gt_assign.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
