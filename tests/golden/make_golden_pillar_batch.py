"""Generate tests/golden/pillar_batch_ref.npz by running the reference's own N4 functions as they are:
``points_to_voxel`` per frame + ``merge_second_batch`` (data/preprocess.py:16-42), ``sparse_sum_for_anchors_mask``
/ ``fused_get_anchors_area`` (libs/ops/box_np_ops.py:772-806, cut out with ``ast``: their module imports the whole
detector) with the cumsum chain of data/preprocess.py:272-277, and ``points_to_bev`` (libs/ops/point_cloud/
bev_ops.py, imported by path: it only needs numba + numpy).
Build-container only:  python tests/golden/make_golden_pillar_batch.py"""
import ast
import importlib.util
import os
import sys
from collections import defaultdict

import numpy as np

PP = "/root/reference/PAPC/models/detect/pointpillars/"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def cut(path, names, ns):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(body) == len(names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)


def by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def frames(rng, sizes, F=4):
    out = []
    for n in sizes:
        p = rng.uniform(-1.0, 1.0, (n, F)).astype(np.float32)
        p[:, 0] = p[:, 0] * 6 + 5
        p[:, 1] *= 5
        p[:, 2] = p[:, 2] * 2.5 - 1
        p[:, 3] = rng.random(n).astype(np.float32)
        out.append(p)
    return out


if __name__ == "__main__":
    pc_ops = by_path("ref_pc_ops", PP + "libs/ops/point_cloud/point_cloud_ops.py")
    bev_ops = by_path("ref_bev_ops", PP + "libs/ops/point_cloud/bev_ops.py")
    ns = {"np": np, "defaultdict": defaultdict}
    cut(PP + "data/preprocess.py", ["merge_second_batch"], ns)
    cut(PP + "libs/ops/box_np_ops.py", ["sparse_sum_for_anchors_mask", "fused_get_anchors_area"], ns)
    rng = np.random.default_rng(11)
    out = {}
    vs = np.array([0.4, 0.4, 4.0], np.float32)
    rg = np.array([0.0, -4.0, -3.0, 10.0, 4.0, 1.0], np.float32)
    out["voxel_size"], out["range"] = vs, rg
    # ---- batched voxelisation + merge: ragged frames, an empty one, one that hits the max_voxels break
    for case, sizes, max_points, max_voxels in (("a", (900, 0, 1500, 37), 5, 400), ("b", (3000, 2500), 8, 120)):
        fr = frames(rng, sizes)
        ex = []
        for p in fr:
            v, c, n = pc_ops.points_to_voxel(p, vs, rg, max_points, True, max_voxels)
            ex.append({"voxels": v, "num_points": n, "coordinates": c, "num_voxels": np.array([v.shape[0]], np.int64)})
        merged = ns["merge_second_batch"](ex)
        out[f"{case}_sizes"] = np.array(sizes, np.int64)
        out[f"{case}_cfg"] = np.array([max_points, max_voxels], np.int64)
        for i, p in enumerate(fr):
            out[f"{case}_points{i}"] = p
        assert "num_voxels" not in merged   # popped by the reference (:20)
        for k in ("voxels", "num_points", "coordinates"):
            out[f"{case}_{k}"] = merged[k]
        out[f"{case}_frame_voxels"] = np.array([e["voxels"].shape[0] for e in ex], np.int32)
        if case == "a":
            # ---- anchors mask chain (data/preprocess.py:270-277) on the merged coordinates of frame 0
            coors = ex[0]["coordinates"]
            grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
            dense = ns["sparse_sum_for_anchors_mask"](coors, tuple(grid[::-1][1:]))
            out["am_coors"], out["am_grid"], out["am_dense"] = coors, grid, dense.copy()
            cum = dense.cumsum(0).cumsum(1)
            x1 = rng.uniform(-1.0, 9.0, 300).astype(np.float32)
            y1 = rng.uniform(-4.5, 3.0, 300).astype(np.float32)
            bv = np.stack([x1, y1, x1 + rng.uniform(0.1, 3.0, 300).astype(np.float32),
                           y1 + rng.uniform(0.1, 3.0, 300).astype(np.float32)], axis=1).astype(np.float32)
            bv[:5, 2] = -0.5 + bv[:5, 2] * 0   # boxes entirely left of the map: negative index, NumPy wraps
            area = ns["fused_get_anchors_area"](cum, bv, vs, rg, grid)
            out["am_cum"], out["am_anchors_bv"], out["am_area"] = cum, bv, area
    # ---- points_to_bev: slices, with / without reflectivity, the max_voxels break
    bvs = np.array([0.2, 0.2, 1.0], np.float32)
    pts = frames(rng, (4000,))[0]
    pts[:, :3] = np.round(pts[:, :3] * 8) / 8   # equal heights inside a cell: the '>' keeps the earlier point
    out["bev_points"], out["bev_voxel_size"] = pts, bvs
    out["bev_plain"] = bev_ops.points_to_bev(pts, bvs, rg, False)
    out["bev_refl"] = bev_ops.points_to_bev(pts, bvs, rg, True)
    out["bev_refl_break"] = bev_ops.points_to_bev(pts, bvs, rg, True, max_voxels=700)
    np.savez_compressed(os.path.join(HERE, "pillar_batch_ref.npz"), **out)
    for k, v in out.items():
        print(f"{k:22s} {str(np.asarray(v).dtype):8s} {np.asarray(v).shape}")
