"""Generate tests/golden/pillar_glue_ref.npz by running two pure-NumPy pieces of the reference as they are
(cut out with ``ast``; their modules import the whole detector): ``VoxelGenerator.__init__`` / properties
(PAPC/models/detect/pointpillars/core/voxel_generator.py:5-43) and ``merge_second_batch``
(data/preprocess.py:16-42, the batch layout the pillar encoder consumes).
Build-container only:  python tests/golden/make_golden_pillar_glue.py"""
import ast
import os
from collections import defaultdict

import numpy as np

PP = "/root/reference/PAPC/models/detect/pointpillars/"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, names, ns):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(body) == len(names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)


if __name__ == "__main__":
    ns = {"np": np, "defaultdict": defaultdict, "points_to_voxel": None}
    cut(PP + "core/voxel_generator.py", ["VoxelGenerator"], ns)
    cut(PP + "data/preprocess.py", ["merge_second_batch"], ns)
    out = {}
    for tag, vs, rg in (("yaml", (0.16, 0.16, 4.0), (0.0, -39.68, -3.0, 69.12, 39.68, 1.0)),
                        ("default", (0.2, 0.2, 4.0), (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)),
                        ("odd", (0.3, 0.7, 0.25), (-10.0, -7.0, -1.0, 11.5, 7.35, 1.1))):
        g = ns["VoxelGenerator"](vs, rg, 100, 12000)
        out[f"{tag}_args"] = np.array(list(vs) + list(rg), np.float64)
        out[f"{tag}_voxel_size"], out[f"{tag}_range"], out[f"{tag}_grid"] = g.voxel_size, g.point_cloud_range, g.grid_size
        assert g.max_num_points_per_voxel == 100
    rng = np.random.default_rng(5)
    batch = []
    for i, p in enumerate((7, 0, 12)):
        batch.append({"voxels": rng.random((p, 5, 4)).astype(np.float32),
                      "num_points": rng.integers(1, 6, p).astype(np.int32),
                      "coordinates": rng.integers(0, 50, (p, 3)).astype(np.int32),
                      "num_voxels": np.array([p], np.int64)})
        for k in ("voxels", "num_points", "coordinates"):
            out[f"batch{i}_{k}"] = batch[-1][k]
    merged = ns["merge_second_batch"](batch)
    for k, v in merged.items():
        out[f"merged_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "pillar_glue_ref.npz"), **out)
    for k, v in out.items():
        print(f"{k:22s} {str(np.asarray(v).dtype):8s} {np.asarray(v).shape}")
