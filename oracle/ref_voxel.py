"""Load the reference's OWN numba voxeliser by file path -- build container only.

/root/reference does not exist on the GPU box, so nothing on the -m gpu / smoke / bench
paths imports this module; it is used by tests/golden/make_golden.py (fixture generation)
and by the not-gpu tests when /root/reference is present.  Importing the reference
*package* would pull ``libs.*`` (paddle); the file itself only needs time/numba/numpy
(point_cloud_ops.py:1-4).
"""
import importlib.util
import os

REF_FILE = ("/root/reference/PAPC/models/detect/pointpillars/libs/ops/point_cloud/"
            "point_cloud_ops.py")


def available():
    if not os.path.exists(REF_FILE):
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


_MOD = None


def module():
    global _MOD
    if _MOD is None:
        spec = importlib.util.spec_from_file_location("papc_ref_point_cloud_ops", REF_FILE)
        _MOD = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_MOD)
    return _MOD


def points_to_voxel(*args, **kwargs):
    return module().points_to_voxel(*args, **kwargs)
