"""Load the reference's OWN numba.cuda NMS module (PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/
nms_gpu.py) from its staged copy under oracle/_ref/ -- TEST INFRASTRUCTURE / bench baseline only.

The module's pybind11 build of nms.so (:8-19) and its ``libs.*`` imports are cut out with ``ast`` (as
tests/golden/make_golden_nms.py does); nothing else is modified.  With a GPU the kernels compile through
numba.cuda as the reference runs them (bench.py's N3 baseline); with NUMBA_ENABLE_CUDASIM=1 they run in the
simulator (how the golden vectors were made)."""
import ast
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(path=None):
    path = path or os.path.join(HERE, "_ref", "nms_gpu.py")
    tree = ast.parse(open(path).read())
    body = []
    for n in tree.body:
        if isinstance(n, ast.Try):
            continue                                     # the nms.so build / import
        if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("libs."):
            continue
        if isinstance(n, ast.FunctionDef) and n.name == "nms_gpu_cc":
            continue                                     # wrapper of the pybind11 module
        body.append(n)
    ns = {}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns
