"""Generate tests/golden/pc_normalize.npz by RUNNING THE REFERENCE'S OWN ``pc_normalize``
(PAPC/models/layers/pointnet2_basic_layers.py:17-23).  The module imports paddle at its top, so the
function's source is cut out with ``ast`` and executed with NumPy alone -- nothing is restated.
Build-container only (needs /root/reference):  python tests/golden/make_golden_pc_normalize.py"""
import ast
import os

import numpy as np

REF = "/root/reference/PAPC/models/layers/pointnet2_basic_layers.py"
OUT = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "pc_normalize")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    rng = np.random.default_rng(0)
    x32 = rng.uniform(-1.0, 1.0, (1024, 3)).astype(np.float32)
    x64 = rng.normal(0.0, 3.0, (257, 3))
    np.savez_compressed(os.path.join(OUT, "pc_normalize.npz"), x32=x32, y32=ns["pc_normalize"](x32),
                        x64=x64, y64=ns["pc_normalize"](x64))
    print("pc_normalize golden written:", ns["pc_normalize"](x32).dtype, ns["pc_normalize"](x64).dtype)
