"""Per-CTA %globaltimer timeline of every grouped-MLP layer launch of the bench step (triage build:
PAPC_NVCC_EXTRA=-DPAPC_TT_TRIAGE; run with PAPC_TT_GCLK=1).  No host synchronisation between the
launches, so programmatic dependent launch and the real kernel-to-kernel boundaries are observed.
usage: PAPC_TT_GCLK=1 python tools/prof_gclk.py [out.txt]   then   python tools/gclk_summary.py out.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from papc_b200 import _lib, sa_stack, synth  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "gclk.txt")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
B = bench.B_PER_GPU
xyz = torch.from_numpy(synth.clouds(B, bench.N_POINTS, seed=0)).to(dev)
st1 = torch.from_numpy(synth.fps_start(B, bench.N_POINTS, seed=1)).to(dev)
st2 = torch.zeros(B, dtype=torch.int64, device=dev)
model = sa_stack.SSGSetAbstractionStack().to(dev)
for i, sa in enumerate(model.layers_()):
    sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(bench.SA_CFG[i][3], bench.SA_CFG[i][4], seed=2 + i))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if os.environ.get("PAPC_GCLK_GRAPH", "1") == "1":
    # as bench.py runs the step: the captured forward replayed (the 3 warm-up passes and the capture take
    # the first slices; every replay rewrites the capture's slices -> the LAST 8 launches of the dump)
    g = sa_stack.GraphedForward(lambda x: model(x, None, start_idx=(st1, st2)), xyz)
    for _ in range(3):
        flush.fill_(1)                  # cold L2, as between the bench's timed iterations
        torch.cuda.synchronize()
        g.replay()
else:
    for _ in range(3):
        flush.fill_(1)
        torch.cuda.synchronize()
        model(xyz, None, start_idx=(st1, st2))
torch.cuda.synchronize()
os.makedirs(os.path.dirname(out), exist_ok=True)
n = _lib.lib().papc_tt_gclk_dump(out.encode())
print("launches dumped:", n, "->", out)
