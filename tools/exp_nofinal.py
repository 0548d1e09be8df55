"""Upper bound of what removing the last-CTA finalisation from the layer kernels could gain: one normal pass
writes scale / shift, then PAPC_TT_DBG=1024 (experiment build) skips every finalisation and the replayed graph
is timed.  usage: PAPC_B200_LIB=papc_b200/lib/libpapc_b200_exp.so python tools/exp_nofinal.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from papc_b200 import sa_stack, synth
dev = torch.device("cuda:0")
B = 32
xyz = torch.from_numpy(synth.clouds(B, 1024, seed=0)).to(dev)
st1 = torch.from_numpy(synth.fps_start(B, 1024, seed=1)).to(dev)
st2 = torch.zeros(B, dtype=torch.int64, device=dev)
model = sa_stack.SSGSetAbstractionStack().to(dev)
for i, sa in enumerate(model.layers_()):
    sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(bench.SA_CFG[i][3], bench.SA_CFG[i][4], seed=2 + i))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(tag):
    g = sa_stack.GraphedForward(lambda x: model(x, None, start_idx=(st1, st2)), xyz)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    t = 0.0
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    print(f"{tag}: {t/30*1e3:.1f} us per step")
run("normal")
os.environ["PAPC_TT_DBG"] = "1024"
run("no finalisation (stale scale/shift)")
