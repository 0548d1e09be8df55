#!/bin/bash
# N-GPU bench check with tight timeouts (a hang must not burn the budget): both arms, as the driver launches them
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_r2_ref_n$N.json 2> gpurun_out/bench_r2_ref_n$N.err
echo "ref exit $?"; grep "^{" gpurun_out/bench_r2_ref_n$N.json | cut -c1-200
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "bench exit $?"; grep -v "^$" gpurun_out/bench_r2_n$N.err | grep -v "OMP_NUM\|\*\*\*" | tail -5 | cut -c1-300
python - <<P
import json
for l in open("gpurun_out/bench_r2_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["config"]["collective"])
        print("e2e", d["e2e"]["value"]); print("strong", d.get("strong_scaling")); print("syncbn", d.get("syncbn"))
P
if [[ "$N" == "2" ]]; then
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -s --timeout 280 -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
fi
