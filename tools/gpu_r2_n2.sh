#!/bin/bash
# 2-GPU check: NCCL SyncBN parity test, then the N=2 bench line (all-gather inside the graph, strong-scaling and SyncBN modes)
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -s --timeout 800 -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log; tail -8 gpurun_out/pytest_dist.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "bench exit $?"; grep -v "^$" gpurun_out/bench_r2_n$N.err | tail -8 | cut -c1-300
python - <<P
import json
d=json.load(open("gpurun_out/bench_r2_n$N.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["config"]["collective"])
print("e2e", d["e2e"]); print("strong", d.get("strong_scaling")); print("syncbn", d.get("syncbn"))
P
PAPC_GATHER_IN_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extra > gpurun_out/bench_r2_n${N}_gather_outside.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n${N}_gather_outside.json')); print('gather outside graph:', d['value'], d['ms_per_step'])"
