#!/bin/bash
# A/B one environment switch of the library inside ONE gpurun call: every value is benched `reps` times,
# interleaved, and the ms/step of each run is printed (the switches are read at launch / capture time, so
# one process per value).
# usage: gpurun --timeout 900 -- 'bash tools/ab_bench.sh PAPC_TT_DBG 0,256 [reps=2] [steps=30]'
#   PAPC_TT_DBG bits that are A/B switches in normal builds: 256 = BatchNorm sums from the partial rows
#   instead of the integer-atomic words, 512 = no tensor-map prefetch;  PAPC_TT_TMA2D=0 = per-row bulk
#   copies instead of the 2-D tensor copy;  PAPC_TT_PDL=0 = no programmatic dependent launch;
#   PAPC_OVERLAP_SAMPLING=0 = FPS / ball query on the main stream.
var="$1"; IFS=',' read -ra vals <<< "$2"; reps="${3:-2}"; steps="${4:-30}"
for ((r = 0; r < reps; ++r)); do
  for v in "${vals[@]}"; do
    ms=$(env "$var=$v" timeout 300 python bench.py --steps "$steps" --warmup 5 2>/dev/null |
         python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms/step  %.2f M points/s  e2e %.4f ms' % (d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step']))")
    echo "$var=$v  run $r:  $ms"
  done
done
