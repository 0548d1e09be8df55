#!/bin/bash
# A/B library builds inside ONE gpurun call: bench every listed .so `reps` times, interleaved.
# usage: gpurun --timeout 900 -- 'bash tools/ab_lib.sh "papc_b200/lib/libpapc_b200.so papc_b200/lib/libpapc_b200_x.so" [reps=2] [steps=30]'
libs="$1"; reps="${2:-2}"; steps="${3:-30}"
for ((r = 0; r < reps; ++r)); do
  for l in $libs; do
    ms=$(PAPC_B200_LIB=$l timeout 300 python bench.py --steps "$steps" --warmup 5 --no-extra 2>/dev/null |
         python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('%.4f ms/step  %.2f M points/s  e2e %.4f ms' % (d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step']))")
    echo "$(basename $l)  run $r:  $ms"
  done
done
