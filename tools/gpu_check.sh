#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list.  Logs land in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tests] [bench] [ncu]'
mkdir -p gpurun_out
what="${*:-tests smoke bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ "$what" == *tc* ]]; then
  timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
  echo "pytest_tc exit $?" >> gpurun_out/pytest_tc.log
  tail -80 gpurun_out/pytest_tc.log | cut -c1-400
fi
if [[ "$what" == *tests* ]]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_tc.py -p no:cacheprovider > gpurun_out/pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest.log
  tail -60 gpurun_out/pytest.log
fi
if [[ "$what" == *layers* ]]; then
  timeout 300 python tools/prof_layer.py > gpurun_out/layers.log 2>&1
  echo "layers exit $?"; tail -12 gpurun_out/layers.log
  PAPC_MLP_TC=0 timeout 300 python tools/prof_layer.py > gpurun_out/layers_simt.log 2>&1
fi
if [[ "$what" == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
fi
if [[ "$what" == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [[ "$what" == *ncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -3 gpurun_out/ncu_bench.log | cut -c1-300
fi
if [[ "$what" == *triage* ]]; then
  for d in 0 1 2 4 3 5 6 7; do
    echo "== PAPC_TT_DBG=$d" >> gpurun_out/triage.log
    PAPC_TT_DBG=$d timeout 300 python tools/prof_layer.py sa1.l3 sa2.l2 sa1.l2 >> gpurun_out/triage.log 2>&1
  done
  cat gpurun_out/triage.log
fi
if [[ "$what" == *full* ]]; then
  # one --set full capture of every grouped-MLP layer kernel of ONE warm step (the second pass)
  timeout 1200 ncu --set full --clock-control none --import-source on \
     -k regex:"mlp_layer_tt|fps_reg" -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-10} -f -o gpurun_out/step_full python tools/prof_step.py 2 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-300
fi
if [[ "$what" == *prims* ]]; then
  timeout 300 python tools/prof_prims.py > gpurun_out/prims.log 2>&1
  echo "--- PAPC_FPS_WIDE=1" >> gpurun_out/prims.log
  PAPC_FPS_WIDE=1 timeout 300 python tools/prof_prims.py >> gpurun_out/prims.log 2>&1
  cat gpurun_out/prims.log
fi
if [[ "$what" == *fpsshape* ]]; then
  for w in 0 2 3; do echo "--- PAPC_FPS_WIDE=$w"; PAPC_FPS_WIDE=$w timeout 300 python tools/prof_prims.py 2>&1 | grep -v 2048; done | tee gpurun_out/fps_shapes.log
fi
if [[ "$what" == *pillarncu* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vox_|pfn_|scatter_" -s 13 -c 13 -f \
     -o gpurun_out/pillars_full python tools/prof_pillars.py > gpurun_out/ncu_pillars.log 2>&1
  echo "ncu pillars exit $?"; tail -2 gpurun_out/ncu_pillars.log | cut -c1-200
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ball_query" -s 2 -c 2 -f \
     -o gpurun_out/bq_full python tools/prof_step.py 2 > gpurun_out/ncu_bq.log 2>&1
  echo "ncu bq exit $?"
fi
