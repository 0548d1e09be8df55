#!/bin/bash
# round-2 evidence run (one gpurun call): full GPU test suite, bench line, ncu launch list of the bench command,
# one `ncu --set full` capture of every kernel of ONE warm step.  Outputs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_r2.txt 2>&1
if [[ "$*" == *onlyfull* ]]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mlp_layer_tt|sample_group|fps_reg|ball_query|point_moments|pool_finish|prep_wimg|prep_ximg" -s 19 -c 19 -f -o gpurun_out/r02_step_full python tools/prof_step.py 2 > gpurun_out/ncu_full_r2.log 2>&1
echo "ncu full exit $?"; exit 0
fi
if [[ "$*" != *notests* ]]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r2_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2_full.log; tail -4 gpurun_out/pytest_r2_full.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r2_final.err; cut -c1-400 gpurun_out/bench_r2_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_bench_r2.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"mlp_layer_tt|sample_group|fps_reg|ball_query|point_moments|pool_finish|prep_wimg|prep_ximg" -s 19 -c 19 -f \
   -o gpurun_out/r02_step_full python tools/prof_step.py 2 > gpurun_out/ncu_full_r2.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full_r2.log | cut -c1-200
ls -la gpurun_out/r02_step_full.ncu-rep
