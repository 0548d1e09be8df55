#!/bin/bash
# round-2 GPU check: new tests (full-size chains, reference model files), then the bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_reference_models.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_r2_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r2_new.log
grep -v "^$" gpurun_out/pytest_r2_new.log | tail -60 | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
echo "bench exit $?"; tail -5 gpurun_out/bench_r2.err; cut -c1-1500 gpurun_out/bench_r2.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
echo "ref exit $?"; cut -c1-400 gpurun_out/bench_r2_ref.json
