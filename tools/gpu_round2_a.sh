#!/bin/bash
# first GPU pass of round 2: gpu tests, short bench, sanitizer logs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json | cut -c1-600
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_memcheck.log
