#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2_compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_compute_sanitizer_racecheck.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -c 1800 gpurun_out/r2_bench_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
