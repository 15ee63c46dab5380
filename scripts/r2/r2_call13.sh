#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "msm or sharded or golden" > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest13.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_compute_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/r2_compute_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2_compute_sanitizer_synccheck.log
