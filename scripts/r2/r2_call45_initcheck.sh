#!/bin/bash
mkdir -p gpurun_out/c45
timeout 800 compute-sanitizer --tool initcheck python scripts/sanitize_small.py > gpurun_out/c45/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/c45/sanitizer_initcheck.log; grep -c "Uninitialized" gpurun_out/c45/sanitizer_initcheck.log
