#!/bin/bash
mkdir -p gpurun_out/c43
python -m pytest tests -m gpu -q > gpurun_out/c43/gputest.log 2>&1; tail -4 gpurun_out/c43/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
