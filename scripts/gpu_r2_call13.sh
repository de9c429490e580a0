#!/bin/bash
# Round 2, GPU call 13: weights-stationary conv plan, second version (TMA weight ring, per-tap readiness, division-free issue loop,
# vectorized epilogue): timeline + the flat conv tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/ws_timeline.py > $O/ws_timeline.log 2>&1; echo "ws_timeline rc=$?"; cat $O/ws_timeline.log | tail -40
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q > $O/pytest_flat.log 2>&1; echo "pytest flat rc=$?"; tail -15 $O/pytest_flat.log
du -sh $O
