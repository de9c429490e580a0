#!/bin/bash
# Round 2, GPU call 18: tensor-core GRU forward recurrence (mma.sync, split bf16): timing against the FFMA kernels + the GRU tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/time_gru.py > $O/time_gru.log 2>&1; echo "time_gru rc=$?"; cat $O/time_gru.log | tail -30
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gru" > $O/pytest_gru.log 2>&1; echo "pytest gru rc=$?"; tail -8 $O/pytest_gru.log
