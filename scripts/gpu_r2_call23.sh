#!/bin/bash
# Round 2, GPU call 23: event timings of the HBM / latency kernels after the round's last kernel changes + one ncu --set full pass
# (one launch per kernel and size) for the new kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.log 2>&1; echo "hbm rc=$?"; cat $O/hbm_kernels.log | cut -c1-150
timeout 900 ncu --set full --clock-control none -k regex:'lstm_fwd_mma|lstm_bwd_mma|gru_fwd_mma|gru_bwd_mma|maxpool_.*_v4|conv_c1_3x3|adam_v4' -o /tmp/prof_new -f \
    python scripts/hbm_kernels.py --once > $O/prof_new.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_new.ncu-rep --page raw --csv > /tmp/ncu_new_raw.csv 2>/dev/null; python scripts/ncu_pick.py /tmp/ncu_new_raw.csv > $O/ncu_new_kernels.csv; cut -c1-260 $O/ncu_new_kernels.csv
du -sh $O
