#!/bin/bash
# Round 2, GPU call 26: ncu --set full rows of the final conv kernels at the bench shape (weights-stationary forward, stacked weight gradient).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'conv_ws_kernel|conv_wgrad_flat_kernel' -s 4 -c 6 -o /tmp/prof_fin -f python scripts/flat_timeline.py > $O/prof_fin.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_fin.ncu-rep --page raw --csv > /tmp/ncu_fin_raw.csv 2>/dev/null; python scripts/ncu_pick.py /tmp/ncu_fin_raw.csv > $O/ncu_final_conv_kernels.csv; cut -c1-300 $O/ncu_final_conv_kernels.csv
