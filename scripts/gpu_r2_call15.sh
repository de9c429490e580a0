#!/bin/bash
# Round 2, GPU call 15: weights-stationary plan with the two-tile first pass: timeline, flat tests, ncu --set full of the kernel,
# launch list of the config-3 step, GPU suite, default bench line (all legs).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/ws_timeline.py > $O/ws_timeline.log 2>&1; echo "ws_timeline rc=$?"; grep -v "ldtm\|staged" $O/ws_timeline.log | head -24
timeout 600 python -m pytest tests/test_gpu_flat.py -m gpu -q > $O/pytest_flat.log 2>&1; echo "pytest flat rc=$?"; tail -3 $O/pytest_flat.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ws_kernel -s 2 -c 2 -o /tmp/prof_ws -f python scripts/ws_once.py > $O/prof_ws.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_ws.ncu-rep --page raw --csv > $O/ncu_full_conv_ws_raw.csv 2>/dev/null; python scripts/ncu_pick.py $O/ncu_full_conv_ws_raw.csv > $O/ncu_full_conv_ws.csv; cat $O/ncu_full_conv_ws.csv
ncu -i /tmp/prof_ws.ncu-rep --page details --csv > $O/ncu_full_conv_ws_details.csv 2>/dev/null; wc -c $O/ncu_full_conv_ws_details.csv
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1"
env $QUICK TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c3.log 2>&1
echo "launches rc=$?"; python scripts/launch_summary.py $O/launches_c3.csv > $O/launches_c3_summary.txt 2>&1; head -24 $O/launches_c3_summary.txt
gzip -f $O/launches_c3.csv
timeout 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"; python scripts/bench_summary.py $O/bench_default.json default
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1"
env $QUICK timeout 500 python bench.py --config 2 --steps 30 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err; python scripts/bench_summary.py $O/bench_c2.json c2
du -sh $O
