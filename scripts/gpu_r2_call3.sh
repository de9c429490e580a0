#!/bin/bash
# Round 2, GPU call 3: cluster LSTM kernels (DSMEM h exchange) — GPU suite, config-3 bench (default and TPGSR_PREPACK=1), launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -12 $O/pytest_gpu.log
summ() { python scripts/bench_summary.py $O/bench_$1.json $1; }
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 20 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; summ $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
BARGS="--config 3" run_bench c3 TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1
BARGS="--config 3" run_bench c3_prepack $QUICK TPGSR_PREPACK=1
timeout 200 python scripts/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.log 2>&1; echo "hbm rc=$?"; grep -i "lstm\|fail" $O/hbm_kernels.log
env $QUICK TPGSR_BENCH_NO_ROOFLINE=1 TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c3.log 2>&1
echo "launches rc=$?"; wc -l $O/launches_c3.csv; python scripts/launch_summary.py $O/launches_c3.csv > $O/launches_c3_summary.txt 2>&1; head -32 $O/launches_c3_summary.txt
gzip -f $O/launches_c3.csv
du -sh $O
