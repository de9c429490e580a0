#!/bin/bash
# Round 2, GPU call 21: float4 max-pool / Adam kernels, 32-bit index arithmetic in split_bf16_flat: GPU suite, config 2 / 3 lines, launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name | cut -c1-150
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_ROOFLINE=1"
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3 $QUICK
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_b $QUICK
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2 $QUICK
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1"
env $QUICK TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c3.log 2>&1
echo "launches rc=$?"; python scripts/launch_summary.py $O/launches_c3.csv > $O/launches_c3_summary.txt 2>&1; head -36 $O/launches_c3_summary.txt
gzip -f $O/launches_c3.csv
du -sh $O
