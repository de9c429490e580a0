#!/bin/bash
# Round 2, GPU call 10: tall plan of the flat convs (kw-stacked 9x9 convs as 2-D halo tiles): flat / ops / model suites, A/B on configs 2, 3, 4.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_flat.py -m gpu -q -x > $O/pytest_flat.log 2>&1; echo "pytest flat rc=$?"; tail -6 $O/pytest_flat.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 30 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1"
BARGS="--config 2" run_bench c2 $QUICK
BARGS="--config 2" run_bench c2_notall $QUICK TPGSR_FLAT_TALL=0
BARGS="--config 3" run_bench c3 $QUICK
BARGS="--config 3" run_bench c3_notall $QUICK TPGSR_FLAT_TALL=0
BARGS="--config 4" run_bench c4 $QUICK
env $QUICK TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c2.csv python bench.py --config 2 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c2.log 2>&1
echo "launches rc=$?"; python scripts/launch_summary.py $O/launches_c2.csv > $O/launches_c2_summary.txt 2>&1; head -30 $O/launches_c2_summary.txt
gzip -f $O/launches_c2.csv
du -sh $O
