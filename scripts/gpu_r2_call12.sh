#!/bin/bash
# Round 2, GPU call 12: first run of the weights-stationary conv plan (weights resident in tensor memory): correctness + timeline, the flat
# conv tests, then the suite and the default bench if that is green.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/ws_timeline.py > $O/ws_timeline.log 2>&1; echo "ws_timeline rc=$?"; cat $O/ws_timeline.log | tail -40
timeout 600 python -m pytest tests/test_gpu_flat.py tests/test_gpu_ops.py -m gpu -q -x > $O/pytest_flat.log 2>&1; echo "pytest flat rc=$?"; tail -8 $O/pytest_flat.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 $O/pytest_gpu.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1"
BARGS="--steps 30 --warmup 5" run_bench c3 $QUICK
BARGS="--steps 30 --warmup 5" run_bench c3_nows $QUICK TPGSR_FLAT_WS=0
BARGS="--config 2 --steps 30 --warmup 5" run_bench c2 $QUICK
du -sh $O
