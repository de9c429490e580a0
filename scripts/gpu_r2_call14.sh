#!/bin/bash
# Round 2, GPU call 14: weights-stationary conv plan integrated: GPU suite, smoke, config 3 / 2 bench lines with the plan on and off.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1"
BARGS="--steps 30 --warmup 5" run_bench c3 $QUICK
BARGS="--steps 30 --warmup 5" run_bench c3_nows $QUICK TPGSR_FLAT_WS=0
BARGS="--config 2 --steps 30 --warmup 5" run_bench c2 $QUICK
BARGS="--config 2 --steps 30 --warmup 5" run_bench c2_nows $QUICK TPGSR_FLAT_WS=0
du -sh $O
