#!/bin/bash
# Round 2, GPU call 20: tensor-core LSTM recurrences (register-resident weight fragments, bulk-copy exchange) in the step:
# GPU suite, smoke, config 3 / 4 bench lines against the FFMA cluster kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name | cut -c1-150
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_ROOFLINE=1"
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3 $QUICK
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_ffma $QUICK TPGSR_LSTM_MODE=0
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_b $QUICK
BARGS="--config 4 --steps 20 --warmup 5" run_bench c4 $QUICK
BARGS="--config 5 --steps 30 --warmup 5" run_bench c5 $QUICK
du -sh $O
