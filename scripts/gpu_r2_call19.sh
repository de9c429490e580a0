#!/bin/bash
# Round 2, GPU call 19: tensor-core GRU recurrences in the step: GPU suite, smoke, config 2 / 3 bench lines against the FFMA kernels.
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
for rep in 1 2; do
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2_$rep $QUICK
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2_ffma_$rep $QUICK TPGSR_GRU_MODE=0
done
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3 $QUICK
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_ffma $QUICK TPGSR_GRU_MODE=0
du -sh $O
