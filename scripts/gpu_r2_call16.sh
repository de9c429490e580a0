#!/bin/bash
# Round 2, GPU call 16: A/B of the first-pass width of the weights-stationary plan (2 tiles vs 1) on the timeline and on configs 2 / 3.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/ws_timeline.py > $O/ws_timeline.log 2>&1; echo "ws_timeline rc=$?"; grep "B=" $O/ws_timeline.log
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; python scripts/bench_summary.py $O/bench_$name.json $name | cut -c1-260
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
for rep in 1 2; do
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2_first2_$rep $QUICK
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2_first1_$rep $QUICK TPGSR_WS_FIRST=1
BARGS="--config 2 --steps 40 --warmup 5" run_bench c2_nows_$rep $QUICK TPGSR_FLAT_WS=0
done
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_first2 $QUICK
BARGS="--config 3 --steps 40 --warmup 5" run_bench c3_first1 $QUICK TPGSR_WS_FIRST=1
du -sh $O
