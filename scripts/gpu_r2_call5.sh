#!/bin/bash
# Round 2, GPU call 5: pipelined plan of the flat convs (tile after tile, per-box halo barriers, 4-warp epilogue under the next tile's MMAs):
# GPU suite, A/B against TPGSR_FLAT_PIPE=0 on configs 2 / 3, the reduced-product backward switches, per-CTA timeline, ncu of the conv.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_stats.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 $O/pytest_gpu.log
summ() { python scripts/bench_summary.py $O/bench_$1.json $1; }
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 20 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; summ $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
BARGS="--config 2" run_bench c2 $QUICK
BARGS="--config 2" run_bench c2_nopipe $QUICK TPGSR_FLAT_PIPE=0
BARGS="--config 3" run_bench c3 $QUICK
BARGS="--config 3" run_bench c3_nopipe $QUICK TPGSR_FLAT_PIPE=0
BARGS="--config 3" run_bench c3_bwd2 $QUICK TPGSR_BWD_PRODUCTS=2
BARGS="--config 3" run_bench c3_bwd2_w1 $QUICK TPGSR_BWD_PRODUCTS=2 TPGSR_WGRAD_PRODUCTS=1
BARGS="--config 5" run_bench c5 TPGSR_BENCH_NO_CPU=1
timeout 300 python scripts/flat_timeline.py > $O/flat_timeline.log 2>&1; echo "timeline rc=$?"; head -40 $O/flat_timeline.log
TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1 TPGSR_BENCH_NO_ROOFLINE=1 TPGSR_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none \
    -k regex:'conv_flat_kernel' -c 12 -o /tmp/prof_conv -f python bench.py --config 2 --steps 1 --warmup 3 --no-graph --no-e2e > $O/prof_conv.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_conv.ncu-rep --page raw --csv > $O/ncu_full_conv_flat_pipe.csv 2>/dev/null; wc -c $O/ncu_full_conv_flat_pipe.csv
du -sh $O
