#!/bin/bash
# Round 2, GPU call 4: valid/full flat convs (CRNN conv6), InfoGen on the tensor path, teacher on the side stream: GPU suite, the cascade
# parity test with fp32 CUDA-core contractions (is the percent-level cascade gradient error conditioning or a kernel bug?), benches, launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_stats.jsonl
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -12 $O/pytest_gpu.log
mv $O/parity_stats.jsonl $O/parity_stats_default.jsonl 2>/dev/null
TPGSR_CONV_MODE=fp32 timeout 600 python -m pytest tests/test_gpu_parity_full.py -m gpu -q > $O/pytest_parity_fp32.log 2>&1; echo "pytest parity fp32 rc=$?"; tail -5 $O/pytest_parity_fp32.log
mv $O/parity_stats.jsonl $O/parity_stats_fp32.jsonl 2>/dev/null
summ() { python scripts/bench_summary.py $O/bench_$1.json $1; }
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" timeout 500 python bench.py --steps 20 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; summ $name
}
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
BARGS="--config 3" run_bench c3 TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1
BARGS="--config 4" run_bench c4 $QUICK
BARGS="--config 2" run_bench c2 $QUICK
env $QUICK TPGSR_BENCH_NO_ROOFLINE=1 TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c3.log 2>&1
echo "launches rc=$?"; wc -l $O/launches_c3.csv; python scripts/launch_summary.py $O/launches_c3.csv > $O/launches_c3_summary.txt 2>&1; head -30 $O/launches_c3_summary.txt
gzip -f $O/launches_c3.csv
du -sh $O
