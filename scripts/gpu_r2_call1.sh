#!/bin/bash
# Round 2, first GPU call: (1) the whole GPU suite + the never-run experimental tests, (2) bench lines for configs 3 and 2 with the new legs
# (eager cuDNN baseline, inference, parity, CPU port), (3) A/B of the default-off switches on config 3, (4) the launch list of ONE graph-free
# config-3 step, (5) CUDA-event timings of the HBM kernels at B=48 and out-of-L2 batches + one ncu --set full pass over them.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $O/pytest_gpu.log
TPGSR_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q > $O/pytest_experimental.log 2>&1; echo "pytest experimental rc=$?"
tail -6 $O/pytest_experimental.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
    r, rc, eg, inf, par = d.get('roofline', {}), d.get('roofline_crnn', {}), d.get('gpu_eager_baseline', {}), d.get('inference', {}), d.get('parity', {})
    print(sys.argv[1], 'cfg', d['config']['config_id'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
          'launches', d['launches_per_step'], '| conv us', r.get('launch_us'), 'frac', r.get('frac'), '| crnn conv us', rc.get('launch_us'), 'frac', rc.get('frac'),
          '| eager', eg.get('value'), eg.get('error'), '| infer', inf.get('value'), inf.get('error'), '| parity', par.get('sr_rel_l2_vs_oracle'), par.get('psnr'), par.get('ssim'),
          '| cpu', d.get('cpu_baseline', {}).get('value'), '| loss', d['config']['last_loss'])
except Exception as e:
    print(sys.argv[1], 'no line:', e)
PY
}
run_bench() {   # name, args..., env via leading VAR=val words
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 $BARGS > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?"; tail -2 $O/bench_$name.err; summ $name
}
BARGS="--config 3" run_bench c3 A=1
BARGS="--config 2" run_bench c2 TPGSR_BENCH_NO_CPU=1
BARGS="--config 4" run_bench c4 TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1
QUICK="TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 TPGSR_BENCH_NO_INFER=1"
BARGS="--config 3" run_bench c3_occ2 $QUICK TPGSR_FLAT_OCC2=1
BARGS="--config 2" run_bench c2_occ2 $QUICK TPGSR_FLAT_OCC2=1
BARGS="--config 3" run_bench c3_bwd2 $QUICK TPGSR_BWD_PRODUCTS=2
BARGS="--config 3" run_bench c3_bwd2_w1 $QUICK TPGSR_BWD_PRODUCTS=2 TPGSR_WGRAD_PRODUCTS=1
BARGS="--config 3" run_bench c3_bf16 $QUICK TPGSR_CONV_MODE=bf16
timeout 300 python bench.py --impl reference --config 3 --steps 2 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref_c3.err; echo "ref arm rc=$?"; cat $O/bench_ref_c3.json | cut -c1-400
# launch list of one eager (graph-free) config-3 step: only the timed step is inside the profiler range
env $QUICK TPGSR_BENCH_NO_ROOFLINE=1 TPGSR_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-graph --no-e2e > $O/launches_c3.log 2>&1
echo "launches rc=$?"; wc -l $O/launches_c3.csv; python scripts/launch_summary.py $O/launches_c3.csv > $O/launches_c3_summary.txt 2>&1; head -45 $O/launches_c3_summary.txt
# HBM kernels: event timings, then one ncu --set full pass (one launch per kernel and size)
timeout 300 python scripts/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.log 2>&1; echo "hbm rc=$?"; cat $O/hbm_kernels.log
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'tps_sample|pixelshuffle_mish|image_loss|bicubic_gray|lstm_fwd|lstm_bwd|gru_fwd|gru_bwd' -o $O/prof_hbm -f \
    python scripts/hbm_kernels.py --once > $O/prof_hbm.log 2>&1; echo "ncu hbm rc=$?"
ls -la $O | head -60
