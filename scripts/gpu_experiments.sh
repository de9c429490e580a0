#!/bin/bash
# One gpurun call for the switches that were written but not measured in round 1 (each default-off):
#   TPGSR_FLAT_OCC2=1      flat convs tiled for two co-resident CTAs per SM        TPGSR_PREPACK=1  persistent packed weights
#   TPGSR_BWD_PRODUCTS=2   gradient GEMMs with dY as one bf16 plane (two products)     TPGSR_WGRAD_PRODUCTS=1  one-product weight gradient
# First the kernels' own tests (tests/test_gpu_zz_experimental.py), then the flat-conv / model suites with a switch on, then bench.py
# (no CPU leg) beside the default.
set -u
mkdir -p gpurun_out
run_bench() {   # name, env...
  local name=$1; shift
  env "$@" TPGSR_BENCH_NO_CPU=1 timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "bench $name rc=$?"; python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
    r = d.get('roofline', {})
    print(sys.argv[1], 'images/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'conv us', r.get('launch_us'), 'frac', r.get('frac'))
except Exception as e:
    print(sys.argv[1], 'no line:', e)
PY
}
TPGSR_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q > gpurun_out/pytest_experimental.log 2>&1; echo "pytest experimental rc=$?"
tail -5 gpurun_out/pytest_experimental.log
TPGSR_FLAT_OCC2=1 timeout 400 python -m pytest tests/test_gpu_flat.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/pytest_occ2.log 2>&1; echo "pytest occ2 rc=$?"
tail -3 gpurun_out/pytest_occ2.log
run_bench default TPGSR_FLAT_OCC2=0 TPGSR_BENCH_INFER=1
run_bench occ2 TPGSR_FLAT_OCC2=1
run_bench occ2_bf16 TPGSR_FLAT_OCC2=1 TPGSR_CONV_MODE=bf16
run_bench default_bf16 TPGSR_CONV_MODE=bf16
run_bench bwd2 TPGSR_BWD_PRODUCTS=2
run_bench bwd2_w1 TPGSR_BWD_PRODUCTS=2 TPGSR_WGRAD_PRODUCTS=1
run_bench bwd2_w1_occ2 TPGSR_BWD_PRODUCTS=2 TPGSR_WGRAD_PRODUCTS=1 TPGSR_FLAT_OCC2=1
TPGSR_BWD_PRODUCTS=2 TPGSR_WGRAD_PRODUCTS=1 timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_cascade.py -m gpu -q > gpurun_out/pytest_bwd2_w1.log 2>&1; echo "pytest bwd2+w1 rc=$? (the default suite's gradient tolerances applied to the reduced-product backward)"
tail -5 gpurun_out/pytest_bwd2_w1.log
