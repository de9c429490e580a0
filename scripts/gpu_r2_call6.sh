#!/bin/bash
# Round 2, GPU call 6: regrouped TPS kernels (+ mask=False), backward-products tests in the default suite; HBM-kernel timings at B=48 and
# out-of-L2 batches; one ncu --set full pass over them (raw CSV exported on the box: the .ncu-rep stays there, gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_stats.jsonl
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 $O/pytest_gpu.log
timeout 300 python scripts/hbm_kernels.py --json $O/hbm_kernels.json > $O/hbm_kernels.log 2>&1; echo "hbm rc=$?"; cat $O/hbm_kernels.log
timeout 900 ncu --set full --clock-control none -k regex:'tps_sample|pixelshuffle_mish|image_loss_fwd|image_loss_bwd|bicubic_gray|lstm_fwd|lstm_bwd|gru_fwd|gru_bwd' \
    -o /tmp/prof_hbm -f python scripts/hbm_kernels.py --once > $O/prof_hbm.log 2>&1; echo "ncu hbm rc=$?"
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > /tmp/prof_hbm_raw.csv 2>/dev/null
python scripts/ncu_pick.py /tmp/prof_hbm_raw.csv > $O/ncu_hbm_kernels.csv; wc -l $O/ncu_hbm_kernels.csv; head -50 $O/ncu_hbm_kernels.csv
python scripts/bench_summary.py /dev/null x >/dev/null 2>&1
TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 timeout 400 python bench.py --config 3 --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; python scripts/bench_summary.py $O/bench_c3.json c3
du -sh $O
