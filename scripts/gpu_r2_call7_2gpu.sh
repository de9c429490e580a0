#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): NCCL parity of the product train steps (replicas bit-identical, equal to the single-process emulation),
# the single-GPU suite after the fast-mish / TPS changes, and 2-GPU bench lines for configs 3 and 4 (NCCL_DEBUG left as the launcher sets it).
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 $O/pytest_multi.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/pytest_gpu.log
for cfg in 3 4; do
  NCCL_DEBUG=INFO TPGSR_BENCH_CONFIG=$cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_c$cfg.json 2> $O/bench_n2_c$cfg.err
  echo "bench n2 c$cfg rc=$?"; python scripts/bench_summary.py $O/bench_n2_c$cfg.json n2_c$cfg; grep -c "NCCL INFO" $O/bench_n2_c$cfg.err; grep -m2 "nranks\|comm .* rank" $O/bench_n2_c$cfg.err | cut -c1-200
  wc -l $O/bench_n2_c$cfg.json
done
TPGSR_BENCH_NO_CPU=1 TPGSR_BENCH_NO_EAGER=1 timeout 400 python bench.py --config 3 --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; python scripts/bench_summary.py $O/bench_c3.json c3
tail -c 3000 $O/bench_n2_c3.err > $O/bench_n2_c3.err.tail; rm -f $O/bench_n2_c3.err $O/bench_n2_c4.err
du -sh $O
