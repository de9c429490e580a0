#!/bin/bash
# which kernel change moves the worst gradient tensors of the stu_iter=3 cascade test (conditioning probe)
mkdir -p gpurun_out
for cfg in "A=1" "TPGSR_LSTM_MODE=0" "TPGSR_GRU_MODE=0" "TPGSR_FLAT_WS=0" "TPGSR_LSTM_MODE=0 TPGSR_GRU_MODE=0 TPGSR_FLAT_WS=0"; do
  rm -f gpurun_out/parity_stats.jsonl
  env $cfg timeout 300 python -m pytest tests/test_gpu_parity_full.py -m gpu -q -k "stu_iter3" > gpurun_out/probe.log 2>&1
  echo "== $cfg: $(tail -1 gpurun_out/probe.log)"
  python - <<'PY'
import json
for line in open('gpurun_out/parity_stats.jsonl'):
    d = json.loads(line); v = d.get('stats', d)
    if v.get('worst_vs_floor'):
        print('   median err', round(v['median_err'][0], 4), 'worst ratios', [(round(a, 2), round(b, 3), n) for a, b, c, n in v['worst_vs_floor'][:4]])
PY
done
