#!/bin/bash
# round-2 GPU call T (1 GPU): pass-ratio sweep at 1,250 queries (shard and full corpus) with the selection merges; final suite + smoke
mkdir -p gpurun_out
for r in 8 10 12 16; do
  VQ_PASS_RATIO=$r timeout 300 python bench.py --rows 1500000 --steps 60 --no-cpu-baseline --no-parity-check > gpurun_out/t_shard_ratio$r.json 2>> gpurun_out/t.err
done
for r in 8 10 12; do
  VQ_PASS_RATIO=$r timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-parity-check > gpurun_out/t_full_ratio$r.json 2>> gpurun_out/t.err
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["result_sha"])
PY
tail -3 gpurun_out/t.err
