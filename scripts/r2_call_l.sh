#!/bin/bash
# round-2 GPU call L (1 GPU): selection merge (restructured gather) vs sort merge
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -x -q 2>&1 | tail -3
for impl in select sort; do
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --rows 1500000 --steps 50 --no-cpu-baseline --no-parity-check > gpurun_out/l_shard_$impl.json
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --config 5 --rows 1500000 --steps 300 > gpurun_out/l_c5_$impl.json
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-parity-check > gpurun_out/l_full_$impl.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"merge_" -c 20 --csv --log-file gpurun_out/l_launches.csv python bench.py --rows 1500000 --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > /dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    if "latency" in d: print(f, [(r["nq"], r["p50_us"]) for r in d["latency"]["per_batch_size"]])
    else: print(f, round(d["ms_per_step"],3), d["result_sha"])
PY
tail -6 gpurun_out/l_launches.csv | cut -d, -f5,9,15
