#!/bin/bash
# round-2 GPU call S (1 GPU): single-pass mode for few queries (in-kernel pruning + two-level merge) vs the primed schedule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_dropin.py -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
for mx in 32 8 0; do
  VQ_SINGLE_PASS_NQ=$mx timeout 600 python bench.py --config 5 --rows 1500000 --steps 400 > gpurun_out/s_c5_shard_sp$mx.json 2>> gpurun_out/s.err
done
for mx in 32 0; do
  VQ_SINGLE_PASS_NQ=$mx timeout 600 python bench.py --config 5 --steps 300 > gpurun_out/s_c5_12m_sp$mx.json 2>> gpurun_out/s.err
done
VQ_SINGLE_PASS_NQ=128 timeout 300 python bench.py --nq 64 --rows 1500000 --steps 100 --no-cpu-baseline > gpurun_out/s_nq64_sp.json 2>> gpurun_out/s.err
VQ_SINGLE_PASS_NQ=0 timeout 300 python bench.py --nq 64 --rows 1500000 --steps 100 --no-cpu-baseline --no-parity-check > gpurun_out/s_nq64_mp.json 2>> gpurun_out/s.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|merge_|prep_rows" -c 30 --csv --log-file gpurun_out/s_launches_nq1.csv python bench.py --config 5 --rows 1500000 --steps 2 --warmup 1 --nq 1 > /dev/null 2>> gpurun_out/s.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    if "latency" in d: print(f, [(r["nq"], r["p50_us"], r["p99_us"]) for r in d["latency"]["per_batch_size"]], d.get("result_sha_per_batch_size",{}).get("32"))
    else: print(f, round(d["ms_per_step"],3), d["result_sha"], d.get("parity_check",{}).get("ok"))
PY
tail -8 gpurun_out/s_launches_nq1.csv | cut -d, -f5,9,15
tail -5 gpurun_out/s.err
