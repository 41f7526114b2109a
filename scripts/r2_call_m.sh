#!/bin/bash
# round-2 GPU call M (1 GPU): programmatic dependent launch in the latency regime: suite + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/m_pytest.log
for rows in 1500000 12000000; do
  timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/m_c5_${rows}_pdl.json 2>> gpurun_out/m.err
  VQ_NO_PDL=1 timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/m_c5_${rows}_nopdl.json 2>> gpurun_out/m.err
done
timeout 300 python bench.py --nq 128 --no-cpu-baseline > gpurun_out/m_nq128_pdl.json 2>> gpurun_out/m.err
VQ_NO_PDL=1 timeout 300 python bench.py --nq 128 --no-cpu-baseline --no-parity-check > gpurun_out/m_nq128_nopdl.json 2>> gpurun_out/m.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    if "latency" in d: print(f, [(r["nq"], r["p50_us"], r["p99_us"]) for r in d["latency"]["per_batch_size"]], d.get("result_sha_per_batch_size",{}).get("32"))
    else: print(f, round(d["ms_per_step"],3), d["result_sha"], d.get("parity_check",{}).get("ok"))
PY
tail -5 gpurun_out/m.err
