#!/bin/bash
# round-2 GPU call U (1 GPU): K1 with k-major tiled fp32 storage + TMA ring + FFMA2: full GPU suite, then same-box A/B against the
# round-1 build (gpurun_ab/lib_r1.so) through the engine
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/u_pytest.log
for rep in 1 2; do
VIQUAE_B200_LIB=$PWD/gpurun_ab/lib_r1.so timeout 300 python bench.py --rows 1500000 --dtype f32 --steps 10 --no-cpu-baseline --no-parity-check > gpurun_out/u_ab_r1_$rep.json 2>> gpurun_out/u.err
timeout 300 python bench.py --rows 1500000 --dtype f32 --steps 10 --no-cpu-baseline > gpurun_out/u_ab_kmajor_$rep.json 2>> gpurun_out/u.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_ab_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["result_sha"], (d.get("parity_check") or {}).get("ok"))
PY
tail -3 gpurun_out/u.err
