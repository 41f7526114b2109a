#!/bin/bash
# round-2 GPU call U (1 GPU): same-box A/B of the fp32 exact mode through the engine: round-1 K1 (HEAD build) vs the TMA-fed FFMA2 K1
mkdir -p gpurun_out
for rep in 1 2; do
VIQUAE_B200_LIB=$PWD/gpurun_ab/lib_r1.so timeout 300 python bench.py --rows 1500000 --dtype f32 --steps 10 --no-cpu-baseline --no-parity-check > gpurun_out/u_ab_r1_$rep.json 2>> gpurun_out/u.err
timeout 300 python bench.py --rows 1500000 --dtype f32 --steps 10 --no-cpu-baseline --no-parity-check > gpurun_out/u_ab_tma_$rep.json 2>> gpurun_out/u.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_ab_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["result_sha"])
PY
tail -3 gpurun_out/u.err
