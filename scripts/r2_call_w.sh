#!/bin/bash
# round-2 GPU call W (1 GPU): compute-sanitizer over the rewritten K1 (TMA ring, tiled storage kernels); small-batch A/B of the fp32 mode
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for args in "f32 6000 40 96 20" "f32 3000 300 100 10"; do
  timeout 600 $CS --tool memcheck --error-exitcode 9 python scripts/gpu_check.py $args > gpurun_out/w_memcheck_$(echo $args | tr ' ' '_').log 2>&1; echo "memcheck $args rc=$?"
done
timeout 900 $CS --tool racecheck --error-exitcode 9 python scripts/gpu_check.py f32 6000 40 96 20 > gpurun_out/w_racecheck_f32.log 2>&1; echo "racecheck f32 rc=$?"
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_search.py -x -q -k "tiled" > gpurun_out/w_memcheck_tiled_test.log 2>&1; echo "memcheck tiled test rc=$?"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\| OK$\|passed\|failed" gpurun_out/w_*check*.log
for nq in 1 32 256; do
VIQUAE_B200_LIB=$PWD/gpurun_ab/lib_r1.so timeout 300 python bench.py --rows 1500000 --dtype f32 --nq $nq --steps 10 --no-cpu-baseline --no-parity-check > gpurun_out/w_ab_r1_nq$nq.json 2>> gpurun_out/w.err
timeout 300 python bench.py --rows 1500000 --dtype f32 --nq $nq --steps 10 --no-cpu-baseline --no-parity-check > gpurun_out/w_ab_new_nq$nq.json 2>> gpurun_out/w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/w_ab_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["result_sha"])
PY
tail -3 gpurun_out/w.err
