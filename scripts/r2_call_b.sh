#!/bin/bash
# round-2 GPU call B: ridge profile (nq=256), refreshed sweep, the other BASELINE configs on one GPU
mkdir -p gpurun_out
for nq in 128 192 256 384 512 1024; do
  timeout 300 python bench.py --no-cpu-baseline --no-parity-check --nq $nq --steps 20 > gpurun_out/b_sweep_nq$nq.json 2>> gpurun_out/b_sweep.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 4 -c 1 -o gpurun_out/b_prof_tc_nq256 -f python bench.py --nq 256 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/b_ncu256.log 2>&1; echo "ncu256 rc=$?"
timeout 600 python bench.py --config 1 > gpurun_out/b_config1.json 2> gpurun_out/b_config1.err; echo "config1 rc=$?"
timeout 600 python bench.py --config 3 > gpurun_out/b_config3.json 2> gpurun_out/b_config3.err; echo "config3 rc=$?"
timeout 900 python bench.py --config 4 > gpurun_out/b_config4.json 2> gpurun_out/b_config4.err; echo "config4 rc=$?"
timeout 600 python bench.py --config 5 > gpurun_out/b_config5.json 2> gpurun_out/b_config5.err; echo "config5 rc=$?"
timeout 900 python bench.py > gpurun_out/b_default.json 2> gpurun_out/b_default.err; echo "default rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_reference.json 2> gpurun_out/b_reference.err; echo "reference rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d.get('roofline',{}).get('frac',0),4), d.get('result_sha'), d.get('parity_check',{}).get('ok'), (d.get('clocks') or {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/b_*.err
