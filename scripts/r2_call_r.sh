#!/bin/bash
# round-2 GPU call R (1 GPU): the final build end to end — suite, smoke, every BASELINE config, reference arm, sweep, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r_smoke.log 2>&1; tail -1 gpurun_out/r_smoke.log
timeout 900 python bench.py > gpurun_out/r_default.json 2> gpurun_out/r_default.err; echo "default rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r_reference.json 2> gpurun_out/r_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --config 1 > gpurun_out/r_config1.json 2> gpurun_out/r_config1.err; echo "config1 rc=$?"
timeout 600 python bench.py --config 3 > gpurun_out/r_config3.json 2> gpurun_out/r_config3.err; echo "config3 rc=$?"
timeout 900 python bench.py --config 4 > gpurun_out/r_config4.json 2> gpurun_out/r_config4.err; echo "config4 rc=$?"
timeout 600 python bench.py --config 5 > gpurun_out/r_config5.json 2> gpurun_out/r_config5.err; echo "config5 rc=$?"
timeout 600 python bench.py --config 5 --rows 1500000 > gpurun_out/r_config5_shard.json 2> gpurun_out/r_config5_shard.err
for nq in 128 256 512 4096; do
  timeout 300 python bench.py --no-cpu-baseline --no-parity-check --nq $nq --steps 20 > gpurun_out/r_sweep_nq$nq.json 2>> gpurun_out/r_sweep.err
done
timeout 300 python bench.py --no-cpu-baseline --no-parity-check --rows 1500000 --steps 50 > gpurun_out/r_shard_nq1250.json 2>> gpurun_out/r_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|merge_|prep_rows" -c 60 --csv --log-file gpurun_out/r_launches_nq1250.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > /dev/null 2>> gpurun_out/r_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 4 -c 1 -o gpurun_out/r_prof_tc_nq1250 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_select -s 3 -c 1 -o gpurun_out/r_prof_merge_select -f python bench.py --rows 1500000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/r_ncu2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d.get('roofline',{}).get('frac',0),4), d.get('result_sha'), d.get('parity_check',{}).get('ok'), (d.get('clocks') or {}).get('sm_mhz'))
        if 'latency' in d and 'per_batch_size' in d['latency']:
            print('   ', [(r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
    except Exception as e:
        print(f, 'ERR', e)
PY
for f in gpurun_out/r_*.err; do echo "== $f"; tail -n 2 $f; done
bash scripts/r2_call_o.sh
