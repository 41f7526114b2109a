#!/bin/bash
# round-2 GPU call G (1 GPU): suite after the latency-mode merges / stream-kernel removal, latency regime again
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
for rows in 1500000 12000000; do
  timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/g_c5_${rows}_auto.json 2>> gpurun_out/g.err
  VQ_PASS_RATIO=8 timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/g_c5_${rows}_ratio8.json 2>> gpurun_out/g.err
done
timeout 300 python bench.py --nq 128 --no-cpu-baseline --no-parity-check > gpurun_out/g_nq128.json 2>> gpurun_out/g.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/g_launches_nq1_shard.csv python bench.py --config 5 --rows 1500000 --steps 2 --warmup 1 --nq 1 > /dev/null 2>> gpurun_out/g.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/g_smoke.log 2>&1; tail -2 gpurun_out/g_smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/g_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'latency' in d: print(f, [ (r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
        else: print(f, d['ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 5 gpurun_out/g.err
