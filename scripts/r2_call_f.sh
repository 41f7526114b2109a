#!/bin/bash
# round-2 GPU call F (1 GPU): suite + latency schedule A/B + does the mma.sync streaming kernel win anywhere?
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
for rows in 1500000 12000000; do
  timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/f_c5_${rows}_auto.json 2>> gpurun_out/f.err
  VQ_PASS_RATIO=8 timeout 600 python bench.py --config 5 --rows $rows --steps 500 > gpurun_out/f_c5_${rows}_ratio8.json 2>> gpurun_out/f.err
done
for rows in 20000 100000 500000; do
  for v in auto stream; do
    timeout 300 python bench.py --config 5 --rows $rows --steps 300 --variant $v > gpurun_out/f_small_${rows}_${v}.json 2>> gpurun_out/f.err
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/f_launches_nq1_shard.csv python bench.py --config 5 --rows 1500000 --steps 2 --warmup 1 --nq 1 > /dev/null 2>> gpurun_out/f.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, [ (r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 5 gpurun_out/f.err
