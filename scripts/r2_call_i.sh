#!/bin/bash
# round-2 GPU call I (1 GPU): finer L2 sharing window sweep (nq = 1250 and 4096), latency after the memset-node removal
mkdir -p gpurun_out
for w in "1 2" "1 3" "2 2" "2 3" "2 4" "4 4" "4 5" "4 6" "3 6" "4 8"; do
  set -- $w
  VQ_SYNC_EVERY=$1 VQ_SYNC_WINDOW=$2 timeout 300 python bench.py --no-cpu-baseline --no-parity-check --steps 30 > gpurun_out/i_sync_$1_$2.json 2>> gpurun_out/i.err
done
for w in "2 3" "2 4" "4 6" "4 8"; do
  set -- $w
  VQ_SYNC_EVERY=$1 VQ_SYNC_WINDOW=$2 timeout 300 python bench.py --no-cpu-baseline --no-parity-check --nq 4096 --steps 10 > gpurun_out/i_nq4096_sync_$1_$2.json 2>> gpurun_out/i.err
  VQ_SYNC_EVERY=$1 VQ_SYNC_WINDOW=$2 timeout 300 python bench.py --no-cpu-baseline --no-parity-check --rows 1500000 --steps 50 > gpurun_out/i_shard_sync_$1_$2.json 2>> gpurun_out/i.err
done
timeout 600 python bench.py --config 5 --rows 1500000 --steps 500 > gpurun_out/i_c5_shard.json 2>> gpurun_out/i.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'latency' in d: print(f, [ (r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
        else: print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['result_sha'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 5 gpurun_out/i.err
