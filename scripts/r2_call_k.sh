#!/bin/bash
# round-2 GPU call K (2 GPUs): suite incl. multi-GPU with the selection merges, A/B against the sort merges
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/k_pytest.log
for impl in select sort; do
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --rows 1500000 --steps 50 --no-cpu-baseline --no-parity-check > gpurun_out/k_shard_$impl.json 2>> gpurun_out/k.err
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-parity-check > gpurun_out/k_full_$impl.json 2>> gpurun_out/k.err
  VQ_MERGE_IMPL=$impl timeout 300 python bench.py --config 5 --rows 1500000 --steps 300 > gpurun_out/k_c5_$impl.json 2>> gpurun_out/k.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|merge_|prep_rows" -c 40 --csv --log-file gpurun_out/k_launches_shard.csv python bench.py --rows 1500000 --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > /dev/null 2>> gpurun_out/k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'latency' in d: print(f, [ (r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
        else: print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['result_sha'])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -E "merge|tc_topk" gpurun_out/k_launches_shard.csv | tail -10 | cut -d, -f5,9,15
tail -n 5 gpurun_out/k.err
