#!/bin/bash
# round-2 GPU call D (1 GPU): tests after the fusion-grid / merge / pass-schedule changes, config 3 with the fit leg, latency and shard-size A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/d_pytest.log
timeout 600 python bench.py --config 3 > gpurun_out/d_config3.json 2> gpurun_out/d_config3.err; echo "config3 rc=$?"
timeout 600 python bench.py --config 5 > gpurun_out/d_config5.json 2> gpurun_out/d_config5.err; echo "config5 rc=$?"
timeout 600 python bench.py --config 5 --rows 1500000 > gpurun_out/d_config5_shard.json 2> gpurun_out/d_config5_shard.err; echo "config5 shard rc=$?"
timeout 600 python bench.py --rows 1500000 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/d_shard_auto.json 2> gpurun_out/d_shard.err
VQ_MERGE_THREADS=256 timeout 600 python bench.py --rows 1500000 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/d_shard_m256.json 2>> gpurun_out/d_shard.err
VQ_MERGE_THREADS=128 timeout 600 python bench.py --rows 1500000 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/d_shard_m128.json 2>> gpurun_out/d_shard.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/d_default.json 2> gpurun_out/d_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), 'kern', round(d.get('roofline',{}).get('kernel_ms',0) or 0,3), d.get('result_sha'), d.get('parity_check',{}).get('ok'), (d.get('clocks') or {}).get('sm_mhz'))
        if 'latency' in d and 'per_batch_size' in d['latency']:
            for r in d['latency']['per_batch_size']: print('   ', r)
        if 'fusion_fit' in d: print('   ', d['fusion_fit']); print('   ', d['stages_ms'])
    except Exception as e:
        print(f, 'ERR', e)
PY
for f in gpurun_out/d_*.err; do echo "== $f"; tail -n 3 $f; done
