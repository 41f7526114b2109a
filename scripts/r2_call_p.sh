#!/bin/bash
# round-2 GPU call P (2 GPUs): multi-GPU tests after the load-batching fix of the pull-merge, N = 2 bench and latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_search.py -x -q 2>&1 | tail -3
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/p_n2_peer.json 2> gpurun_out/p_n2.err; echo "n2 rc=$?"
VIQUAE_B200_TRANSPORT=nccl timeout 600 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity-check > gpurun_out/p_n2_nccl.json 2>> gpurun_out/p_n2.err
timeout 600 $TR2 bench.py --gpus 2 --config 5 --rows 3000000 > gpurun_out/p_n2_config5.json 2>> gpurun_out/p_n2.err; echo "n2 config5 rc=$?"
timeout 600 python bench.py --config 5 --rows 1500000 > gpurun_out/p_n1_config5_shard.json 2>> gpurun_out/p_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/p_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'kern', round(d.get('roofline',{}).get('kernel_ms',0) or 0,3), d.get('result_sha'), d.get('parity_check',{}).get('ok'))
        if 'latency' in d and 'per_batch_size' in d['latency']:
            print('   ', [(r['nq'], r['p50_us'], r['p99_us']) for r in d['latency']['per_batch_size']])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/p_n2.err | tail -5
