#!/bin/bash
# round-2 GPU call Q (8 GPUs): final build after the pull-merge load batching — headline at N = 8 (both transports), N = 1 on the same box, config 5 at 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/q_n8_peer.json 2> gpurun_out/q_n8_peer.err; echo "n8 peer rc=$?"
timeout 600 $TR bench.py --gpus 8 --config 5 > gpurun_out/q_n8_config5.json 2> gpurun_out/q_n8_config5.err; echo "n8 config5 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/q_n1.json 2> gpurun_out/q_n1.err; echo "n1 rc=$?"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --no-parity-check > gpurun_out/q_n4_peer.json 2> gpurun_out/q_n4_peer.err; echo "n4 rc=$?"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity-check > gpurun_out/q_n2_peer.json 2> gpurun_out/q_n2_peer.err; echo "n2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/q_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'kern', round(d.get('roofline',{}).get('kernel_ms',0) or 0,3), d.get('result_sha'), d.get('parity_check',{}).get('ok'), (d.get('clocks') or {}).get('sm_mhz'), 'launches', d.get('gpu_launches'))
        if 'latency' in d and 'per_batch_size' in d['latency']:
            for r in d['latency']['per_batch_size']: print('   ', r)
    except Exception as e:
        print(f, 'ERR', e)
PY
for f in gpurun_out/q_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $f | tail -n 3; done
