#!/bin/bash
# round-2 GPU call Y (8 GPUs): final-tree confirmation of the sharded path — headline at N = 8 (result_sha must equal N = 1, 2) and config 5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench_n8.json 2> gpurun_out/y_n8.err
timeout 240 $TR bench.py --gpus 8 --config 5 --no-cpu-baseline > gpurun_out/y_bench_n8_config5.json 2>> gpurun_out/y_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/y_bench_n8.json").read().strip().splitlines()[-1])
print("N=8", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["result_sha"], d["parity_check"]["ok"])
d=json.loads(open("gpurun_out/y_bench_n8_config5.json").read().strip().splitlines()[-1])
print("config5 N=8", {k: d[k] for k in d if k in ("value", "ms_per_step")}, json.dumps(d.get("latency", d.get("config", {})))[:600])
PY
tail -2 gpurun_out/y_n8.err
