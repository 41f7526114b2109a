#!/bin/bash
# round-2 GPU call X (1 GPU): final verification of the tree — full GPU suite, smoke, default bench line, reference arm, fp32 ingest rate
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/x_bench_default.json 2>> gpurun_out/x.err
python - <<'PY'
import json, time, torch, sys
d=json.loads(open("gpurun_out/x_bench_default.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["result_sha"], d["parity_check"]["ok"], d["gpu_launches"])
sys.path.insert(0, ".")
from viquae_b200.index import B200FlatIndex
x = torch.randn(1_500_000, 768, device="cuda")
for dt in ("f32", "bf16"):
    idx = B200FlatIndex(d=768, dtype=dt, capacity=1_500_000)
    torch.cuda.synchronize(); t = time.time(); idx.add(x); torch.cuda.synchronize(); dt_s = time.time() - t
    print(f"ingest device-resident fp32 rows -> {dt}: {x.numel() * 4 / dt_s / 1e9:.0f} GB/s of source")
    t = time.time(); r = idx.reconstruct_n(0, 100000); print(f"  reconstruct_n 100k rows: {time.time() - t:.3f} s")
    idx.close()
PY
tail -3 gpurun_out/x.err
