#!/bin/bash
# round-2 GPU call A: parity of the warp-uniform issue loops, A/B against the round-1 kernel on the same box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
free -g > gpurun_out/a_mem.txt 2>&1; nproc >> gpurun_out/a_mem.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/a_bench_new.json 2> gpurun_out/a_bench_new.err; echo "bench new rc=$?"
VIQUAE_B200_LIB=$PWD/viquae_b200/libviquae_b200_prev.so timeout 600 python bench.py --no-cpu-baseline --no-parity-check > gpurun_out/a_bench_prev.json 2> gpurun_out/a_bench_prev.err; echo "bench prev rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-parity-check > gpurun_out/a_bench_new2.json 2> gpurun_out/a_bench_new2.err; echo "bench new2 rc=$?"
for nq in 256 4096; do
  timeout 600 python bench.py --no-cpu-baseline --no-parity-check --nq $nq --steps 10 > gpurun_out/a_bench_new_nq$nq.json 2>> gpurun_out/a_bench_new.err
  VIQUAE_B200_LIB=$PWD/viquae_b200/libviquae_b200_prev.so timeout 600 python bench.py --no-cpu-baseline --no-parity-check --nq $nq --steps 10 > gpurun_out/a_bench_prev_nq$nq.json 2>> gpurun_out/a_bench_prev.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 4 -c 1 -o gpurun_out/a_prof_tc_nq1250 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/a_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d.get('result_sha'), d.get('parity_check',{}).get('ok'), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
