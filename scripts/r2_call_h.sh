#!/bin/bash
# round-2 GPU call H (1 GPU): L2 sharing window sweep at the headline workload, launch list + full capture of the final build
mkdir -p gpurun_out
for w in "4 8" "2 4" "2 6" "4 6" "4 12" "8 16"; do
  set -- $w
  VQ_SYNC_EVERY=$1 VQ_SYNC_WINDOW=$2 timeout 300 python bench.py --no-cpu-baseline --no-parity-check --steps 30 > gpurun_out/h_sync_$1_$2.json 2>> gpurun_out/h.err
done
timeout 300 python bench.py --no-cpu-baseline --no-parity-check --steps 30 > gpurun_out/h_sync_default_again.json 2>> gpurun_out/h.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|merge_lists|prep_rows" -c 60 --csv --log-file gpurun_out/h_launches_nq1250.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check > /dev/null 2>> gpurun_out/h.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 4 -c 1 -o gpurun_out/h_prof_tc_nq1250 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/h_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 3 -c 1 -o gpurun_out/h_prof_tc_nq1 -f python bench.py --nq 1 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/h_ncu1.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/h_sync_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['result_sha'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 5 gpurun_out/h.err
