#!/bin/bash
# round-2 GPU call V2 (1 GPU): ncu capture of the new K1 main pass (4 passes per search at 1.5 M rows: launch index 3); config 4 bench line
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:f32_topk -s 3 -c 1 -o gpurun_out/v_prof_f32_main -f python bench.py --rows 1500000 --dtype f32 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --config 4 --no-cpu-baseline > gpurun_out/v_bench_config4.json 2>> gpurun_out/v.err; tail -c 1200 gpurun_out/v_bench_config4.json
tail -3 gpurun_out/v.err
