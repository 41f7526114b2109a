#!/bin/bash
# round-2 GPU call V (1 GPU): tiled-fp32-storage test; ncu capture of the new K1 main pass; config 1 bench line (f32x3 / f32 exact / bf16)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "tiled or fp32_mode or zero_norm or faiss_flat or ties" > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:f32_topk -s 4 -c 1 -o gpurun_out/v_prof_f32_main -f python bench.py --rows 1500000 --dtype f32 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --config 1 > gpurun_out/v_bench_config1.json 2>> gpurun_out/v.err; tail -c 1500 gpurun_out/v_bench_config1.json
tail -3 gpurun_out/v.err
