#!/bin/bash
# round-2 GPU call V (1 GPU): ncu source-level capture of the TMA-fed K1 through the A/B harness (small shape)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:f32_topk -o gpurun_out/v_k1_tma -f ./scripts/ubench/k1_variants 1280 151552 1 0x400 > gpurun_out/v_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/v_ncu.log
