#!/bin/bash
# round-2 GPU call O (1 GPU): compute-sanitizer over the new kernels (memcheck + racecheck), final ncu captures, all configs
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for args in "tc_cg1 40000 7 128 50" "tc_cg2 30000 300 256 100" "auto 3000 1 64 10" "f32 6000 40 96 20"; do
  timeout 600 $CS --tool memcheck --error-exitcode 9 python scripts/gpu_check.py $args > gpurun_out/o_memcheck_$(echo $args | tr ' ' '_').log 2>&1; echo "memcheck $args rc=$?"
done
timeout 900 $CS --tool racecheck --error-exitcode 9 python scripts/gpu_check.py tc_cg1 40000 7 128 50 > gpurun_out/o_racecheck_tc_cg1.log 2>&1; echo "racecheck tc_cg1 rc=$?"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fuse.py tests/test_gpu_inbatch.py -x -q > gpurun_out/o_memcheck_fuse_inbatch.log 2>&1; echo "memcheck fuse+inbatch rc=$?"
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fuse.py tests/test_gpu_inbatch.py -x -q > gpurun_out/o_racecheck_fuse_inbatch.log 2>&1; echo "racecheck fuse+inbatch rc=$?"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\| OK$\|passed\|failed" gpurun_out/o_*check*.log
