#!/bin/bash
# ncu --set full of the fused stage L+W kernel (normal and skeleton-only), after a clean run of the same command
for d in 0 7; do
HTN_STACK_DEBUG=$d timeout 300 ncu --set full --clock-control none --import-source on -k regex:stack_gemm -s 6 -c 1 -f -o gpurun_out/r2_stack_dbg$d \
  python bench.py --steps 10 --warmup 3 --no-cpu --no-groundstate > gpurun_out/ncu_stack_$d.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
