#!/bin/bash
# GPU suite, full bench line, launch list and ncu --set full of the two headline launches (each after a clean run)
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputest.log 2>&1; tail -3 gpurun_out/r2_gputest.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-groundstate > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 90 --csv --log-file gpurun_out/r2_heff_ac_launch_list.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu --no-groundstate > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stack_gemm|grouped_gemm' -s 12 -c 2 -f -o gpurun_out/r2_heff_ac_final \
  python bench.py --steps 20 --warmup 3 --no-cpu --no-groundstate > gpurun_out/ncu_final.log 2>&1
ls -la gpurun_out/r2_heff_ac_final.ncu-rep gpurun_out/r2_heff_ac_launch_list.csv
