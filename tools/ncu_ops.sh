#!/bin/bash
# ncu --set full captures of the non-headline kernels; each op first runs clean, then under ncu.
set -x
for op in ac2 transfer krylov qr svd; do
  timeout 300 python tools/profile_ops.py $op > gpurun_out/r2_ops_$op.txt 2>&1 || { tail -5 gpurun_out/r2_ops_$op.txt; continue; }
  cat gpurun_out/r2_ops_$op.txt
  case $op in
    svd) extra="-k regex:svd_round -s 300 -c 2";;
    krylov) extra="-k regex:multi -s 8 -c 3";;
    qr) extra="-k regex:qr_ -s 1 -c 1";;
    *) extra="-s 12 -c 8";;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on $extra -f -o gpurun_out/r2_ops_$op python tools/profile_ops.py $op > gpurun_out/r2_ops_${op}_ncu.log 2>&1
done
ls -la gpurun_out/r2_ops_*.ncu-rep
