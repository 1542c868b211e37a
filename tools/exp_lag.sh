#!/bin/bash
for d in 32 33 34 40 41; do
HTN_MIX_LAG=0 HTN_STACK_DEBUG=$d timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg $d', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
grep "stack timeline" gpurun_out/e.log | tail -1
done
