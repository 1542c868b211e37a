#!/bin/bash
for cfg in "1 0" "2 0" "2 128" "2 256" "2 384" "1 128"; do set -- $cfg
HTN_STACK_NG=$1 HTN_STACK_DEBUG=$2 timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NG $1 dbg $2', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
