#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_heff.py tests/test_gpu_twosite.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
for cfg in "1 0" "1 64" "2 0"; do set -- $cfg
HTN_STACK_NG=$1 HTN_STACK_DEBUG=$2 timeout 120 python bench.py --steps 300 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NG $1 dbg $2', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
HTN_STACK_DEBUG=32 timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --no-groundstate 2>&1 >/dev/null | grep "stack timeline\|mix chunks" | sed -n 4,5p | cut -c1-400
