#!/bin/bash
# ablation of the fused stage L+W kernel (HTN_STACK_DEBUG bits: 1 no A loads, 2 no T stores, 4 no mix work, 16 mix as its own launch)
for d in 0 1 2 4 3 6 7 16; do
HTN_STACK_DEBUG=$d timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg $d', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
HTN_STACK=0 timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('classic', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
HTN_PLAN_DEBUG=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-groundstate 2>&1 >/dev/null | grep "\[htn\]" | head -40
