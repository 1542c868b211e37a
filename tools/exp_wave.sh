#!/bin/bash
for w in 4 8 16 24 48 1000; do
HTN_WAVE_MB=$w timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('waveMB $w', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done
