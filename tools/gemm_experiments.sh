#!/bin/bash
for c in 1 2 3; do
HTN_GEMM_CTAS=$c python bench.py --steps 30 --warmup 5 --no-cpu --no-groundstate 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas/SM=$c bench', d['value'], d['stages_ms'])"
done
