#!/bin/bash
python bench.py --steps 200 --warmup 5 --no-cpu --no-groundstate 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['stages_ms'])"
python tools/dense_gemm_check.py 2>&1 | grep "n=2048\|n= 167"
