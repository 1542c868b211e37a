#!/bin/bash
python bench.py --steps 100 --warmup 5 --no-cpu --no-groundstate 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['stages_ms'])"
python -m pytest tests/test_gpu_heff.py -x -q 2>&1 | tail -2
