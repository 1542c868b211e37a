#!/bin/bash
# A/B of library variants built into build/variants/ (see DESIGN.md section 8): swaps each over
# hubbardtn_b200/libhtn.so on the GPU box's scratch copy and runs the short bench.
for v in "$@"; do
  cp build/variants/libhtn_$v.so hubbardtn_b200/libhtn.so
  timeout 150 python bench.py --steps 500 --warmup 20 --no-cpu --no-groundstate 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VARIANT $v', round(d['value'],1), d.get('stages_ms'), round(d['roofline']['frac'],4))"
done
