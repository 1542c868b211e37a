#!/bin/bash
NSITES=1 timeout 300 python tools/real_c4.py 1024 6 2>&1 | tail -2
python bench.py --steps 300 --warmup 5 --no-cpu --no-groundstate 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4-synthetic', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
timeout 300 python tools/run_config.py C2 1024 1e-11 2>&1 | grep -E "IDMRG2|VUMPS:|cumulative" 
timeout 200 python -m pytest tests/test_gpu_heff.py tests/test_gpu_mps.py -x -q 2>&1 | tail -2
