for cfg in "6 32" "6 48"; do set -- $cfg
for d in 0 4 16; do
HTN_STACK_TPJ=$1 HTN_STACK_FALL=0 HTN_WAVE_MB=$2 HTN_STACK_DEBUG=$d timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu --no-groundstate 2>gpurun_out/e.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tpj $1 waveMB $2 dbg $d', round(d['value']), {k:round(v,4) for k,v in d['stages_ms'].items()})"
done; done
