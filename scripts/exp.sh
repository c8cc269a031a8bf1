mkdir -p gpurun_out/s6; rm -f gpurun_out/s6/*
for ta in 0 5; do
  for c in 2; do
    RRX_TICKET_AHEAD=$ta python bench.py --config $c --size 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6/ta${ta}_cfg$c.json 2> gpurun_out/s6/ta${ta}_cfg$c.err
  done
done
RRX_TICKET_AHEAD=5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "survey or config_corpora or straddling or config2_full or sizes_around" 2>&1 | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s6/ta*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))
    except Exception as e: print(f,'ERR',open(f).read()[-300:], open(f.replace('.json','.err')).read()[-500:])
PY
