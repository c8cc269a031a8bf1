mkdir -p gpurun_out/s6
for ta in 0 1 2; do
  for c in 2 4 5; do
    RRX_TICKET_AHEAD=$ta python bench.py --config $c --size 1073741824 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s6/ta${ta}_cfg$c.json 2> gpurun_out/s6/ta${ta}_cfg$c.err
  done
done
python bench.py --config 1 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/s6/cfg1.json 2>&1
python bench.py --config 3 --size 1073741824 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/s6/cfg3.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s6/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],3))
    except Exception as e: print(f,'ERR',open(f).read()[-300:])
PY
