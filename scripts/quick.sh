# quick GPU check: parity subset + bench lines for all configs
mkdir -p gpurun_out/q
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_full_size or survey or config_corpora or straddling or specialised" > gpurun_out/q/pytest.log 2>&1; tail -2 gpurun_out/q/pytest.log
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/q/cfg2.json 2> gpurun_out/q/cfg2.err
for c in 3 4 5; do python bench.py --config $c --size 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/q/cfg$c.json 2> gpurun_out/q/cfg$c.err; done
python bench.py --config 1 --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/q/cfg1.json 2> gpurun_out/q/cfg1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/q/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms'],3), d['kernel_variant'])
    except Exception as e: print(f,'ERR',open(f).read()[-300:], open(f.replace('.json','.err')).read()[-800:])
PY
