# A/B of compile-time variants of the run-time compiled kernel: bash scripts/ab_defs.sh <config> "<defs1>" "<defs2>" ...
c=$1; shift
mkdir -p gpurun_out/ab
i=0
for d in "$@"; do
  i=$((i+1))
  RRX_JIT_DEFS="$d" timeout 300 python bench.py --config $c --size 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --sub none > gpurun_out/ab/d${c}_$i.json 2> gpurun_out/ab/d${c}_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab/d${c}_$i.json").read().strip().splitlines()[-1]); print("config $c [$d]:", round(d['value'],1), "GB/s kernel_ms", round(d['roofline']['kernel_ms'],4))
except Exception as e:
    print("config $c [$d] ERR", open("gpurun_out/ab/d${c}_$i.err").read()[-300:])
PY
done
