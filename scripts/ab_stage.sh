# A/B of the tile staging variants of the run-time compiled kernels (bulk copies / row pitch / byte-table copies)
mkdir -p gpurun_out/ab
for c in 2 3 5; do
  for v in "0 132 32" "0 132 16" "0 144 16" "1 144 16"; do
    set -- $v
    RRX_JIT_BULK=$1 RRX_JIT_PITCH=$2 RRX_JIT_BT=$3 timeout 300 python bench.py --config $c --size 1073741824 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --sub none > gpurun_out/ab/c${c}_$1_$2_$3.json 2> gpurun_out/ab/c${c}_$1_$2_$3.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab/c${c}_$1_$2_$3.json").read().strip().splitlines()[-1]); print("config $c bulk=$1 pitch=$2 bt=$3:", round(d['value'],1), "GB/s kernel_ms", round(d['roofline']['kernel_ms'],4))
except Exception as e:
    print("config $c $v ERR", open("gpurun_out/ab/c${c}_$1_$2_$3.err").read()[-300:])
PY
  done
done
