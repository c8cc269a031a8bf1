# GPU box: the new regression tests, then several fuzz processes side by side (half with the straddle variant forced)
# usage: bash scripts/fuzz_round.sh <outdir under gpurun_out> <minutes> [first seed]
out=gpurun_out/${1:-fz}; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "row_boundary_inside or picks_its_kernel" > $out/pytest_new.log 2>&1; tail -n 2 $out/pytest_new.log
min=${2:-3}; s0=${3:-1}
for i in 0 1 2; do
  timeout 400 python scripts/fuzz_gpu.py --seed $((s0 + 2 * i)) --minutes $min --jit-every $((3 + 4 * i)) > $out/fuzz$((s0 + 2 * i)).log 2>&1 &
  RRX_JIT_VARIANT=1 timeout 400 python scripts/fuzz_gpu.py --seed $((s0 + 2 * i + 1)) --minutes $min --jit-every $((3 + 3 * i)) > $out/fuzz$((s0 + 2 * i + 1)).log 2>&1 &
done
wait
tail -n 3 $out/fuzz*.log
