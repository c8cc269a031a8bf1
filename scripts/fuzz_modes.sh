# GPU box: fuzz of the one-shot calls' extra modes (anchors, non-overlapping view) and of streaming host ingestion
# usage: bash scripts/fuzz_modes.sh <outdir under gpurun_out> <minutes>
out=gpurun_out/${1:-fzm}; mkdir -p $out; min=${2:-2}
timeout 300 python scripts/fuzz_gpu.py --mode extra --seed 11 --minutes $min > $out/extra11.log 2>&1 &
timeout 300 python scripts/fuzz_gpu.py --mode extra --seed 12 --minutes $min > $out/extra12.log 2>&1 &
RRX_HOST_CHUNK=1048576 timeout 300 python scripts/fuzz_gpu.py --mode stream --seed 13 --minutes $min > $out/stream13.log 2>&1 &
RRX_HOST_CHUNK=1048576 timeout 300 python scripts/fuzz_gpu.py --mode stream --seed 14 --minutes $min > $out/stream14.log 2>&1 &
wait
tail -n 4 $out/*.log | cut -c1-300
