# ncu --set full capture of the scan kernel of one config (1 GiB) + the run-time compiled cubin.
# usage: bash scripts/prof_cfg.sh <config> [size]   ->  gpurun_out/prof_c<config>/
c=$1; size=${2:-1073741824}; out=gpurun_out/prof_c$c
mkdir -p $out
python bench.py --config $c --size $size --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --sub none > $out/plain.json 2> $out/plain.err || { tail -5 $out/plain.err; exit 1; }
RRX_JIT_DUMP=$out/jit.cubin ncu --set full --clock-control none --import-source on -k regex:rrx_scan -s 4 -c 1 -o $out/scan_full -f python bench.py --config $c --size $size --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --sub none > $out/ncu_f.log 2>&1
ls -la $out | tail -5
