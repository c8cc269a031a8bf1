# ncu --set full capture of config <c> with RRX_JIT_DEFS="<defs>" -> gpurun_out/prof_<tag>/
c=$1; defs=$2; tag=$3; out=gpurun_out/prof_$tag
mkdir -p $out
RRX_JIT_DEFS="$defs" RRX_JIT_DUMP=$out/jit.cubin ncu --set full --clock-control none --import-source on -k regex:rrx_scan -s 4 -c 1 -o $out/scan_full -f python bench.py --config $c --size 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --sub none > $out/ncu_f.log 2>&1
ls $out | tr '\n' ' '
