# ncu captures of the config-2 scan kernel: launch list + one full-set capture of the scan kernel
mkdir -p gpurun_out/prof
python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof/plain.json 2> gpurun_out/prof/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/prof/launches.csv python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof/ncu_l.log 2>&1
RRX_JIT_DUMP=gpurun_out/prof/jit_cfg2.cubin ncu --set full --clock-control none --import-source on -k regex:rrx_scan -s 4 -c 1 -o gpurun_out/prof/scan_full -f python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof/ncu_f.log 2>&1
ls -la gpurun_out/prof
