mkdir -p gpurun_out/prof5
RRX_JIT_DUMP=gpurun_out/prof5/jit_cfg5.cubin ncu --set full --clock-control none --import-source on -k regex:rrx_scan -s 4 -c 1 -o gpurun_out/prof5/scan_full -f python bench.py --config 5 --size 1073741824 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof5/ncu_f.log 2>&1
ls -la gpurun_out/prof5
