# round-end evidence: launch list of the default bench command + ncu --set full captures of every config's scan kernel
mkdir -p gpurun_out/final
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --sub none > gpurun_out/final/plain.json 2> gpurun_out/final/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --sub none > gpurun_out/final/ncu_l.log 2>&1
for c in 2 2v 3 4 5; do timeout 300 bash scripts/prof_cfg.sh $c > /dev/null 2>&1; done
ls gpurun_out/prof_c*/scan_full.ncu-rep
