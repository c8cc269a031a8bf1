# GPU box, round end: the whole -m gpu suite (with one fuzz process of the extra modes beside it), then the default bench line
out=gpurun_out/${1:-fin}; mkdir -p $out
timeout 200 python scripts/fuzz_gpu.py --mode extra --seed 12 --minutes 2 > $out/extra12.log 2>&1 &
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -n 3 $out/pytest_gpu.log
wait
tail -n 2 $out/extra12.log | cut -c1-300
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json
