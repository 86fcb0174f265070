timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_latest.log 2>&1; tail -c 200 gpurun_out/bench_latest.log
