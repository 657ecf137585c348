timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; cat gpurun_out/bench_n1b.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|filter_kernel" -c 26 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.json 2>/dev/null; tail -8 gpurun_out/r01_launches.csv | cut -c1-200
