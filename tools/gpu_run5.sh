timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1d.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['filter_kernel_avg_ms'], d['clocks'])"
