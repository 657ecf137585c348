timeout 300 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -25
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_sql_surface.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_n1e.json 2> gpurun_out/bench_n1e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1e.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['filter_kernel_avg_ms'], d['clocks'])"
