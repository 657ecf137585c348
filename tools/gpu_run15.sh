timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_n1g.json 2> gpurun_out/bench_n1g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1g.json')); print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e']['value'], 'scan', d['roofline']['avg_launch_ms'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])"
