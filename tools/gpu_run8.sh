timeout 900 python bench.py > gpurun_out/bench_n1f.json 2> gpurun_out/bench_n1f.err; tail -2 gpurun_out/bench_n1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1f.json'))
print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e']['value'], 'scan ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])
print(json.dumps(d.get('batched'), indent=1)[:1800])
print(d.get('cpu_baseline'))"
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
