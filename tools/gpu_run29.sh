# round 1, session 2, call 10 (1 GPU): final parity + bench + profiles of the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r01e_bench_n1.json 2> gpurun_out/r01e_bench_n1.err; tail -2 gpurun_out/r01e_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01e_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e'], 'scan', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])
for k,v in d['batched'].items(): print(k, v['queries_per_s'], v['ms_per_batch'], v['roofline']['achieved'], v['roofline']['frac'])
print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01e_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r01e_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|filter_kernel" -c 40 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
tail -4 gpurun_out/r01e_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 30 -c 1 -f -o gpurun_out/r01e_scan_full python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
ls -la gpurun_out/r01e*.ncu-rep
