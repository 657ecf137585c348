timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; tail -1 gpurun_out/r01_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e']['value'], 'scan', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])
for k,v in d['batched'].items(): print(k, round(v['queries_per_s']), round(v['ms_per_batch'],2), v['roofline']['achieved'], v['roofline']['frac'])
print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r01_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|filter_kernel" -c 30 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -f -o gpurun_out/r01_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -f -o gpurun_out/r01_tc_int8_full python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
