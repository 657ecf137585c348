# round 1, session 2, call 5 (1 GPU): full parity + timings after carve-out fix / batch schedule defaults / batch TVFs; bench + profiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== single query 10M / 1.25M"
timeout 200 python tools/quick_bench.py --n 10000000 --iters 100 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --n 1250000 --iters 400 2>&1 | tail -1
echo "== bench"
timeout 600 python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/r01d_bench_n1.err; tail -2 gpurun_out/r01d_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01d_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e']['value'], 'scan', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])
for k,v in d['batched'].items(): print(k, v['queries_per_s'], v['ms_per_batch'], v['roofline']['achieved'], v['roofline']['frac'])
print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r01d_bench_reference.json
echo "== ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|filter_kernel" -c 30 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
tail -4 gpurun_out/r01d_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 -f -o gpurun_out/r01d_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 9 -c 1 -f -o gpurun_out/r01d_tc_int8_full python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r01d_batch_launches_int8_b1024.csv python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
