# round 1, session 2, call 4 (1 GPU): two-stream scan/filter overlap + async batch levels (buckets): parity, timing, schedule sweep, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== single query 10M / 1.25M"
timeout 200 python tools/quick_bench.py --n 10000000 --iters 100 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --n 1250000 --iters 400 2>&1 | tail -1
echo "== batch schedule sweep 10M"
SW="batch_m0=1024,batch_growth=32;batch_m0=512,batch_growth=16;batch_m0=256,batch_growth=16;batch_m0=256,batch_growth=8;batch_m0=128,batch_growth=8;batch_m0=128,batch_growth=4"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "$SW" 2>&1 | grep -v "iter 0" | tail -20
echo "== batch schedule sweep 1.25M"
timeout 300 python tools/quick_batch.py --n 1250000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "$SW" 2>&1 | grep -v "iter 0" | tail -20
echo "== bf16 768"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 768 --vtype 3 --metric 4 --nq 1024 --iters 3 --sweep "batch_m0=1024,batch_growth=32;batch_m0=256,batch_growth=8" 2>&1 | grep -v "iter 0" | tail -6
echo "== bench"
timeout 600 python bench.py > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; tail -2 gpurun_out/r01c_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01c_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','single_query_latency_ms')}, 'e2e', d['e2e']['value'], 'scan', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['clocks'])
for k,v in d['batched'].items(): print(k, v)
print(d['cpu_baseline'])"
