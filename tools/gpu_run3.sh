timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1c.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['filter_kernel_avg_ms'], d['clocks'])"
for sb in 12288 16384 6144; do python tools/quick_bench.py --iters 50 --opts "stage_bytes=$sb" 2>&1 | tail -1; done
python tools/quick_bench.py --iters 30 --vtype 1 --dim 384 --n 4000000 2>&1 | tail -1
python tools/quick_bench.py --iters 30 --vtype 3 --dim 768 --n 4000000 --metric 4 2>&1 | tail -1
python tools/quick_bench.py --iters 30 --vtype 4 --dim 1536 --n 3000000 --metric 3 2>&1 | tail -1
python tools/quick_bench.py --iters 30 --vtype 2 --dim 128 --n 10000000 --metric 5 2>&1 | tail -1
