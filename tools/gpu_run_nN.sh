N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$N.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'scan ms', d['roofline']['avg_launch_ms'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['top1'])"
