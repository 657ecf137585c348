# round 1, session 2, call 6 (8 GPUs): bench at N=8 and N=4 (grouped exchange, two-stream engine, sharded batch extras)
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 800 --warmup 8 > gpurun_out/r01d_bench_n$N.json 2> gpurun_out/r01d_bench_n$N.err
  tail -2 gpurun_out/r01d_bench_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/r01d_bench_n$N.json'))
print('N=$N', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'scan ms', d['roofline']['avg_launch_ms'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['top1'], d.get('batched'), d['clocks'])"
done
