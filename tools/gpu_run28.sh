# round 1, session 2, call 9 (2 GPUs): NCCL tests (grouped exchange incl. strided submit, sharded batch) + bench N=2: default and forced fused groups
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
for MODE in default fused; do
  EXTRA=""; [ $MODE = fused ] && EXTRA="--engine-opt fuse_mb=100000 --no-batched"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 400 --warmup 8 $EXTRA > gpurun_out/r01e_bench_n2_$MODE.json 2> gpurun_out/r01e_bench_n2_$MODE.err
  tail -1 gpurun_out/r01e_bench_n2_$MODE.err
  python -c "
import json; d=json.load(open('gpurun_out/r01e_bench_n2_$MODE.json'))
print('$MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'scan ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['top1'], d.get('batched'), d['clocks'])"
done
