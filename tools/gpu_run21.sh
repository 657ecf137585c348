# round 1, session 2, call 2 (2 GPUs): NCCL tests of the grouped exchange + sharded batch path, bench at N=2 with group sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
for G in 1 4 8; do
  EXTRA="--no-batched"; [ $G = 4 ] && EXTRA=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus 2 --steps 400 --warmup 5 --group $G $EXTRA > gpurun_out/r01b_bench_n2_g$G.json 2> gpurun_out/r01b_bench_n2_g$G.err
  tail -2 gpurun_out/r01b_bench_n2_g$G.err
  python -c "
import json; d=json.load(open('gpurun_out/r01b_bench_n2_g$G.json'))
print('G=$G', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'scan ms', d['roofline']['avg_launch_ms'], 'filter', d['roofline']['filter_kernel_avg_ms'], d['top1'], d.get('batched'))"
done
