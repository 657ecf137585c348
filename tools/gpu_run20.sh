# round 1, session 2, call 1 (1 GPU): parity of everything changed (slots, grouped exchange heads, sharded batch merge, 8-warp epilogue) + epilogue sweep + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "epi2=0;epi2=1;epi2=2" 2>&1 | tail -14
timeout 300 python tools/quick_batch.py --n 10000000 --dim 768 --vtype 3 --metric 4 --nq 1024 --iters 2 --sweep "epi2=0;epi2=1;epi2=2" 2>&1 | tail -10
timeout 600 python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; tail -2 gpurun_out/r01b_bench_n1.err; cut -c1-1500 gpurun_out/r01b_bench_n1.json
