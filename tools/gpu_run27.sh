# round 1, session 2, call 8 (1 GPU): multi-query scan launches: parity + throughput; batch: norm prefetch, B-stationary with 8-warp epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== single query 10M / 1.25M"
timeout 200 python tools/quick_bench.py --n 10000000 --iters 100 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --n 1250000 --iters 400 2>&1 | tail -1
echo "== groups 1.25M / 10M"
timeout 200 python tools/quick_group.py --n 1250000 2>&1 | tail -4
timeout 200 python tools/quick_group.py --n 10000000 --queries 160 2>&1 | tail -4
echo "== batch 10M int8: default / bstat"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "bstat=0;bstat=1" 2>&1 | grep -v "iter 0" | tail -6
