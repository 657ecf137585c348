timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 2>&1 | tail -2
timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --opts bstat=1 2>&1 | tail -2
timeout 200 python tools/quick_batch.py --n 10000000 --dim 128 --vtype 5 --metric 4 --nq 1024 --iters 3 2>&1 | tail -2
timeout 200 python tools/quick_batch.py --n 10000000 --dim 128 --vtype 5 --metric 4 --nq 1024 --iters 3 --opts bstat=1 2>&1 | tail -2
