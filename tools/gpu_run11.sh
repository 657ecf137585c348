timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 --opts batch_debug=1 2>&1 | tail -22
timeout 200 python tools/quick_batch.py --n 6000000 --dim 1536 --vtype 4 --metric 3 --nq 256 --k 100 --iters 2 --opts batch_debug=1 2>&1 | tail -22
