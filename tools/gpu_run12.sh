timeout 300 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 2>&1 | tail -3
timeout 200 python tools/quick_batch.py --n 10000000 --dim 768 --vtype 3 --metric 4 --nq 1024 2>&1 | tail -2
timeout 200 python tools/quick_batch.py --n 6000000 --dim 1536 --vtype 4 --metric 3 --nq 256 --k 100 2>&1 | tail -2
timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 --opts batch_debug=1 2>&1 | grep -E "refine|finish|sort" | tail -9
