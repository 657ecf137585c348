timeout 200 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 2>&1 | tail -5
timeout 200 python tools/quick_batch.py --n 10000000 --dim 768 --vtype 3 --metric 4 --nq 1024 2>&1 | tail -5
timeout 200 python tools/quick_batch.py --n 6000000 --dim 1536 --vtype 4 --metric 3 --nq 256 --k 100 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 1 -f -o gpurun_out/r01_tc_full python tools/quick_batch.py --n 10000000 --dim 768 --vtype 3 --metric 4 --nq 1024 --iters 1 > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 4 -c 1 -f -o gpurun_out/r01_filter_full2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
