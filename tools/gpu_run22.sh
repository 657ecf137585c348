# round 1, session 2, call 3 (1 GPU): CTA-list filter regression + timing, epilogue sweep (8/16 warps), batch stage timings, small-shard profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== single query 10M / 1.25M"
timeout 200 python tools/quick_bench.py --n 10000000 --iters 50 2>&1 | tail -1
timeout 200 python tools/quick_bench.py --n 1250000 --iters 200 2>&1 | tail -1
echo "== batch sweep 10M"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "epi2=2;epi2=4;epi2=5" 2>&1 | tail -12
echo "== batch stages 10M / 1.25M"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 --opts batch_debug=1 2>&1 | tail -32
timeout 300 python tools/quick_batch.py --n 1250000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 --opts batch_debug=1 2>&1 | tail -30
echo "== ncu small shard"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_kernel|filter_kernel" -s 10 -c 20 --csv --log-file gpurun_out/r01b_launches_1250k.csv python tools/quick_bench.py --n 1250000 --iters 20 > /dev/null 2>&1
tail -8 gpurun_out/r01b_launches_1250k.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel" -s 6 -c 1 -f -o gpurun_out/r01b_scan_1250k python tools/quick_bench.py --n 1250000 --iters 5 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"filter_kernel" -s 6 -c 1 -f -o gpurun_out/r01b_filter python tools/quick_bench.py --n 1250000 --iters 5 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
