# round 1, session 2, call 7 (1 GPU): adaptive row partition of the single-query scan: parity + effect; tc ncu capture; C4 shard timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== balance 10M"; timeout 200 python tools/quick_balance.py --n 10000000 2>&1 | tail -7
echo "== balance 1.25M"; timeout 200 python tools/quick_balance.py --n 1250000 --iters 400 2>&1 | tail -7
echo "== C4 shard: uint8 cosine dim1536 6.25M rows k=100 B=256"
timeout 300 python tools/quick_batch.py --n 6250000 --dim 1536 --vtype 4 --metric 3 --k 100 --nq 256 --iters 3 2>&1 | tail -4
echo "== tc ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 5 -c 1 -f -o gpurun_out/r01d_tc_int8_full python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
