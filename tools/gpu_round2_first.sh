# Round 2, first GPU call (1 GPU, ~4 min): validate the two experimental tensor-core epilogue variants prepared at the
# end of round 1 (compiled, never run), and re-check the fused group launches on a full-size shard.
#   epi_max=1 : running-maximum pre-test (int8/uint8, L2/DOT)      tc_n=128 : 128-query tiles, 4 TMEM accumulator buffers
# quick_batch prints "match single-query path" for the LAST configuration of a sweep: run each variant last once.
mkdir -p gpurun_out
echo "== int8 L2: default vs epi_max (correctness of epi_max checked at the end)"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "epi_max=0;epi_max=1" 2>&1 | grep -v "iter 0" | tail -6
echo "== int8 L2: tc_n=128 (4 accumulator buffers)"
timeout 300 python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 3 --sweep "tc_n=0;tc_n=128" 2>&1 | grep -v "iter 0" | tail -6
echo "== uint8 DOT with epi_max, bf16 with tc_n=128"
timeout 300 python tools/quick_batch.py --n 4000000 --dim 384 --vtype 4 --metric 4 --nq 512 --iters 2 --sweep "epi_max=1" 2>&1 | tail -3
timeout 300 python tools/quick_batch.py --n 4000000 --dim 768 --vtype 3 --metric 4 --nq 1024 --iters 2 --sweep "tc_n=128" 2>&1 | tail -3
echo "== fused group launches on a full shard (was slower than single launches before the cta_time fix: re-measure)"
timeout 200 python tools/quick_group.py --n 10000000 --queries 160 2>&1 | tail -4
echo "== SQL end to end (SURVEY 8d: true SQL path at n <= 1M), ours vs the reference AVX2 build"
timeout 600 python tools/sql_bench.py --n 1000000 --dim 384 --queries 50 --which both 2>&1 | cut -c1-700
