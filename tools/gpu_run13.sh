timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"refine|final_sort|replay|all_pairs|RadixSort|tc_scan|row_norm|fill_kernel|max_norm" -c 80 --csv --log-file gpurun_out/r01_batch_launches.csv python tools/quick_batch.py --n 10000000 --dim 384 --vtype 5 --metric 1 --nq 1024 --iters 2 > gpurun_out/ncu_batch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r01_batch_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-40:]: print(r[4][:70].ljust(72), r[-1])
PY
