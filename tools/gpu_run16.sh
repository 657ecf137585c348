for sb in 6144 8192 12288 16384 24576; do
timeout 300 python bench.py --no-cpu-baseline --no-batched --steps 200 --engine-opt stage_bytes=$sb > gpurun_out/tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('stage_bytes=$sb', round(d['value'],1), 'scan', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
for rb in 98304 131072; do
timeout 300 python bench.py --no-cpu-baseline --no-batched --steps 200 --engine-opt stage_bytes=12288 --engine-opt ring_bytes=$rb > gpurun_out/tmp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('stage 12288 ring_bytes=$rb', round(d['value'],1), 'scan', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
