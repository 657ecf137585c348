"""print the handful of numbers of a bench.py JSON line that a GPU session looks at first"""
import json
import sys

d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")})
print("e2e", d.get("e2e"))
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "avg_launch_ms", "traffic", "kernel_alone_ms", "kernel_alone_gbs", "tc_kernel_ms_per_batch", "hbm")} if d.get("roofline") else None)
print("clocks", d.get("clocks"))
for k, v in (d.get("batched") or {}).items():
    if isinstance(v, dict):
        print("batched", k, {kk: v.get(kk) for kk in ("queries_per_s", "ms_per_batch", "ms_per_global_batch", "queries_per_s_results_on_every_rank", "parity", "error") if v.get(kk) is not None}, (v.get("roofline") or {}).get("achieved"), (v.get("roofline") or {}).get("frac"))
for k in ("fp_single_query", "sql_e2e", "cpu_baseline", "parity"):
    if d.get(k) is not None:
        print(k, d[k])
