#!/bin/bash
# round-2 last call: the tempField fixture on hardware, then the driver's default bench line with the sampler that
# keeps the load running only until nvidia-smi has printed a sample
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_unstructured_gpu.py -q -m gpu -x -rfE -k "temp_field or atlas_suite" 2>&1 | tail -4 | cut -c1-250 | tee gpurun_out/r2d_tempfield_tests.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_default_20.json 2> gpurun_out/r2d_bench_default_20.err
python - <<'PY'
import json
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"])
try:
    d = json.loads(open("gpurun_out/r2d_bench_default_20.json").read().strip().splitlines()[-1])
    show("hori_diff", d)
    for k, r in d.get("workloads", {}).items(): show(k, r)
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2d_bench_default_20.err").read()[-800:])
PY
