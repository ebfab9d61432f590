#!/bin/bash
# round-2 GPU runs: tests, then the default bench line (with its sub-records)
set -u
mkdir -p gpurun_out
case "${1:-tests}" in
tests)
  timeout 1500 python -m pytest tests -q -m gpu -x -rfE --durations=8 2>&1 | tail -40 | tee gpurun_out/r2_gpu_tests.log
  ;;
refcuda)
  timeout 300 python -m pytest tests/test_reference_cuda_texts_gpu.py -q -m gpu -rfE 2>&1 | grep -E "^E  |FAILED|passed|failed" | cut -c1-400 | tee gpurun_out/r2_refcuda.log
  ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
  tail -c 600 gpurun_out/r2_bench_default.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"])
show("hori_diff", d)
for k, r in d.get("workloads", {}).items(): show(k, r)
PY
  ;;
esac
