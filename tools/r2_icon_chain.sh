#!/bin/bash
# ICON nabla2 at 20 M edges: chained launch variants against the two-launch default
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unstructured_gpu.py -q -m gpu -x -k "icon_laplacian or nabla4 or overwriting or random" 2>&1 | tail -3
for o in "" "ico_chain=1" "ico_chain=1 ico_chain_lag=4096" "ico_chain=1 ico_cache_hints=1" "ico_chain=1 ico_chain_l1=0 ico_cache_hints=1" "ico_chain=1 max_blocks_per_sm=12"; do
  a=""; for kv in $o; do a="$a --opt $kv"; done
  timeout 300 python bench.py --workload icon_nabla2 --steps 10 --no-cpu-baseline --no-e2e $a > gpurun_out/icon_tmp.json 2> gpurun_out/icon_tmp.err
  python - "$o" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/icon_tmp.json").read().strip().splitlines()[-1])
    print("icon_nabla2 [%s] ms %.3f frac %.3f parity %s launches/step %d clocks %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["frac"], d["parity_checked"], d["gpu_launches"] // d["steps"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("icon_nabla2 [%s] FAILED" % sys.argv[1], e); print(open("gpurun_out/icon_tmp.err").read()[-800:])
PY
done | tee gpurun_out/r2_icon_chain4.log
