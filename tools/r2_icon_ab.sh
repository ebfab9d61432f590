#!/bin/bash
# same-box A/B of ICON nabla2 planner knobs at 20 M edges
set -u
mkdir -p gpurun_out
for o in "" "levels_per_thread=65" "levels_per_thread=33" "levels_per_thread=65 unroll_k=4" "levels_per_thread=65 unroll_k=13" "levels_per_thread=65 max_blocks_per_sm=12" ""; do
  a=""; for kv in $o; do a="$a --opt $kv"; done
  timeout 300 python bench.py --workload icon_nabla2 --steps 10 --no-cpu-baseline --no-e2e $a > gpurun_out/icon_tmp.json 2> gpurun_out/icon_tmp.err
  python - "$o" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/icon_tmp.json").read().strip().splitlines()[-1])
    print("icon_nabla2 [%s] ms %.3f frac %.3f parity %s clocks %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["frac"], d["parity_checked"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("icon_nabla2 [%s] FAILED" % sys.argv[1], e); print(open("gpurun_out/icon_tmp.err").read()[-500:])
PY
done | tee gpurun_out/r2_icon_ab2.log
