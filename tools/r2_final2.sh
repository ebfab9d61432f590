#!/bin/bash
# round-2 closing single-GPU session: smoke, the whole GPU suite, the driver's default bench line, ncu of the
# staged ICON variant beside the default kernels (3.15 M edges x 65 levels)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu -x -rfE --durations=3 2>&1 | tail -9 | cut -c1-250 | tee gpurun_out/r2c_gpu_tests_final.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_default_20.json 2> gpurun_out/r2c_bench_default_20.err
python - <<'PY'
import json
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"])
try:
    d = json.loads(open("gpurun_out/r2c_bench_default_20.json").read().strip().splitlines()[-1])
    show("hori_diff", d)
    for k, r in d.get("workloads", {}).items(): show(k, r)
except Exception as e:
    print("bench FAILED", e, open("gpurun_out/r2c_bench_default_20.err").read()[-800:])
PY
S="python tools/sweep_icon.py 1024 list"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ICON -s 2 -c 2 -f -o gpurun_out/r2c_prof_icon_staged \
    $S 'ico_stage=1,levels_per_thread=4' > gpurun_out/ncu_s1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ICON -s 2 -c 2 -f -o gpurun_out/r2c_prof_icon_default \
    $S '' > gpurun_out/ncu_s2.log 2>&1
ls -la gpurun_out/r2c_prof_icon_*.ncu-rep
