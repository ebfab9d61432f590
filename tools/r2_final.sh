#!/bin/bash
# round-2 final single-GPU session: smoke, full GPU test suite, default bench line, launch lists, ICON ncu
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu -x -rfE --durations=6 2>&1 | tail -14 | cut -c1-250 | tee gpurun_out/r2_gpu_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default_20.json 2> gpurun_out/r2_bench_default_20.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_20.json").read().strip().splitlines()[-1])
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"])
show("hori_diff", d); print("  e2e api:", d["e2e"]["api"][:90], "| python mirror:", d.get("e2e_python_mirror", {}).get("value"), d.get("e2e_cpp_class_error"))
for k, r in d.get("workloads", {}).items(): show(k, r)
PY
B="python bench.py --no-cpu-baseline --no-sub-workloads --steps 20 --warmup 3"
$B > gpurun_out/plain_hd.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 60 --csv \
    --log-file gpurun_out/r2_hori_diff_launches.csv $B > gpurun_out/ncu_l1.log 2>&1
T="python bench.py --workload tridiagonal --no-cpu-baseline --steps 20 --warmup 3"
$T > gpurun_out/plain_tri.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 60 --csv \
    --log-file gpurun_out/r2_tridiagonal_launches.csv $T > gpurun_out/ncu_l2.log 2>&1
C="python bench.py --workload icon_nabla2 --no-cpu-baseline --no-e2e --steps 3 --warmup 3"
$C > gpurun_out/plain_icon.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 20 --csv \
    --log-file gpurun_out/r2_icon_nabla2_launches.csv $C > gpurun_out/ncu_l3.log 2>&1
$C > gpurun_out/plain_icon.log 2>&1 && ncu --set full --clock-control none -k regex:ICON_laplacian -s 6 -c 2 -f \
    -o gpurun_out/r2_prof_icon $C > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
