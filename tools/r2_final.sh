#!/bin/bash
# round-2 final single-GPU session: smoke, full GPU test suite, default bench line, float lines, launch lists, ncu
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu -x -rfE --durations=4 2>&1 | tail -10 | cut -c1-250 | tee gpurun_out/r2_gpu_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default_20.json 2> gpurun_out/r2_bench_default_20.err
python bench.py --steps 200 --warmup 5 --no-sub-workloads > gpurun_out/r2_bench_hori_diff_200.json 2> gpurun_out/r2_bench_hori_diff_200.err
for w in hori_diff tridiagonal; do
  python bench.py --workload $w --steps 50 --no-cpu-baseline --no-sub-workloads --opt single_precision=1 > gpurun_out/r2_bench_${w}_f32.json 2> gpurun_out/r2_bench_${w}_f32.err
done
python - <<'PY'
import json
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"]["sm_mhz"], r["clocks"]["reasons"])
d = json.loads(open("gpurun_out/r2_bench_default_20.json").read().strip().splitlines()[-1])
show("hori_diff", d)
for k, r in d.get("workloads", {}).items(): show(k, r)
show("hori_diff 200 steps", json.loads(open("gpurun_out/r2_bench_hori_diff_200.json").read().strip().splitlines()[-1]))
for w in ("hori_diff", "tridiagonal"):
    try: show(w + " f32", json.loads(open("gpurun_out/r2_bench_%s_f32.json" % w).read().strip().splitlines()[-1]))
    except Exception as e: print(w, "f32 FAILED", e, open("gpurun_out/r2_bench_%s_f32.err" % w).read()[-600:])
PY
B="python bench.py --no-cpu-baseline --no-sub-workloads --steps 20 --warmup 3"
$B > gpurun_out/plain_hd.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 60 --csv \
    --log-file gpurun_out/r2_hori_diff_launches.csv $B > gpurun_out/ncu_l1.log 2>&1
$B > gpurun_out/plain_hd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hori_diff.*tma -s 5 -c 2 -f \
    -o gpurun_out/r2_prof_hd $B > gpurun_out/ncu_f1.log 2>&1
T="python bench.py --workload tridiagonal --no-cpu-baseline --steps 20 --warmup 3"
$T > gpurun_out/plain_tri.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 60 --csv \
    --log-file gpurun_out/r2_tridiagonal_launches.csv $T > gpurun_out/ncu_l2.log 2>&1
$T > gpurun_out/plain_tri.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tridiagonal.*ring -s 5 -c 2 -f \
    -o gpurun_out/r2_prof_tri $T > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/r2_prof_tri.ncu-rep
