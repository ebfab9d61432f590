#!/bin/bash
# round-2 single-GPU evidence: default bench line (with sub-records), the per-GPU shape of C5 on one GPU,
# launch lists and ncu --set full captures of the two Cartesian headline kernels
set -u
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_default_200.json 2> gpurun_out/r2_bench_default_200.err
python bench.py --steps 20 --warmup 5 --shape 4096x512x80 --no-cpu-baseline > gpurun_out/r2_bench_hd_c5shape_1gpu.json 2> gpurun_out/r2_bench_hd_c5shape_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>&1
B="python bench.py --no-cpu-baseline --no-sub-workloads --steps 20 --warmup 3"
$B > gpurun_out/plain_hd.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2_hori_diff_launches.csv $B > gpurun_out/ncu_l1.log 2>&1
$B > gpurun_out/plain_hd.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hori_diff.*tma -s 5 -c 2 -f \
    -o gpurun_out/r2_prof_hd $B > gpurun_out/ncu_f1.log 2>&1
T="python bench.py --workload tridiagonal --no-cpu-baseline --steps 20 --warmup 3"
$T > gpurun_out/plain_tri.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2_tridiagonal_launches.csv $T > gpurun_out/ncu_l2.log 2>&1
$T > gpurun_out/plain_tri.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tridiagonal.*ring -s 5 -c 2 -f \
    -o gpurun_out/r2_prof_tri $T > gpurun_out/ncu_f2.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_200.json").read().strip().splitlines()[-1])
def show(k, r):
    if "error" in r: print(k, "ERROR", r["error"]); return
    print(k, "ms %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "parity", r.get("parity_checked"), r.get("max_abs_diff"),
          "e2e %.3g" % r.get("e2e", {}).get("value", float("nan")), r["clocks"])
show("hori_diff", d)
print(" ref cuda:", {k: v for k, v in d.get("reference_cuda_kernel", {}).items() if k != "what"})
for k, r in d.get("workloads", {}).items(): show(k, r)
c = json.loads(open("gpurun_out/r2_bench_hd_c5shape_1gpu.json").read().strip().splitlines()[-1]); show("c5shape_1gpu", c)
PY
tail -c 300 gpurun_out/r2_bench_reference_arm.json
