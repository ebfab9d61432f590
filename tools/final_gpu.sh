#!/bin/bash
# One GPU session: smoke, the three 1-GPU bench lines, the reference arm, and the ncu evidence for the
# hori_diff kernel (launch list + --set full).  Results land in gpurun_out/.
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in hori_diff tridiagonal icon_nabla2; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:hori_diff.*tma -s 5 -c 2 -f \
    -o gpurun_out/prof_hd_final python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/ncu_f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stencil -c 40 --csv \
    --log-file gpurun_out/hd_launches.csv python bench.py --no-cpu-baseline --steps 20 > gpurun_out/n1.log 2>&1
python - <<PY
import json
for w in ("hori_diff", "tridiagonal", "icon_nabla2"):
    d = json.loads(open("gpurun_out/bench_%s.json" % w).read().strip().splitlines()[-1])
    print(w, "ms", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "frac", round(d["roofline"]["frac"], 3),
          "e2e", d.get("e2e", {}).get("value"), d["clocks"])
print(open("gpurun_out/bench_ref.json").read()[-300:])
PY
