#!/bin/bash
# round-2 multi-GPU runs: bash tools/r2_mgpu.sh N   (under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== partition check"; timeout 300 $T --master-port 29631 tests/multigpu_partition_check.py 2>&1 | grep -E "MULTIGPU|rror|Traceback" | tee gpurun_out/r2_partition_check_${N}gpu.log
echo "== slab check"; MULTIGPU_QUICK=1 timeout 300 $T --master-port 29633 tests/multigpu_check.py 2>&1 | grep -E "MULTIGPU|EXCHANGE|rror" | tee gpurun_out/r2_mg_check_${N}gpu.log
run() { # name, args...
  local name=$1; shift
  timeout 600 $T --master-port 29632 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/r2_bench_${name}_${N}gpu.json 2> gpurun_out/r2_bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${name}_${N}gpu.json").read().strip().splitlines()[-1])
    print("${name} N=$N ms/step %.4f value %.4g frac %.3f parity %s e2e %.3g clocks %s" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity_checked"), d.get("e2e", {}).get("value", float("nan")), d["clocks"]))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2_bench_${name}_${N}gpu.err").read()[-1500:])
PY
}
run hd_default --steps 50
J=$((4096 / N))
run hd_c5 --steps 20 --shape 4096x${J}x80
IN=${2:-1024}   # TriMesh nx = ny per GPU (the global mesh is built, renumbered and cut on every rank)
run icon_slabs --workload icon_nabla2 --steps 10 --icon-n $IN
run icon_partition --workload icon_nabla2 --steps 10 --icon-n $IN --icon-partition rcm
