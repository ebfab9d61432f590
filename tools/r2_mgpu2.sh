#!/bin/bash
# round-2 second multi-GPU session: whole-tile strips, NCCL against peer-memory halos, C5 as named
set -u
N=${1:-8}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {
  local name=$1; shift
  timeout 600 $T --master-port 29632 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/r2b_bench_${name}_${N}gpu.json 2> gpurun_out/r2b_bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench_${name}_${N}gpu.json").read().strip().splitlines()[-1])
    print("${name} N=$N ms/step %.4f value %.4g kernel-frac %.3f parity %s clocks %s" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("parity_checked"), d["clocks"]))
except Exception as e:
    print("${name} FAILED", e); print(open("gpurun_out/r2b_bench_${name}_${N}gpu.err").read()[-1200:])
PY
}
run hd_nccl --steps 100 --warmup 5
run hd_p2p --steps 100 --warmup 5 --halo p2p
J=$((4096 / N))
run c5_nccl --steps 50 --shape 4096x${J}x80
run c5_p2p --steps 50 --shape 4096x${J}x80 --halo p2p
