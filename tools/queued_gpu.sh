#!/bin/bash
# GPU work queued at the end of round 1 (everything below was written after the round's GPU minutes
# were spent; CPU-side checks are green).  One `gpurun -- bash tools/queued_gpu.sh` on 1 GPU for part A,
# `gpurun --gpus 2 -- bash tools/queued_gpu.sh multi` for part B.  Results land in gpurun_out/.
set -u
mkdir -p gpurun_out
if [ "${1:-single}" = "single" ]; then
  # A1: first run of the pending parity tests (Atlas-suite stencils, nabla2 on renumbered meshes):
  #     run in-process here to see every case; once green, drop the child-process wrapper and the xfail mark
  B200_PENDING_INPROC=1 timeout 120 python -m pytest tests/test_zz_pending_gpu.py -q -m gpu -rfE 2>&1 | tail -30 | tee gpurun_out/pending_tests.log
  # A2: does the RCM numbering (28 % fewer gather sectors per warp, tests/test_reorder.py) show up as time?
  for r in none rcm morton; do
    timeout 200 python bench.py --workload icon_nabla2 --steps 20 --no-cpu-baseline --icon-reorder $r \
      > gpurun_out/bench_icon_nabla2_$r.json 2> gpurun_out/bench_icon_nabla2_$r.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_icon_nabla2_$r.json").read().strip().splitlines()[-1])
print("icon_nabla2 numbering=$r ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
PY
  done
  # A0: the default bench line now carries "reference_cuda_kernel" (the reference's own CUDA text recompiled
  #     for sm_100a, timed in a child process): hori_diff and tridiagonal
  for w in hori_diff tridiagonal; do
    timeout 200 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w ours ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 3), "| reference CUDA text:", d.get("reference_cuda_kernel"))
PY
  done
  # A3: ICON's diamond nabla2 + Smagorinsky (13 fused edge stages, one launch, 248 B per edge-level)
  timeout 200 python bench.py --workload icon_diamond --steps 20 > gpurun_out/bench_icon_diamond.json 2> gpurun_out/bench_icon_diamond.err
  tail -c 900 gpurun_out/bench_icon_diamond.json
else
  N=${2:-2}
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  # B1: partitioned (general) unstructured mesh over NCCL point-to-point
  timeout 120 $T --master-port 29631 tests/multigpu_partition_check.py 2>&1 | grep -E "MULTIGPU|rror" | tee gpurun_out/partition_check.log
  # B2: peer-memory halo exchange against NCCL at N GPUs (interior slabs have two neighbours from N=3 on)
  for h in nccl p2p; do
    timeout 120 $T --master-port 29632 bench.py --gpus $N --halo $h --no-cpu-baseline \
      > gpurun_out/bench_hd_${N}gpu_$h.json 2> gpurun_out/bench_hd_${N}gpu_$h.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_hd_${N}gpu_$h.json").read().strip().splitlines()[-1])
print("hori_diff N=$N halo=$h ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], d["clocks"])
PY
  done
  MULTIGPU_QUICK=1 timeout 120 $T --master-port 29633 tests/multigpu_check.py 2>&1 | grep -E "MULTIGPU|EXCHANGE" | tee gpurun_out/mg_quick.log
fi
