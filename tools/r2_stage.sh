#!/bin/bash
# round-2 late session: ico_stage (bulk-copy staged rows) on hardware - parity tests, then same-box A/B at 20 M edges;
# solver ring shapes with 4 resident blocks per SM
set -u
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_unstructured_gpu.py -q -m gpu -x -rfE \
  -k "test_icon_laplacian or test_icon_nabla4 or test_icon_diamond_laplacian" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_stage_tests.log
timeout 200 python tools/sweep_icon.py 2582 list '' 'ico_hoist=1' 'ico_hoist=1,max_blocks_per_sm=12' 'ico_stage=1' 'ico_stage=2' \
  'ico_hoist=1,ico_stage=1' 'ico_hoist=1,ico_stage=2' 'ico_stage=2,levels_per_thread=4' 'ico_hoist=1,ico_stage=2,levels_per_thread=4' \
  'ico_stage=2,levels_per_thread=5,max_blocks_per_sm=8' 'ico_stage=1,levels_per_thread=4' '' 2>&1 | grep -v Warn | tee gpurun_out/r2_stage_icon.log
for rep in 1 2; do
for s in 512x512x80 1024x1024x80; do
  timeout 120 python tools/time_modules.py tridiagonal $s dawn_b200/lib/stencils/tridiagonal_solve.so dawn_b200/lib/stencils/tridiagonal_solve_ril12_tms2.so \
    dawn_b200/lib/stencils/tridiagonal_solve_ril11_tms2.so dawn_b200/lib/stencils/tridiagonal_solve_ril10_tms2.so 2>&1 | grep -v Warn
done; done | tee gpurun_out/r2_stage_tridiagonal.log
