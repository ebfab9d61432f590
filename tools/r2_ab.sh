#!/bin/bash
# A/B of two solver ring shapes, alternating, three shapes
for rep in 1 2 3; do
for s in 512x512x80 1024x1024x80; do
  python tools/time_modules.py tridiagonal $s dawn_b200/lib/stencils/tridiagonal_solve.so dawn_b200/lib/stencils/tridiagonal_solve_ril16_tms2.so dawn_b200/lib/stencils/tridiagonal_solve_ril12_tms3.so 2>&1 | grep -v Warn
done; done
