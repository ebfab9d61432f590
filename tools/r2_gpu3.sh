#!/bin/bash
set -u
mkdir -p gpurun_out
V="'' ring_levels=16,tma_stages=2 ring_levels=4,tma_stages=4 fuse_columns=0,column_ring=0 column_ring=0"
for shape in 64x148x80 64x444x80 64x888x80 64x148x40 64x148x160; do
  echo "== $shape"
  eval SWEEP_SHAPE=$shape timeout 600 python tools/sweep.py tridiagonal list $V 2>&1
done | tee gpurun_out/r2_sweep_tridiagonal3.log
