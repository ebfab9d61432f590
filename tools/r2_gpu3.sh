#!/bin/bash
set -u
mkdir -p gpurun_out
V="ring_producer=0 ring_levels=7,tma_stages=4,block_size_i=32,ring_producer=0 ring_levels=6,tma_stages=4,block_size_i=32,ring_producer=0 ring_levels=8,tma_stages=3,block_size_i=32,ring_producer=0 ring_levels=6,tma_stages=5,block_size_i=32,ring_producer=0 ring_levels=5,tma_stages=5,block_size_i=32,ring_producer=0 ring_levels=7,tma_stages=4,block_size_i=32 ring_levels=14,tma_stages=2,block_size_i=32,ring_producer=0 ring_levels=12,tma_stages=2,block_size_i=32,ring_producer=0 ring_levels=10,tma_stages=3,block_size_i=32,ring_producer=0"
for shape in 512x512x80 1024x1024x80; do
  echo "== $shape"
  eval SWEEP_SHAPE=$shape timeout 600 python tools/sweep.py tridiagonal list $V 2>&1
done | tee gpurun_out/r2_sweep_tridiagonal5.log
