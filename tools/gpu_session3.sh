#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s3.log 2>&1; tail -5 gpurun_out/pytest_gpu_s3.log
for g in 0 1 1 0; do
  echo "== cfg5 LCGP_GRAPHS=$g"
  LCGP_GRAPHS=$g timeout 600 python bench.py --config cfg5_batch 2> gpurun_out/bench_cfg5_s3_g$g.err | tee gpurun_out/bench_cfg5_s3_g$g.json | cut -c1-200
done
for t in 2 4 16; do
  echo "== cfg5 graphs threads=$t"
  timeout 600 python bench.py --config cfg5_batch --threads $t 2>/dev/null | cut -c1-200
done
