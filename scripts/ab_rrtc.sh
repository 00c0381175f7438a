#!/bin/bash
# A/B of planner-kernel build variants on one box (development aid): ab_rrtc.sh "<flags>" ...
for flags in "$@"; do
  (cd /root/repo/pr-ompl_b200/csrc && touch rrtc_batch.cu && make -s NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags" >/dev/null 2>&1 || echo "build failed: $flags")
  echo "== $flags"
  (timeout 120 python bench.py --workload rrtc --no-cpu-baseline 2>/dev/null | cut -c1-130)
done
