#!/bin/bash
# A/B of collect-kernel build variants on one box (development aid): ab_collect.sh "<flags>" ...
cd /root/repo/pr-ompl_b200/csrc
for flags in "$@"; do
  touch knn_filter.cu
  make -s NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags" >/dev/null 2>&1 || echo "build failed: $flags"
  echo "== $flags"
  (cd /root/repo && timeout 300 python scripts/knn_time.py 10000000 4096 16 10)
done
