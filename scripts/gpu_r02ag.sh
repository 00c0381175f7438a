#!/bin/bash
for h in 1.03 1.06 1.1 1.15 1.25; do echo "headroom $h"; PRGPU_NNF_TAU_HEADROOM=$h timeout 300 python scripts/nnf_once.py 100000 2>&1 | tail -1; done | tee gpurun_out/r02ag_headroom.log
