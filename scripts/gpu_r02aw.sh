#!/bin/bash
PRGPU_NNF_TRACE=1 timeout 300 python bench.py --workload nnf --no-cpu-baseline 2>&1 >/dev/null | grep "setup" | tee gpurun_out/r02aw_setup.log
