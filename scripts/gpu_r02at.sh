#!/bin/bash
PRGPU_NNF_TRACE=1 timeout 300 python scripts/nnf_once.py 10 2>&1 | grep "setup\|ok" | tee gpurun_out/r02at_setup.log
