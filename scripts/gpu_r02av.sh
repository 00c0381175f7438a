#!/bin/bash
PRGPU_NNF_TRACE=1 timeout 300 python scripts/nnf_setup_twice.py solve 2>&1 | grep "setup\|create\|close" | tee gpurun_out/r02av_setup.log
