#!/bin/bash
for h in 1.06 1.15; do PRGPU_NNF_TRACE=1 PRGPU_NNF_TAU_HEADROOM=$h timeout 300 python scripts/nnf_once.py 100000 > gpurun_out/r02ah_trace_$h.log 2>&1; tail -1 gpurun_out/r02ah_trace_$h.log; done
