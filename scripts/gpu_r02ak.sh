#!/bin/bash
for b in 2048 1024 512 256 128 64; do echo "budget $b"; PRGPU_NNF_WAVE_BUDGET=$b timeout 300 python scripts/nnf_once.py 100000 2>&1 | tail -1; done | tee gpurun_out/r02ak_budget.log
PRGPU_NNF_TRACE=1 PRGPU_NNF_WAVE_BUDGET=256 timeout 300 python scripts/nnf_once.py 100000 > gpurun_out/r02ak_trace_256.log 2>&1
