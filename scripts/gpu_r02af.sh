#!/bin/bash
O=gpurun_out; T=${1:-r02af}
PRGPU_NNF_TRACE=1 timeout 300 python scripts/nnf_once.py 100000 > $O/${T}_nnf_trace.log 2>&1; echo "nnf rc=$?"; tail -3 $O/${T}_nnf_trace.log
timeout 300 python scripts/nnf_once.py 100000 2>&1 | tail -2
