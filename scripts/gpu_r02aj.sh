#!/bin/bash
O=gpurun_out; T=${1:-r02aj}
for i in 1 2 3 4 5; do PRGPU_NNF_TRACE=1 timeout 300 python scripts/nnf_once.py 100000 > $O/${T}_trace_$i.log 2>&1; tail -1 $O/${T}_trace_$i.log; grep "\[nnf\]" $O/${T}_trace_$i.log | grep -v lazy | awk '$3>20'; done
for i in 1 2 3; do timeout 300 python scripts/nnf_once.py 100000 2>&1 | tail -1; done
