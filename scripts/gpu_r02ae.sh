#!/bin/bash
O=gpurun_out; T=${1:-r02ae}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nnf_lazysp -c 12 -o $O/${T}_full_nnf python scripts/nnf_once.py 200 > $O/${T}_ncu_nnf.log 2>&1; echo "ncu nnf rc=$?"
