#!/bin/bash
O=gpurun_out; T=${1:-r02ai}
for i in 1 2 3; do timeout 300 python scripts/nnf_once.py 100000 2>&1 | tail -1; done | tee $O/${T}_nnf.log
timeout 900 python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py -m gpu -x -q > $O/${T}_pytest_nnf.log 2>&1; echo "pytest nnf rc=$?"; tail -3 $O/${T}_pytest_nnf.log
