#!/bin/bash
O=gpurun_out; T=${1:-r02ad}
timeout 300 python scripts/kd_var.py 10000000 new 2>&1 | tail -5 | tee $O/${T}_kdvar.log
PRGPU_KD_TEAM=8 timeout 300 python scripts/kd_var.py 10000000 team8 2>&1 | tail -5 | tee -a $O/${T}_kdvar.log
timeout 300 python scripts/kd_var.py 1250000 new 2>&1 | tail -5 | tee -a $O/${T}_kdvar.log
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_full_size.py tests/test_gpu_nnf.py -m gpu -x -q > $O/${T}_pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 $O/${T}_pytest_knn.log
