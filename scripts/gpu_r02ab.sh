#!/bin/bash
O=gpurun_out; T=${1:-r02ab}
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -x -q -k "kd" > $O/${T}_pytest_kd.log 2>&1; echo "pytest kd rc=$?"; tail -5 $O/${T}_pytest_kd.log
timeout 300 python scripts/kd_var.py 10000000 pipe 2>&1 | tail -5 | tee $O/${T}_kdvar.log
PRGPU_KD_PIPE=0 timeout 300 python scripts/kd_var.py 10000000 regs 2>&1 | tail -5 | tee -a $O/${T}_kdvar.log
timeout 600 python -m pytest tests/test_gpu_knn.py tests/test_gpu_full_size.py -m gpu -x -q > $O/${T}_pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 $O/${T}_pytest_knn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_search_pipe -s 3 -c 1 -o $O/${T}_full_kd python scripts/kdprof.py > $O/${T}_ncu_kd.log 2>&1; echo "ncu full kd rc=$?"
