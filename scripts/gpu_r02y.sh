#!/bin/bash
O=gpurun_out; T=r02y
timeout 600 python -m pytest tests/test_gpu_check_motion.py -m gpu -x -q > $O/${T}_pytest_cm.log 2>&1; echo "pytest cm rc=$?"; tail -3 $O/${T}_pytest_cm.log
cp pr-ompl_b200/libprgpu.so /tmp/base.so
for v in base w16m2 w32m1 w4m8; do
  if [ $v = base ]; then cp /tmp/base.so pr-ompl_b200/libprgpu.so; else cp variants/libprgpu_$v.so pr-ompl_b200/libprgpu.so; fi
  timeout 300 python scripts/kd_var.py 10000000 $v 2>&1 | tail -5
done | tee $O/${T}_kdvar.log
cp /tmp/base.so pr-ompl_b200/libprgpu.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_search_team -s 3 -c 1 -o $O/${T}_full_kd python scripts/kdprof.py > $O/${T}_ncu_kd.log 2>&1; echo "ncu full kd rc=$?"
