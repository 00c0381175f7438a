#!/bin/bash
O=gpurun_out; T=r02aa
cp pr-ompl_b200/libprgpu.so /tmp/base.so
for v in base p3m4 p3m3 p4m4 p4m3 p2m5 p3m5; do
  if [ $v = base ]; then cp /tmp/base.so pr-ompl_b200/libprgpu.so; else cp variants/libprgpu_$v.so pr-ompl_b200/libprgpu.so; fi
  timeout 300 python scripts/kd_var.py 10000000 $v 2>&1 | tail -5
done | tee $O/${T}_kdvar.log
cp /tmp/base.so pr-ompl_b200/libprgpu.so
for t in 4 2; do PRGPU_KD_TEAM=$t timeout 300 python scripts/kd_var.py 10000000 team$t 2>&1 | tail -5; done | tee -a $O/${T}_kdvar.log
