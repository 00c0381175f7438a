#!/bin/bash
O=gpurun_out; T=${1:-r02ac}
for p in 3 4 5 6; do PRGPU_KD_PPL=$p timeout 300 python scripts/kd_var.py 10000000 pipe_ppl$p 2>&1 | tail -5; done | tee $O/${T}_kdvar.log
for p in 5; do PRGPU_KD_PIPE=0 PRGPU_KD_PPL=$p timeout 300 python scripts/kd_var.py 10000000 regs_ppl$p 2>&1 | tail -5; done | tee -a $O/${T}_kdvar.log
for t in 7 6; do PRGPU_KD_TEAM=$t PRGPU_KD_PPL=3 timeout 300 python scripts/kd_var.py 10000000 pipe_ppl3_team$t 2>&1 | tail -5; done | tee -a $O/${T}_kdvar.log
