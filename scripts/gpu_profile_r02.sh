#!/bin/bash
# end-of-round evidence: launch lists of the default bench + --set full captures of the dominant kernels
O=gpurun_out; T=${1:-r02m}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/${T}_bench_plain.json 2>/dev/null; echo "plain bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_knn.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/${T}_ncu_knn.log 2>&1; echo "launches knn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:kd_search_team -c 2 -o $O/${T}_full_kd python scripts/kdprof.py > $O/${T}_ncu_kd.log 2>&1; echo "full kd rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_cm.csv python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm.log 2>&1; echo "launches cm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'cm_sample2|cm_lists|cm_certify_b' -c 4 -o $O/${T}_full_cm python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm_full.log 2>&1; echo "full cm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'rrtc_' -c 2 -o $O/${T}_full_rrtc python scripts/rrtc_once.py 4096 > $O/${T}_ncu_rrtc.log 2>&1; echo "full rrtc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'nnf_flow_dp|nnf_lazysp' -c 3 -o $O/${T}_full_nnf python scripts/nnf_once.py 60 > $O/${T}_ncu_nnf.log 2>&1; echo "full nnf rc=$?"
