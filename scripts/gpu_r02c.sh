#!/bin/bash
# K2 pair-dealt pipeline: parity tests, timing, ncu; bounds test; kd_search block-size sweep
O=gpurun_out; T=r02c
python -m pytest tests/test_gpu_check_motion.py tests/test_gpu_bounds.py "tests/test_gpu_full_size.py::test_check_motion_full_size" -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
python bench.py --workload check_motion --no-cpu-baseline > $O/${T}_bench_cm.json 2> $O/${T}_bench_cm.err; echo "bench cm rc=$?"
python -c "
import json; d=json.load(open('$O/${T}_bench_cm.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_cm.csv python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_launches.py $O/${T}_launches_cm.csv $O/${T}_launches_cm_agg.csv | head -20
ncu --set full --clock-control none --import-source on -k regex:'cm_sample2|cm_lists|cm_certify_b' -c 8 -o $O/${T}_full_cm python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm_full.log 2>&1; echo "ncu full rc=$?"
for W in 8 4 2 1; do
  rm -f pr-ompl_b200/csrc/kdtree.o; make -s -C pr-ompl_b200/csrc EXTRA="-DKD_WARPS_N=$W" > /dev/null 2>&1
  echo "== KD_WARPS=$W"; python scripts/kdbench.py 2>&1 | grep -E "N=10000000 Q=(4096|32768) k=16|N=1250000 Q=4096 k=16"
done
rm -f pr-ompl_b200/csrc/kdtree.o; make -s -C pr-ompl_b200/csrc > /dev/null 2>&1
