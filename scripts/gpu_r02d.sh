#!/bin/bash
O=gpurun_out; T=${1:-r02d}
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
python bench.py --workload check_motion --no-cpu-baseline > $O/${T}_bench_cm.json 2> $O/${T}_bench_cm.err; echo "bench cm rc=$?"
python -c "
import json; d=json.load(open('$O/${T}_bench_cm.json')); print('cm', d['value'], d['ms_per_step'], d['e2e']['value'])"
python bench.py --workload rrtc --no-cpu-baseline > $O/${T}_bench_rrtc.json 2> $O/${T}_bench_rrtc.err; echo "bench rrtc rc=$?"
python -c "
import json; d=json.load(open('$O/${T}_bench_rrtc.json')); print('rrtc', d['value'], d['ms_per_step'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_cm.csv python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_launches.py $O/${T}_launches_cm.csv $O/${T}_launches_cm_agg.csv | head -12
ncu --set full --clock-control none --import-source on -k regex:'cm_sample2|cm_lists|cm_certify_b' -c 4 -o $O/${T}_full_cm python scripts/cm_once.py 1000000 > $O/${T}_ncu_cm_full.log 2>&1; echo "ncu full rc=$?"
