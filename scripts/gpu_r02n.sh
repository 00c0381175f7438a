#!/bin/bash
# round-2 evidence run: smoke, all GPU tests, both bench arms, launch list of the bench, --set full of the dominant kernels
O=gpurun_out; T=${1:-r02n}
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -10 $O/${T}_pytest.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_knn.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $O/${T}_ncu_knn.log 2>&1; echo "launches knn rc=$?"
ncu --set full --clock-control none --import-source on -k regex:kd_search_team -s 3 -c 1 -o $O/${T}_full_kd python scripts/kdprof.py > $O/${T}_ncu_kd.log 2>&1; echo "full kd rc=$?"
