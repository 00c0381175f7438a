#!/bin/bash
# usage (inside gpurun): scripts/gpu_round.sh TAG  -- tests, bench (both arms), ncu launch list + full capture
TAG=${1:-r02a}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_knn.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/${TAG}_ncu_knn.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:kd_search -c 2 -o $O/${TAG}_full_kd python scripts/kdprof.py > $O/${TAG}_ncu_kd.log 2>&1; echo "ncu full kd rc=$?"
