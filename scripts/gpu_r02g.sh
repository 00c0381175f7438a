#!/bin/bash
O=gpurun_out; T=${1:-r02g}
python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py -m gpu -x -q --durations=5 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
timeout 600 python bench.py --workload nnf --no-cpu-baseline > $O/${T}_bench_nnf.json 2> $O/${T}_bench_nnf.err; echo "bench nnf rc=$?"; tail -3 $O/${T}_bench_nnf.err
python -c "
import json; d=json.load(open('$O/${T}_bench_nnf.json')); print('nnf', d['value'], d['ms_per_step'], d['config'], d['extra'])"
