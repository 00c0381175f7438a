#!/bin/bash
O=gpurun_out; T=${1:-r02k}
timeout 1200 python -m pytest tests/test_gpu_knn.py tests/test_gpu_full_size.py -m gpu -x -q --durations=5 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extra > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print('knn', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['gpu_launches'])
PY
timeout 600 python scripts/kdbench.py 2>&1 | grep -E "k=16|k=1:"
