#!/bin/bash
O=gpurun_out; T=${1:-r02i}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 $O/${T}_pytest.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print('knn', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])
for k in ('check_motion','rrtc','nnf'):
    b=d['extra'][k]; print(k, b.get('value'), b.get('ms_per_step'), b.get('e2e',{}).get('value'))
print(d['extra']['nnf']['config']['ms_per_lazy_iteration'], d['extra']['nnf']['extra'])
PY
