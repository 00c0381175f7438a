#!/bin/bash
O=gpurun_out; T=${1:-r02x}
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print('knn', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'), d['gpu_launches'], d['clocks'])
for k in ('check_motion','rrtc','nnf'):
    b=d['extra'][k]; print(k, b.get('value'), b.get('ms_per_step'), b.get('e2e',{}).get('value'), b['roofline'].get('traffic'))
print(d['extra']['nnf']['config']['ms_per_lazy_iteration'], d['extra']['dropin_a'])
r=json.load(open("$O/${T}_bench_ref.json")); print('ref', r['value'], r['config'])
PY
