#!/bin/bash
O=gpurun_out; T=${1:-r02o}
timeout 900 python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
echo "== NNF device loop"; timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf dev', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'], d['extra']['lazysp_loop'])"
echo "== NNF host loop"; PRGPU_NNF_HOST_LOOP=1 timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf host', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'])"
true
true
timeout 300 python scripts/kd_slices.py
