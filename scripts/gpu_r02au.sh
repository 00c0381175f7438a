#!/bin/bash
O=gpurun_out; T=r02au
PRGPU_NNF_TRACE=1 timeout 300 python scripts/nnf_once.py 10 2>&1 | grep "setup\|ok" | tee $O/${T}_setup.log
timeout 900 python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py tests/test_gpu_plugin.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 300 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf', d['value'], c['setup_s'], c['solve_s'], d['e2e'])"
