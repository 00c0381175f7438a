#!/bin/bash
for i in 1 2; do
echo "== NNF device loop ($i)"; timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf dev', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'], d['extra']['lazysp_loop']['sm_clocks_during_solve'], d['extra']['lazysp_loop']['phase_share_of_loop_kernel']['repair'])"
done
echo "== NNF host loop"; PRGPU_NNF_HOST_LOOP=1 timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf host', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'], d['extra']['lazysp_loop']['sm_clocks_during_solve'])"
