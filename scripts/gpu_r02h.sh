#!/bin/bash
O=gpurun_out; T=${1:-r02h}
timeout 900 python -m pytest tests/test_gpu_knn.py "tests/test_gpu_full_size.py::test_knn_sweep_full_size" -m gpu -x -q --durations=5 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${T}_pytest.log
for t in 0 -1 4 2; do
  echo "== PRGPU_KD_TEAM=$t"; PRGPU_KD_TEAM=$t timeout 600 python scripts/kdbench.py 2>&1 | grep -E "k=16"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== NNF device loop"; timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf dev', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'], d['extra']['lazysp_loop']['wide_waves_handed_to_full_sweep'])"
echo "== NNF host loop"; PRGPU_NNF_HOST_LOOP=1 timeout 600 python bench.py --workload nnf --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('nnf host', d['value'], c['setup_s'], c['solve_s'], c['ms_per_lazy_iteration'])"
