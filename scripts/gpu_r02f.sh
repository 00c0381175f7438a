#!/bin/bash
O=gpurun_out; T=${1:-r02f}
python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py tests/test_gpu_plugin.py -m gpu -x -q --durations=25 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 $O/${T}_pytest.log
for mb in 10 12; do
  rm -f pr-ompl_b200/csrc/check_motion.o; make -s -C pr-ompl_b200/csrc EXTRA="-DCM_MINB=$mb" > /dev/null 2>&1
  echo "== CM_MINB=$mb"; python bench.py --workload check_motion --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cm', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
rm -f pr-ompl_b200/csrc/check_motion.o; make -s -C pr-ompl_b200/csrc > /dev/null 2>&1
