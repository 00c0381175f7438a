#!/bin/bash
O=gpurun_out; T=${1:-r02e}
python -m pytest tests/test_gpu_nnf.py tests/test_gpu_nnf_golden.py tests/test_gpu_plugin.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${T}_pytest.log
timeout 600 python bench.py --workload nnf --no-cpu-baseline > $O/${T}_bench_nnf.json 2> $O/${T}_bench_nnf.err; echo "bench nnf rc=$?"; tail -3 $O/${T}_bench_nnf.err
python -c "
import json; d=json.load(open('$O/${T}_bench_nnf.json')); print('nnf', d['value'], d['ms_per_step'], d['config'])"
for mb in 5 6 8; do
  rm -f pr-ompl_b200/csrc/check_motion.o; make -s -C pr-ompl_b200/csrc EXTRA="-DCM_MINB=$mb" > /dev/null 2>&1
  echo "== CM_MINB=$mb"; python bench.py --workload check_motion --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cm', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
rm -f pr-ompl_b200/csrc/check_motion.o; make -s -C pr-ompl_b200/csrc > /dev/null 2>&1
