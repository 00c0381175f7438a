#!/bin/bash
O=gpurun_out; T=${1:-r02al}
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_check_motion.py tests/test_gpu_plugin.py tests/test_gpu_rrtc_batch.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
timeout 300 python scripts/dropin_once.py 2>&1 | tail -1 | tee $O/${T}_dropin.log
PRGPU_NO_HOST_WINDOW=1 timeout 300 python scripts/dropin_once.py 2>&1 | tail -1 | tee -a $O/${T}_dropin.log
