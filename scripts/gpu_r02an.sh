#!/bin/bash
O=gpurun_out; T=${1:-r02an}
python scripts/latency_probe.py 2>&1 | tail -7 | tee $O/${T}_latency.log
timeout 300 python scripts/dropin_once.py 2>&1 | tail -1 | tee $O/${T}_dropin.log
timeout 1200 python -m pytest tests/test_gpu_check_motion.py tests/test_gpu_plugin.py tests/test_gpu_rrtc_batch.py tests/test_gpu_nnf.py tests/test_gpu_bounds.py tests/test_gpu_full_size.py -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'check_motion_kernel|is_valid_kernel' -c 20 --csv --log-file $O/${T}_launches.csv python scripts/latency_probe.py > /dev/null 2>&1
