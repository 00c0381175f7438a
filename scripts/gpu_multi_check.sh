#!/bin/bash
# usage (inside gpurun --gpus N): scripts/gpu_multi_check.sh N TAG
N=${1:-2}; TAG=${2:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err; echo rc=$?
tail -3 gpurun_out/${TAG}_bench${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench${N}.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["extra"].get("query_sharded_c_abi"), d["extra"]["phase_breakdown"]); print(json.dumps(d["extra"].get("strong_scaling_one_batch"), indent=1)); print(json.dumps(d["extra"].get("configs4_lockstep_node_sharded_nn"), indent=1))
for k in ("check_motion","rrtc"):
    b=d["extra"][k]; print(k, b.get("value"), b.get("ms_per_step"), b.get("e2e",{}).get("value"), b.get("extra"))
PY
