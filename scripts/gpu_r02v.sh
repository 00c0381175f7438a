#!/bin/bash
for c in 32 64 128 256; do
  rm -f pr-ompl_b200/csrc/rrtc_batch.o; make -s -C pr-ompl_b200/csrc EXTRA="-DRB_CUT=$c" > /dev/null 2>&1
  echo "== RB_CUT=$c"; python bench.py --workload rrtc --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('rrtc', d['value'], d['ms_per_step'], d['config']['solved_fraction'])"
done
rm -f pr-ompl_b200/csrc/rrtc_batch.o; make -s -C pr-ompl_b200/csrc > /dev/null 2>&1
