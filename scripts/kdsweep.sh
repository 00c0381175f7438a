#!/bin/bash
# rebuild kdtree.cu with different (slots per lane, min blocks per SM) and time the headline shape
for cfg in "2 4" "2 2" "4 2" "4 3" "1 4" "8 2"; do
  set -- $cfg
  rm -f pr-ompl_b200/csrc/kdtree.o
  make -s -C pr-ompl_b200/csrc EXTRA="-DKD_PPL=$1 -DKD_MINB=$2" > /dev/null 2>&1
  echo "== PPL=$1 MINB=$2"
  python scripts/kdbench.py 2>&1 | grep -E "N=10000000 Q=(4096|32768) k=16|N=1250000 Q=4096 k=16"
done
rm -f pr-ompl_b200/csrc/kdtree.o; make -s -C pr-ompl_b200/csrc > /dev/null 2>&1
