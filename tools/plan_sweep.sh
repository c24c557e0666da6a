#!/bin/bash
# Shared-memory plan sweep of the hot kernels on one box: groups per stage (KBBQ_KPS), dinuc replicas (KBBQ_DREP),
# ring depth (KBBQ_MIN_STAGES .. KBBQ_MAX_STAGES).  bash tools/plan_sweep.sh R L N
LIB=kbbq-py_b200/kbbq/libkbbq_b200.so
R=$1; RL=$2; N=$3
echo "default plan: $(python tools/ab_kernels.py $LIB -- $R $RL $N | tail -1)"
while read k d smin smax; do
  echo "kps $k drep $d stages $smin..$smax: $(KBBQ_KPS=$k KBBQ_DREP=$d KBBQ_MIN_STAGES=$smin KBBQ_MAX_STAGES=$smax python tools/ab_kernels.py $LIB -- $R $RL $N | tail -1)"
done <<CFG
4 32 2 2
3 32 3 3
3 32 2 2
2 32 4 4
2 32 5 5
2 32 3 3
1 32 8 8
4 32 3 3
3 32 4 4
CFG
