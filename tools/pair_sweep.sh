#!/bin/bash
# Fused (tb200_respair) vs two-launch timing of every residual-pair shape of the generators at config-2 sizes.
# usage: [SNAKES="1 0"] [MODES="fused unfused staged"] [CS="32 64 128"] [KDS="3,1 3,5 ..."] bash tools/pair_sweep.sh [dtype=f16]
dt=${1:-f16}
SNAKES=${SNAKES:-"1 0"}; MODES=${MODES:-"fused unfused staged"}; CS=${CS:-"32 64 128"}
KDS=${KDS:-"3,1 3,5 7,1 7,5 11,1 11,5"}
for sn in $SNAKES; do for C in $CS; do L=$((192000*32/C)); for kd in $KDS; do
  for mode in $MODES; do
    if [ $mode != unfused ]; then export TB200_PLAN_DEBUG=1; else unset TB200_PLAN_DEBUG; fi
    timeout 60 python tools/pair_micro.py $C ${kd/,/ } $L 64 $sn $dt 4 $mode 2>&1 | tail -2 | grep -v "^tb200 plan" | sed 's/tb200 respair plan: //; s/tb200 staged conv plan: //' | cut -c1-160
  done
done; done; done
