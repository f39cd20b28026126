#!/bin/bash
# time the climatology sweep for several (keep, pool rows) settings
for kp in "$@"; do
  set -- $kp
  echo "KEEP=$1 POOL_ROWS=$2"
  XMHW_B200_KEEP=$1 XMHW_B200_POOL_ROWS=$2 timeout 120 python tools/quick_time.py c3q 2>&1 | grep -E "rep1|plan"
done
