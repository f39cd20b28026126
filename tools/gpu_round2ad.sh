#!/bin/bash
mkdir -p gpurun_out
for lib in "" "$PWD/xmhw_b200/_xmhw_b200_sums3.so"; do
  echo "== lib=$lib"
  XMHW_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-api 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms/step %.2f' % d['ms_per_step'], d['clocks'])
print({k.replace('xmhw_',''):round(v,2) for k,v in d['kernel_ms'].items()})"
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
done
cuobjdump -res-usage xmhw_b200/_xmhw_b200.so 2>/dev/null | grep -A1 "exceed4" | grep REG
cuobjdump -res-usage xmhw_b200/_xmhw_b200_sums3.so 2>/dev/null | grep -A1 "exceed4" | grep REG
