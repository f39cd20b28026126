#!/bin/bash
# top-K sweep with unit slots in tensor memory: parity first (short timeouts: a TMEM mistake can hang), then timing
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweeps_forced" > gpurun_out/r02p_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_tests.log
tail -15 gpurun_out/r02p_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02p_tests.log; then exit 0; fi
{
echo "== quarter topk"; XMHW_B200_SWEEP=topk timeout 120 bash tools/kms.sh global025_quarter
echo "== quarter topk tmem"; XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 120 bash tools/kms.sh global025_quarter
echo "== full topk tmem"; XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 200 bash tools/kms.sh global025_30yr
echo "== full general"; timeout 200 bash tools/kms.sh global025_30yr
echo "== quarter_w2 topk tmem"; XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 120 bash tools/kms.sh quarter_w2
} 2>&1 | tee gpurun_out/r02p_kms.log
XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_tm -s 1 -c 1 -o gpurun_out/sweep2tm_r02p_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2tm_r02p.log 2>&1
