#!/bin/bash
# occupancy scaling of the two sweeps: narrower windows = smaller pools = more resident warps
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r02o_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02o_tests.log
tail -3 gpurun_out/r02o_tests.log
{
for wl in quarter_w1 quarter_w2 quarter_w3 global025_quarter; do
  echo "== $wl general"; bash tools/kms.sh $wl
  echo "== $wl topk"; XMHW_B200_SWEEP=topk bash tools/kms.sh $wl
done
echo "== quarter_w2 topk wpb2"; XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_WPB=2 bash tools/kms.sh quarter_w2
echo "== quarter_w1 topk wpb2"; XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_WPB=2 bash tools/kms.sh quarter_w1
} 2>&1 | tee gpurun_out/r02o_kms.log
XMHW_B200_SWEEP=topk ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_kernel -s 1 -c 1 -o gpurun_out/sweep2_r02o_w2 \
    python bench.py --workload quarter_w2 --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2_r02o.log 2>&1
