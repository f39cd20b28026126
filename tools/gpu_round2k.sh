#!/bin/bash
# round 2k: general sweep with persistent blocks + near/far scratch + L2 access-policy window; top-K warp pairs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_tests.log
tail -3 gpurun_out/r02k_tests.log
python - <<'PY'
import torch, ctypes
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persistingL2CacheMaxSize", getattr(p, "persisting_l2_cache_max_size", None), "accessPolicyMaxWindowSize", getattr(p, "access_policy_max_window_size", None))
PY
run() { echo "== $*"; env "$@" bash tools/kms.sh global025_30yr 2>&1; }
{
run A=default
run XMHW_B200_SWEEP_L2=0
run XMHW_B200_SWEEP_NEAR=48
run XMHW_B200_SWEEP_NEAR=48 XMHW_B200_SWEEP_L2=0
run XMHW_B200_SWEEP_NEAR=4
run XMHW_B200_SWEEP_NEAR=12
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_PAIR=1
run XMHW_B200_SWEEP=topk
} | tee gpurun_out/r02k_kms.log
ncu --set full --clock-control none --import-source on -k regex:clim_sweep_kernel -s 1 -c 1 -o gpurun_out/sweep_r02k_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep_r02k.log 2>&1
XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_pair -s 1 -c 1 -o gpurun_out/pair_r02k_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/pair_r02k.log 2>&1
ls -la gpurun_out | tail -5
