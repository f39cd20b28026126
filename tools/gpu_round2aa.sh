#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02aa_tests.log
tail -3 gpurun_out/r02aa_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02aa_tests.log; then exit 0; fi
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr A=auto
run global025_quarter A=auto
run regional_40yr A=auto
} | tee gpurun_out/r02aa_kms.log
python - <<'PY'
# cost of the transpose BASELINE north_star item (1) asks for: time-major -> cell-major copy of the config-3 series
import torch, time
T, n = 10957, 1036800
x = torch.empty((T, n), dtype=torch.float32, device="cuda").normal_()
y = torch.empty((n, T), dtype=torch.float32, device="cuda")
for _ in range(2): y.copy_(x.t())
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): y.copy_(x.t())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("transpose %d x %d f32 (library copy kernel): %.2f ms = %.0f GB/s read+write" % (T, n, ms, 2 * T * n * 4 / ms / 1e6))
PY
