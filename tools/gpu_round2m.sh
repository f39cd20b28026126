#!/bin/bash
# where does the public-API time go (config 3)?  cProfile of xmhw.threshold + xmhw.detect(compact=True)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu --api-profile gpurun_out/api_profile_r02m.txt > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02m.json').readline()); print(d['e2e']['api'])"
head -70 gpurun_out/api_profile_r02m.txt
python - <<'PY'
import time, torch, numpy as np
def t(f, *a):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(*a); torch.cuda.synchronize(); return time.perf_counter()-t0, r
for gb in (1, 6):
    n = gb << 30
    dt, x = t(lambda: torch.empty(n, dtype=torch.uint8, pin_memory=True)); print("pinned alloc %d GB: %.2f s" % (gb, dt)); del x
    a = np.empty(n, np.uint8); a[::4096] = 1
    rt = torch.cuda.cudart()
    dt, _ = t(lambda: rt.cudaHostRegister(a.ctypes.data, n, 0)); print("hostRegister %d GB (touched): %.2f s" % (gb, dt))
    dt, _ = t(lambda: rt.cudaHostUnregister(a.ctypes.data)); print("hostUnregister %d GB: %.2f s" % (gb, dt))
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    dt, _ = t(lambda: torch.from_numpy(a).copy_(d)); print("D2H to pageable %d GB: %.2f s (%.1f GB/s)" % (gb, dt, gb * 1.0737 / dt))
    dt, _ = t(lambda: d.copy_(torch.from_numpy(a))); print("H2D from pageable %d GB: %.2f s (%.1f GB/s)" % (gb, dt, gb * 1.0737 / dt))
    del d, a
print("threads", torch.get_num_threads())
PY
