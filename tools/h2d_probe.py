import time, torch
for gb in (4, 16):
    n = gb * (1 << 30) // 4
    t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float32, pin_memory=True); t1 = time.perf_counter()
    d = torch.empty(n, dtype=torch.float32, device="cuda"); torch.cuda.synchronize()
    for _ in range(2):
        t2 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        t4 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"{gb} GB: pin alloc {t1-t0:.2f}s  H2D {gb/(t3-t2):.1f} GB/s  D2H {gb/(t5-t4):.1f} GB/s")
    del h, d
