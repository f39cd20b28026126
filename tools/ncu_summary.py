"""Summarise an .ncu-rep (raw metrics + SASS hot spots) into text for profiles/."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.max', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name)
    for i, h in enumerate(hdr):
        if h in keys:
            print("  %-62s %s %s" % (h, r[i], units[i]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
# may contain several kernels; take the first block
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for bi, st in enumerate(start):
    end = start[bi + 1] - 1 if bi + 1 < len(start) else len(rows)
    h = rows[st]; data = [r for r in rows[st + 1:end] if len(r) == len(h)]
    ci, si, so = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    tot = sum(int(r[ci]) for r in data); tots = max(1, sum(int(r[si]) for r in data))
    print("SASS instructions:", len(data), "executed:", tot, "samples:", tots)
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = sorted(((sum(int(r[i] or 0) for r in data), h[i]) for i in stall), reverse=True)
    print("  stall reasons:", ", ".join("%s %.1f%%" % (n, 100.0 * v / tots) for v, n in agg[:7]))
    ops = {}
    for r in data:
        op = r[so].split()[0] if not r[so].strip().startswith("@") else r[so].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ci])
    print("  top opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:14]))
    if len(sys.argv) > 2:
        seg = int(sys.argv[2])
        for a in range(0, len(data), seg):
            s = data[a:a + seg]
            print("  sass[%4d:%4d] inst %.1f%% samples %.1f%%  first: %s" % (a, a + seg, 100.0 * sum(int(r[ci]) for r in s) / tot,
                  100.0 * sum(int(r[si]) for r in s) / tots, s[0][so][:50]))
