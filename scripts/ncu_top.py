"""Top stalled SASS instructions + headline metrics of one kernel from an .ncu-rep (run here, no GPU).
usage: ncu_top.py report.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, vals = raw[0], raw[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
for h, v in zip(hdr, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.05):
        print(f"{h:90s} {v}")
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hd = rows[hi]; col = {h: i for i, h in enumerate(hd)}
ins = [r for r in rows[hi + 1:] if len(r) >= len(hd) and r[0].startswith("0x")]
tot = sum(int(r[col["# Samples"]] or 0) for r in ins)
sc = [h for h in hd if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in ins:
    for c in sc: agg[c] = agg.get(c, 0) + int(r[col[c]] or 0)
print("samples", tot, {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
top = sorted(range(len(ins)), key=lambda i: -int(ins[i][col["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = ins[i]; s = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[c]] or 0), c) for c in sc), reverse=True)[:2]
    print(i, "%5.2f%%" % (100 * s / tot), r[col["Source"]][:78], [(c[6:], v) for v, c in st])
