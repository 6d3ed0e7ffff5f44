"""Stall-reason samples per kernel phase from an ncu source-page export.

usage: ncu_phase_stalls.py <report.ncu-rep> <cubin> <kernel-substring> <header-with-phases>
Phases = the BBD_HD functions of the header; an instruction belongs to the phase of the most
recent line of that header seen in address order (inlined helpers inherit it).
"""
import csv, re, subprocess, sys, os, collections

rep, cubin, kern, header = sys.argv[1:5]
src = open(header).read().splitlines()
funcs = [(i + 1, re.search(r"BBD_HD\s+\S+\s+\*?(\w+)\(", l).group(1)) for i, l in enumerate(src) if re.match(r"^BBD_HD ", l)]
def phase_of_line(n):
    name = None
    for start, f in funcs:
        if start <= n:
            name = f
    return name
SKIP = {"w9", "w9p", "w9_2", "w9pp_2", "ld2", "rs_center", "rs_candidate", "rs_xch", "make_strip", "is_owned"}
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out)); hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
insts = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(insts[0][0], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
line_of = {}; cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m: line_of[int(m.group(1), 16)] = cur
phase = "prologue"; agg = collections.defaultdict(lambda: collections.Counter()); inst = collections.Counter()
for r in insts:
    k = line_of.get(int(r[0], 16) - base)
    if k and k[0] == os.path.basename(header):
        f = phase_of_line(k[1])
        if f and f not in SKIP: phase = f
    inst[phase] += int(r[col["Instructions Executed"]] or 0)
    for c in stall_cols:
        agg[phase][c] += int(r[col[c]] or 0)
tot = sum(sum(v.values()) for v in agg.values()); toti = sum(inst.values())
print(f"{'phase':18s} {'inst%':>6s} {'smpl%':>6s}  top stall reasons (share of the phase's samples)")
for ph, cnt in sorted(agg.items(), key=lambda x: -sum(x[1].values())):
    n = sum(cnt.values())
    if not n: continue
    top = ", ".join(f"{c[6:]} {100 * v / n:.0f}%" for c, v in cnt.most_common(5))
    print(f"{ph:18s} {100 * inst[ph] / toti:6.1f} {100 * n / tot:6.1f}  {top}")
