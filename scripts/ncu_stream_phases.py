"""Stall samples and executed instructions per phase of the streaming kernel from an ncu source-page export.

usage: ncu_stream_phases.py <report.ncu-rep> <lib.so> <kernel-mangled-substring>
An instruction belongs to the phase of the most recent csrc/bbd_stream.cuh line seen in address order (inlined
arithmetic helpers inherit it); phases are line ranges found from marker comments of stream_unit.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
src_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseboostdepth_b200", "csrc", "bbd_stream.cuh")
src = open(src_path).read().splitlines()
def line_of(marker, start=0):
    return next(i + 1 for i, l in enumerate(src) if i >= start and marker in l)
marks = [
    ("exact projection chain (stream_chain, division)", line_of("BBD_HD void div_exact2(float a1")),
    ("tap addresses + Jacobian pieces (stream_coords)", line_of("BBD_HD void stream_coords(")),
    ("tap gathers (stream_gather)", line_of("BBD_HD void stream_gather(")),
    ("unit set-up", line_of("BBD_HD void stream_unit(")),
    ("TMA issue / wait, planes in", line_of("// =============================== rows in")),
    ("bilinear + ring2 + horizontal sums (P2)", line_of("// =============================== P2: row r")),
    ("SSIM / mix / coefficients (stage B)", line_of("// =============================== stage B: row r-1")),
    ("per-pixel minimum, loss, winner", line_of("// per-pixel minimum: candidates in table order")),
    ("3x3 gather of the coefficients", line_of("// only the winner's coefficients survive")),
    ("backward to depth and P (stage C)", line_of("// =============================== stage C: row r-2")),
    ("segment epilogue + fused finalize", line_of("// ---- pose-gradient partials of this sweep")),
    ("identity pre-pass (other kernel)", line_of("// Identity pre-pass, streaming form")),
]
marks.sort(key=lambda m: m[1])
def phase_of(n):
    name = None
    for nm, ln in marks:
        if ln <= n:
            name = nm
    return name

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
insts = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
base = int(insts[0][0], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m:
        lines[int(m.group(1), 16)] = cur
phase = "unit set-up"
agg = collections.defaultdict(collections.Counter)
inst = collections.Counter()
for r in insts:
    k = lines.get(int(r[0], 16) - base)
    if k and k[0] == "bbd_stream.cuh":
        phase = phase_of(k[1]) or phase
    inst[phase] += int(r[col["Instructions Executed"]] or 0)
    for c in stall_cols:
        agg[phase][c] += int(r[col[c]] or 0)
tot = sum(sum(v.values()) for v in agg.values())
toti = sum(inst.values())
print(f"{'phase':52s} {'inst%':>6s} {'smpl%':>6s}  top stall reasons (share of the phase's samples)")
for ph, cnt in sorted(agg.items(), key=lambda x: -sum(x[1].values())):
    n = sum(cnt.values())
    if not n:
        continue
    top = ", ".join(f"{c[6:]} {100 * v / n:.0f}%" for c, v in cnt.most_common(4))
    print(f"{ph:52s} {100 * inst[ph] / toti:6.1f} {100 * n / tot:6.1f}  {top}")
