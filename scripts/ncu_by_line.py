"""Attribute an ncu source-page export (SASS) to CUDA source lines / inlined functions.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-mangled-substring>
Reads per-instruction executed counts and stall samples from the report and the
line table from nvdisasm -g; prints totals per source function-ish line ranges.
"""
import csv, re, subprocess, sys, collections, os, tempfile

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
sass_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(sass_csv))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
insts = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    insts.append((int(r[0], 16), r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                  int(r[col["Thread Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0)))
base = insts[0][0]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur

by_line = collections.Counter(); by_line_samples = collections.Counter(); by_op = collections.Counter()
tot_i = tot_s = 0
for addr, src, ni, nti, ns in insts:
    key = line_of.get(addr - base)
    k = (key[0], key[1]) if key else ("?", 0)
    by_line[k] += ni; by_line_samples[k] += ns
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    by_op[op.split(".")[0]] += ni
    tot_i += ni; tot_s += ns
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
print("--- by source line (file:line  inst%  samples%)")
for k, v in by_line.most_common(top):
    print(f"{k[0]:>18s}:{k[1]:<5d} {100*v/tot_i:6.2f}%  {100*by_line_samples[k]/max(1,tot_s):6.2f}%")
print("--- by opcode")
for k, v in by_op.most_common(25):
    print(f"{k:>12s} {100*v/tot_i:6.2f}%")

# --- by outermost call site in the kernel body (phase attribution)
by_phase = collections.Counter(); by_phase_s = collections.Counter()
for addr, src, ni, nti, ns in insts:
    key = line_of.get(addr - base)
    if not key:
        k = ("?", 0)
    else:
        chain = [(key[0], key[1])] + list(key[2])
        k = chain[-1]
    by_phase[k] += ni; by_phase_s[k] += ns
print("--- by kernel-level call site (file:line inst% samples%)")
for k, v in by_phase.most_common(20):
    print(f"{k[0]:>18s}:{k[1]:<5d} {100*v/tot_i:6.2f}%  {100*by_phase_s[k]/max(1,tot_s):6.2f}%")
