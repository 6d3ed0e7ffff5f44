#!/bin/bash
# Per-kernel device times of one bench step (ncu, cold-cache serialised): scripts/launch_times.sh [out.csv]
out=${1:-gpurun_out/launches.csv}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[h]; ki, vi = H.index('Kernel Name'), H.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[h + 1:] if len(r) > vi]
ends = [i for i, s in enumerate(seq) if 'd2d_backward' in s[0]]
starts = [i for i, s in enumerate(seq) if 'd2d_forward_kernel' in s[0]]
a = max(i for i in starts if i < ends[-1]); b = ends[-1]
tot = 0
for n, v in seq[a:b + 1]:
    us = v / 1000 if v > 500 else v
    tot += us
    print("%9.1f us  %s" % (us, n[:90]))
print("%9.1f us  total of one step" % tot)
PY
