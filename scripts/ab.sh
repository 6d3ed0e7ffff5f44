#!/bin/bash
# A/B kernel timing of library variants on the GPU box: scripts/ab.sh [workload] lib1.so lib2.so ...
# prints ms/step and the reproj kernel's ms for each (bench.py device-resident leg only).
wl=kitti_640x192_b12_pm1
if [[ "$1" != *.so ]]; then wl=$1; shift; fi
for lib in "$@"; do
  for rep in 1 2; do
    BBD_LIB=$lib python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-graph --steps 40 --warmup 10 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$wl', 'step %.4f ms' % d['ms_per_step'], 'kernel %.4f ms' % d['roofline']['kernel_ms'])"
  done
done
