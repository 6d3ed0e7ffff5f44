#!/bin/bash
# per-call times of the non-fused kernels of a step, everything serialised on one stream: scripts/others.sh lib.so ...
for lib in "$@"; do
  BBD_SIDE_STREAMS=0 BBD_LIB=$lib python bench.py --no-cpu-baseline --no-e2e --no-graph --no-full-step --steps 20 --warmup 5 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'step %.4f' % d['ms_per_step'], ' '.join('%s=%.1f' % (o['call'][4:], o['us']) for o in d['roofline']['others']), 'reproj=%.1f' % (d['roofline']['kernel_ms']*1e3))"
done
