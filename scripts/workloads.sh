#!/bin/bash
# device-resident step and fused-kernel time of every benchmarked workload (bench.py, CUDA-graph replay): scripts/workloads.sh [lib.so]
lib=${1:-}
for wl in kitti_640x192_b12_pm1 trimin_mixed_640x192_b12 trimin_decomp_640x192_b12 trimin_all3_640x192_b12 hires_1024x320_b8_pm1; do
  BBD_LIB=$lib python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-full-step --steps 30 --warmup 8 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', 'step %.4f ms' % d['ms_per_step'], '%.2f G/s' % (d['value']/1e9), r['kernel'], '%.4f' % r['kernel_ms'], 'frac %.3f' % r['frac'])"
done
