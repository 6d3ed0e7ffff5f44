#!/bin/bash
# A/B of the pipelined against the one-warp streaming form on the GPU box: GPU parity tests with the pipelined form, then timings.
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for p in 0 1; do
  for rep in 1 2; do
    BBD_PIPE=$p timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-full-step --steps 40 --warmup 10 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('BBD_PIPE=$p', 'step %.4f ms' % d['ms_per_step'], d['roofline']['kernel'], 'kernel %.4f ms' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
