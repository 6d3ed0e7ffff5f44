#!/bin/bash
# Round-2 evidence on the GPU box: GPU tests (+ parity log), bench line, launch list, one --set full capture of the fused kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
bash scripts/launch_times.sh gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:reproj_stream -s 3 -c 1 -f -o gpurun_out/r02_stream_cur \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-full-step > gpurun_out/ncu_cur.log 2>&1
python -c "import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
cat gpurun_out/r02_launches_summary.txt
