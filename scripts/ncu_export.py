"""Export the judged summaries of one `ncu --set full` capture into profiles/:
  <tag>_raw.csv (raw page), <tag>_by_line.txt (instructions by source line / opcode), <tag>_key.txt (headline metrics)
and refresh profiles/traffic.json for the workload.

usage: ncu_export.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> <tag> [workload]
"""
import csv, json, os, subprocess, sys

rep, lib, kern, tag = sys.argv[1:5]
workload = sys.argv[5] if len(sys.argv) > 5 else "kitti_640x192_b12_pm1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(prof, tag + "_raw.csv"), "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"] + sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
with open(os.path.join(prof, tag + "_key.txt"), "w") as f:
    for k in keys:
        if k in m:
            f.write("%-90s %s %s\n" % (k, m[k][0], m[k][1]))
def num(k):
    v, u = m[k]
    x = float(v.replace(",", ""))
    return x * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]
traffic = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
tj = os.path.join(prof, "traffic.json")
t = json.load(open(tj))
t[workload] = traffic
t["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum of one launch of the fused kernel (ncu --set full, profiles/%s_raw.csv); bench.py copies the entry of its workload into roofline.traffic" % tag
json.dump(t, open(tj, "w"), indent=2)
by = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_by_line.py"), rep, lib, kern, "60"], capture_output=True, text=True).stdout
open(os.path.join(prof, tag + "_by_line.txt"), "w").write(by)
print(open(os.path.join(prof, tag + "_key.txt")).read())
print("traffic", traffic)
