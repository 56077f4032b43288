"""Turn the scratch ncu outputs of one capture into the tracked extracts under profiles/.
usage: python tools/profile_extract.py <tag>   (expects gpurun_out/prof_<tag>.ncu-rep, gpurun_out/launches_<tag>.csv)"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1]
rep = f"gpurun_out/prof_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
pat = re.compile(r"^(Kernel Name|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_\w+)"
                 r"|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active"
                 r"|sm__pipe_(fma|alu|tensor)_cycles_active\.avg\.pct_of_peak_sustained_active"
                 r"|sm__inst_executed_pipe_(xu|lsu|tmem)\.avg\.pct_of_peak_sustained_active|dram__bytes_(read|write)\.sum"
                 r"|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active"
                 r"|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__throughput\.avg\.pct_of_peak_sustained_elapsed"
                 r"|lts__t_sector_hit_rate\.pct)$")
with open(f"profiles/{tag}_ncu_raw_extract.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:40] for r in data])
    for i, h in enumerate(hdr):
        if pat.match(h):
            w.writerow([h, units[i]] + [r[i] for r in data])
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
          "dram__bytes_write.sum"):
    i = hdr.index(k)
    print(k, units[i], [r[i] for r in data])
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = [round(float(r[i]), 3) for r in data]
        if max(v) > 0.1:
            print(f"{h[34:-23]:28s}", v)
try:
    lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    seq = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        tot[row["Kernel Name"]] += v
        cnt[row["Kernel Name"]] += 1
        if "cmcd::" in row["Kernel Name"]:
            seq.append((row["ID"], row["Kernel Name"], v))
    s = sum(tot.values())
    with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
        f.write("kernel,launches,total_ns,share_pct\n")
        for k, v in tot.most_common():
            f.write(f"\"{k}\",{cnt[k]},{v:.0f},{100 * v / s:.3f}\n")
    with open(f"profiles/{tag}_launches_cmcd_sequence.csv", "w") as f:
        f.write("id,kernel,gpu__time_duration_ns\n")
        for i, k, v in seq:
            f.write(f"{i},\"{k}\",{v:.0f}\n")
    for k, v in tot.most_common(4):
        print(f"{100 * v / s:6.2f}% {cnt[k]:4d} {k[:90]}")
except OSError:
    pass
for kern in ("fwd", "bwd"):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:bridge_{kern}_tc"], capture_output=True, text=True).stdout
    open(f"/tmp/{tag}_{kern}.csv", "w").write(src)
    out = subprocess.run([sys.executable, "tools/ncu_src_summary.py", f"/tmp/{tag}_{kern}.csv", str(131072 * 256)], capture_output=True, text=True).stdout
    open(f"profiles/{tag}_{kern}_tc_opmix_stalls.txt", "w").write(out)
    out = subprocess.run([sys.executable, "tools/ncu_cuda_lines.py", rep, f"bridge_{kern}_tc", "30"], capture_output=True, text=True).stdout
    open(f"profiles/{tag}_{kern}_tc_hot_lines.txt", "w").write(out)
