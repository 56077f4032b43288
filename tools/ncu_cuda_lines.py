"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo + --import-source on at capture time).
usage: python tools/ncu_cuda_lines.py <report.ncu-rep> <kernel regex> [top]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, cur_file, out = None, "", []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 1 and r[0] not in ("",):
        d = dict(zip(hdr, r))
        # hdr has two "Source" columns (cuda line, sass); dict keeps the last; the cuda text is r[1]
        try:
            out.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), cur_file, r[0], r[1].strip()[:100],
                        int(d.get("stall_long_sb") or 0), int(d.get("stall_wait") or 0), int(d.get("stall_short_sb") or 0),
                        int(d.get("stall_barrier") or 0)))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
tin = sum(o[1] for o in out) or 1
print(f"total samples {tot}, warp instructions {tin:.3e}")
print("samples%  instr%  file:line  [long_sb wait short_sb barrier]  source")
for s, n, f, ln, src, lsb, w, ssb, bar in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / tot:6.2f} {100 * n / tin:6.2f}  {f}:{ln:>4} [{lsb:6d} {w:6d} {ssb:6d} {bar:6d}] {src}")
