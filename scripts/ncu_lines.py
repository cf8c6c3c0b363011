"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo + --import-source on):
python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP]   -> executed warp instructions and stall samples per CUDA line"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        iex, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    elif hdr and r[0].isdigit():
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[iex] or 0), int(r[ist] or 0)))
        except ValueError:
            pass
tot_i, tot_s = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print(f"total executed {tot_i}, samples {tot_s}")
for f, ln, src, ex, st in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * ex / tot_i:5.1f}% inst {100 * st / tot_s:5.1f}% stall  {f}:{ln}  {src[:110]}")
