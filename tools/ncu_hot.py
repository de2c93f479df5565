"""List the hottest SASS lines (warp stall samples) of one kernel in an ncu report:
   python tools/ncu_hot.py REPORT.ncu-rep KERNEL_REGEX [min_pct]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.2
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
hi = [n for n, x in enumerate(r) if "Source" in x and "# Samples" in x][0]
h = r[hi]
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
rows = []
for x in r[hi + 1:]:
    if len(x) != len(h) or not x[isamp].isdigit():
        if "Source" in x:
            break  # next kernel instance
        continue
    rows.append(x)
tot = sum(int(x[isamp]) for x in rows)
print("total samples", tot, "SASS lines", len(rows), "warp instr", sum(int(x[iex]) for x in rows if x[iex].isdigit()))
cum = 0
for n, x in enumerate(rows):
    s = int(x[isamp]); cum += s
    if s > tot * minp / 100:
        print(f"{n:5d} {x[ia][:64].strip():64s} {s:6d} {100*s/tot:5.1f}% exec {x[iex]:>8s} cum {100*cum/tot:5.1f}%")
