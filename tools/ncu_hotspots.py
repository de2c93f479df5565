"""Source-level hot spots of `ncu --set full --import-source on` reports (kernels compiled with -lineinfo):

  python tools/ncu_hotspots.py OUT.md REPORT.ncu-rep[:label] ...

For every kernel of every report: the source lines (file:line of this repo) that collected the most warp-stall
samples, with their share of the kernel's samples and of its executed instructions.  Runs here (no GPU needed)."""
import csv
import io
import subprocess
import sys


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    kernels, cur, file_ = {}, None, None
    rows = list(csv.reader(io.StringIO(out)))
    i = 0
    while i < len(rows):
        r = rows[i]
        if len(r) >= 2 and r[0] == "File Path":
            file_ = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            cur = kernels.setdefault(r[1], [])
        elif r and r[0] == "Line No":
            hdr = r
            il, isrc = hdr.index("Line No"), hdr.index("Source")
            isamp, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            i += 1
            while i < len(rows) and rows[i] and rows[i][0] not in ("File Path", "Function Name", "Line No", "Kernel Name"):
                q = rows[i]
                if q[il].isdigit() and q[2] == "-":  # a CUDA source line (SASS rows carry an address)
                    try:
                        cur.append((file_, int(q[il]), q[isrc].strip(), int(q[isamp] or 0), int(q[iex] or 0)))
                    except ValueError:
                        pass
                i += 1
            continue
        i += 1
    return kernels


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    md = ["# Source-level hot spots (ncu `--set full --import-source on`, warp-stall samples per CUDA source line)\n",
          "Shares are of the kernel's own samples / executed warp instructions. Inlined device functions are attributed to",
          "their own file:line (`fft_dit.cuh` = the FFT butterflies, `tma_stage.cuh` = mbarrier waits).\n"]
    for spec in reps:
        rep, _, label = spec.partition(":")
        for kname, lines in source_page(rep).items():
            tot_s = sum(l[3] for l in lines) or 1
            tot_e = sum(l[4] for l in lines) or 1
            if tot_s < 20:
                continue
            short = kname.replace("void ", "").replace("<unnamed>::", "").split("(")[0]
            md.append(f"## {label or rep}: `{short}` ({tot_s} samples, {tot_e} warp instructions)\n")
            md.append("| file:line | samples % | instr % | source |")
            md.append("|---|---|---|---|")
            agg = {}
            for f, ln, src, s, e in lines:
                k = (f.split("/")[-1], ln)
                a = agg.setdefault(k, [src, 0, 0])
                a[1] += s
                a[2] += e
            for (f, ln), (src, s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
                md.append(f"| {f}:{ln} | {100 * s / tot_s:.1f} | {100 * e / tot_e:.1f} | `{src[:110].replace('|', '/')}` |")
            md.append("")
    open(out, "w").write("\n".join(md) + "\n")
    print(f"wrote {out}")


if __name__ == "__main__":
    main()
