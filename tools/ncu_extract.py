"""Summarise `ncu --set full` reports into the two files the repo commits under profiles/:

  python tools/ncu_extract.py OUT_PREFIX REPORT.ncu-rep[:label] [REPORT2.ncu-rep[:label] ...]

writes OUT_PREFIX.json ({"kernels": [...]}: what bench.py's roofline.traffic reads) and OUT_PREFIX.md
(a table per report).  Launches of the same kernel inside one report are averaged.  Runs here (no GPU
needed): it only reads reports brought back in gpurun_out/.
"""
import csv
import io
import json
import re
import subprocess
import sys

COLS = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "inst": "smsp__inst_executed.sum",
    "bank_conf": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l2_hit": "lts__t_sector_hit_rate.pct",
}
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6,
              "second": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def short(name):
    m = re.search(r"(?:<unnamed>::)?(\w+)(<[^>(]*>)?", name.replace("void ", ""))
    return (m.group(1) + (m.group(2) or "")) if m else name


def load(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {c: i for i, c in enumerate(hdr)}
    stalls = {STALL.match(c).group(1): i for c, i in ix.items() if STALL.match(c)}
    agg = {}
    for r in body:
        name = short(r[ix["Kernel Name"]])
        k = agg.setdefault(name, {"name": name, "launches": 0, "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]],
                                  "_s": {}, "_st": {}})
        k["launches"] += 1

        def val(col):
            i = ix.get(col)
            if i is None or r[i] in ("", "n/a"):
                return None
            return float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i].split("/")[0], 1.0)
        for key, col in COLS.items():
            v = val(col)
            if v is not None:
                k["_s"][key] = k["_s"].get(key, 0.0) + v
        for sname, i in stalls.items():
            try:
                k["_st"][sname] = k["_st"].get(sname, 0.0) + float(r[i])
            except ValueError:
                pass
    out = []
    for k in agg.values():
        n = k["launches"]
        e = {"name": k["name"], "launches": n, "grid": k["grid"], "block": k["block"]}
        s = {a: b / n for a, b in k["_s"].items()}
        e.update(duration_us=round(s.get("dur_us", 0), 2), dram_bytes_read=int(s.get("dram_rd", 0)),
                 dram_bytes_write=int(s.get("dram_wr", 0)), dram_pct=round(s.get("dram_pct", 0), 1),
                 issue_active_pct=round(s.get("issue_pct", 0), 1), fma_pipe_pct=round(s.get("fma_pct", 0), 1),
                 fp64_pipe_pct=round(s.get("fp64_pct", 0), 1), warps_active_pct=round(s.get("warps_pct", 0), 1),
                 regs=int(s.get("regs", 0)), smem_dynamic=int(s.get("smem_dyn", 0)), warp_inst=int(s.get("inst", 0)),
                 smem_bank_conflicts=int(s.get("bank_conf", 0)), l2_hit_pct=round(s.get("l2_hit", 0), 1))
        top = sorted(((v / n, a) for a, v in k["_st"].items() if a != "selected"), reverse=True)[:3]
        e["top_stalls"] = [[a, round(v, 2)] for v, a in top]
        out.append(e)
    return out


def main():
    prefix, reports = sys.argv[1], sys.argv[2:]
    allk, md = [], ["# ncu --set full --clock-control none extract (cold-cache, serialised launches: compare shares)", ""]
    for spec in reports:
        path, _, label = spec.partition(":")
        ks = load(path)
        for k in ks:
            k["capture"] = label or path
        allk += ks
        md += [f"## {label or path}", "",
               "| kernel | n | grid x block | regs | µs | DRAM rd MB | DRAM wr MB | DRAM % | issue % | FMA % | FP64 % | warps % | smem conflicts | top stalls (cycles / issue) |",
               "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
        for k in ks:
            md.append("| {name} | {launches} | {grid} x {block} | {regs} | {duration_us} | {rd:.2f} | {wr:.2f} | {dram_pct} | "
                      "{issue_active_pct} | {fma_pipe_pct} | {fp64_pipe_pct} | {warps_active_pct} | {smem_bank_conflicts} | {st} |"
                      .format(rd=k["dram_bytes_read"] / 1e6, wr=k["dram_bytes_write"] / 1e6,
                              st=", ".join(f"{a} {v}" for a, v in k["top_stalls"]), **k))
        md.append("")
    json.dump({"kernels": allk}, open(prefix + ".json", "w"), indent=1)
    open(prefix + ".md", "w").write("\n".join(md))
    print(f"{len(allk)} kernels -> {prefix}.json / .md")


if __name__ == "__main__":
    main()
