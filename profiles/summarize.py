#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into a small committed text summary under profiles/.

    python profiles/summarize.py gpurun_out/r1d.ncu-rep profiles/r1_tile_3km.txt "3 km mesh, tile path"

Reads the report with `ncu -i ... --page raw --csv` (metrics per captured launch) and
`--page source --csv --print-source sass` (per-instruction stall samples; needs -lineinfo / --import-source).
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALLS = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_math",
          "stall_mio", "stall_lg", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_sleep",
          "stall_membar", "stall_dispatch", "stall_drain", "stall_misc"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                          text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = ["# %s" % title, "# source report: %s (ncu --set full --clock-control none --import-source on)" % rep, ""]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    for r in data:
        lines.append("launch: %s" % r[kn][:100])
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                lines.append("  %-64s %s %s" % (m, r[i], units[i]))
        lines.append("")
    src = ncu(rep, "--page", "source", "--csv", "--print-source", "sass")
    kernels, cur, h = [], None, None
    for row in csv.reader(io.StringIO(src)):
        if len(row) >= 2 and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
            h = None
            continue
        if row and row[0] == "Address":
            h = row
            cur["hdr"] = h
            continue
        if cur is not None and h and len(row) == len(h):
            cur["rows"].append(row)
    if kernels and kernels[0]["rows"]:
        k = kernels[0]
        h = k["hdr"]
        iS, iI, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
        tot = sum(int(r[iS] or 0) for r in k["rows"]) or 1
        lines.append("warp stall samples of the first captured launch (%d samples, %d SASS instructions):"
                     % (tot, len(k["rows"])))
        for n in STALLS:
            if n in h:
                v = sum(int(r[h.index(n)] or 0) for r in k["rows"])
                if v:
                    lines.append("  %-26s %5.1f %%" % (n, 100.0 * v / tot))
        lines.append("")
        lines.append("hottest instructions (share of samples, executions, SASS):")
        top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iS] or 0))[:12]
        for i in sorted(top):
            r = k["rows"][i]
            lines.append("  %5.1f %%  ex=%-9s %s" % (100.0 * int(r[iS] or 0) / tot, r[iI], r[isrc][:90]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
