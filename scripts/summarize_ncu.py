#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/launches_X_summary.md "<command that was profiled>"
  python scripts/summarize_ncu.py full gpurun_out/prof_X.ncu-rep profiles/prof_X_summary.md
  python scripts/summarize_ncu.py stalls gpurun_out/prof_X.ncu-rep profiles/prof_X_stalls.md <kernel regex> [launch index]
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(u, v)


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi, ui = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        ms = to_ms(r[vi], r[ui])
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("command: `%s`\n\nsource: `%s` (%d launches, %.3f ms of kernel time; per-launch times are cold-cache and serialised: "
                "compare SHARES, not absolutes)\n\n" % (cmd, src, sum(a[0] for a in agg.values()), tot))
        f.write("| kernel | launches | sum ms | share | max ms | mean ms |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| %s | %d | %.3f | %.1f%% | %.3f | %.4f |\n" % (k, a[0], a[1], 100 * a[1] / tot, a[2], a[1] / a[0]))
    print(open(dst).read())


def full(src, dst):
    if src.endswith(".csv"):      # `ncu -i X.ncu-rep --page raw --csv` already run on the GPU box (the report itself was too big to bring back)
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % src)
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "")
            f.write("## %s  (id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (name, r[0]))
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            try:
                t = to_ms(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
                rd *= scale[units[hdr.index("dram__bytes_read.sum")]]
                wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
                f.write("| derived: dram traffic per launch | %.4f | GB |\n| derived: dram GB/s | %.1f | GB/s |\n" %
                        ((rd + wr) / 1e9, (rd + wr) / 1e9 / (t * 1e-3)))
            except Exception:
                pass
            f.write("\n")
    print(open(dst).read())


def stalls(src, dst, kernel, launch="0"):
    """Per-instruction warp-stall samples (ncu source page) of one launch: stall-reason totals + the hottest SASS lines."""
    if src.endswith(".csv.gz"):   # `ncu -i X.ncu-rep --page source --csv | gzip` already run on the GPU box (one kernel per file)
        import gzip
        raw = gzip.open(src, "rt").read()
    elif src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-skip", launch,
                              "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kname = rows[0][1] if rows and len(rows[0]) > 1 else kernel
    hdr = rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:          # the csv repeats every SASS row (once per source view): keep the first by address
        if len(r) == len(hdr) and r[idx["Address"]] not in seen:
            seen.add(r[idx["Address"]])
            data.append(r)

    def val(r, n):
        try:
            return int(r[idx[n]] or 0)
        except ValueError:
            return 0
    names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(val(r, "# Samples") for r in data) or 1
    with open(dst, "a") as f:
        f.write("\n# warp-stall samples per SASS instruction: `%s` (launch %s of `%s`)\n\n" % (kname, launch, src))
        f.write("total samples %d over %d SASS instructions\n\n| stall reason | samples | share |\n|---|---:|---:|\n" % (tot, len(data)))
        for n, v in sorted(((n, sum(val(r, n) for r in data)) for n in names), key=lambda x: -x[1])[:8]:
            f.write("| %s | %d | %.1f%% |\n" % (n, v, 100.0 * v / tot))
        f.write("\n| samples | share | executed | SASS | top stall |\n|---:|---:|---:|---|---|\n")
        for r in sorted(data, key=lambda r: -val(r, "# Samples"))[:25]:
            top = max(names, key=lambda n: val(r, n))
            f.write("| %d | %.1f%% | %s | `%s` | %s |\n" % (val(r, "# Samples"), 100.0 * val(r, "# Samples") / tot,
                                                       r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:70], top))
    print(open(dst).read()[-5000:])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif sys.argv[1] == "stalls":
        stalls(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "0")
    else:
        full(sys.argv[2], sys.argv[3])
