#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
into profiles/rNN_launch_summary.txt and the per-launch DRAM traffic that bench.py reports as roofline.traffic.

    python tools/summarize_launches.py gpurun_out/r01_launches.csv --command "python bench.py ..." \
        --summary profiles/r01_launch_summary.txt --traffic profiles/r01_traffic.json --key C2:tricubic
"""
import argparse
import collections
import csv
import json


def to_unit(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0,
             "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}
    return v * scale[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--command", default="")
    ap.add_argument("--summary")
    ap.add_argument("--traffic")
    ap.add_argument("--key", default="C2:tricubic")
    ap.add_argument("--kernel", default="k_render_pass")
    a = ap.parse_args()
    lines = [l for l in open(a.csv) if l.startswith('"')]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = launches.setdefault(row["ID"], {"name": row["Kernel Name"], "grid": row["Grid Size"]})
        d[row["Metric Name"]] = to_unit(row["Metric Value"], row["Metric Unit"])
    per = collections.OrderedDict()
    for d in launches.values():
        k = per.setdefault(d["name"], {"n": 0, "ms": 0.0, "bytes": 0.0})
        k["n"] += 1
        k["ms"] += d.get("gpu__time_duration.sum", 0.0)
        k["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total = sum(k["ms"] for k in per.values())
    out = ["command: %s   (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)" % a.command]
    for name, k in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        out.append("%-64s launches=%4d  time=%10.3f ms (%5.1f%%)  dram=%8.2f GB (%.2f GB/launch)" %
                   (name[:64], k["n"], k["ms"], 100 * k["ms"] / total, k["bytes"] / 1e9, k["bytes"] / 1e9 / k["n"]))
    passes = [d for d in launches.values() if a.kernel in d["name"]]
    if passes:
        full_grid = max(passes, key=lambda d: d.get("gpu__time_duration.sum", 0))["grid"]
        full = [d for d in passes if d["grid"] == full_grid]
        byt = [d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in passes]
        fbyt = [d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in full]
        out.append("render passes: %d launches, %d with the full grid %s avg %.2f ms, avg DRAM traffic %.2f GB per full-grid pass" %
                   (len(passes), len(full), full_grid, sum(d["gpu__time_duration.sum"] for d in full) / len(full), sum(fbyt) / len(fbyt) / 1e9))
        if a.traffic:
            try:
                t = json.load(open(a.traffic))
            except (OSError, ValueError):
                t = {}
            t[a.key] = sum(byt) / len(byt)
            t["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per %s launch, mean over all %d launches of `%s` under ncu (%s); "
                          "full-grid passes average %.3g bytes" % (a.kernel, len(passes), a.command, a.csv, sum(fbyt) / len(fbyt)))
            json.dump(t, open(a.traffic, "w"), indent=1)
    text = "\n".join(out) + "\n"
    if a.summary:
        open(a.summary, "w").write(text)
    print(text, end="")


if __name__ == "__main__":
    main()
