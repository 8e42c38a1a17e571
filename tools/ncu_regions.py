#!/usr/bin/env python
"""Where a kernel's issue slots go: per-region share of executed instructions, average active lanes and stall samples,
from `ncu -i X.ncu-rep --page source --csv --print-source sass`.  Regions are given as name=lo:hi (hex offsets from
the kernel's first instruction); without regions, prints the control-flow skeleton with offsets to pick them from."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
a0 = int(data[0][ix["Address"]], 16)


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
tot_t = sum(f(r, "Thread Instructions Executed") for r in data)
print("instructions %.3g, avg lanes %.2f, samples %d" % (tot_i, tot_t / tot_i, tot_s))
if len(sys.argv) == 2:
    for r in data:
        if re.search(r"BRA|BSSY|BSYNC|VOTE|WARPSYNC|CALL|EXIT|RET", r[ix["Source"]]):
            print("%6x  %-70s inst %5.2f%% lanes %4.1f" % (int(r[ix["Address"]], 16) - a0, r[ix["Source"]][:70],
                                                          100 * f(r, "Instructions Executed") / tot_i,
                                                          f(r, "Thread Instructions Executed") / max(f(r, "Instructions Executed"), 1)))
    sys.exit(0)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for spec in sys.argv[2:]:
    name, rng = spec.split("=")
    lo, hi = [int(x, 16) for x in rng.split(":")]
    sel = [r for r in data if lo <= int(r[ix["Address"]], 16) - a0 < hi]
    ie = sum(f(r, "Instructions Executed") for r in sel)
    te = sum(f(r, "Thread Instructions Executed") for r in sel)
    sm = sum(f(r, "# Samples") for r in sel)
    top = sorted(((sum(f(r, s) for r in sel), s) for s in stalls), reverse=True)[:3]
    print("%-26s inst %5.1f%%  lanes %5.1f  samples %5.1f%%  %s" % (name, 100 * ie / tot_i, te / max(ie, 1), 100 * sm / tot_s,
          ", ".join("%s %.1f%%" % (s[6:], 100 * v / tot_s) for v, s in top)))
