#!/usr/bin/env python
"""FP32 cost of one ray step, counted in the SASS of the step kernel (k_step) of the built library.

    python tools/sass_costs.py            # writes profiles/r02_step_costs.json and profiles/r02_k_step_sass.txt

The step kernel is one loop; everything floating-point in it belongs to the step (first kick + drift, cell, weights,
separable contraction, second kick, containment test); the refill code and the coefficient gathers are integer / memory
instructions.  FLOP per ray step = 2 x FFMA + FMUL + FADD + MUFU between the loop head and its backward branch.
bench.py multiplies this by the ray steps it counted on the device to report the FP32 ceiling of the roofline block.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mitsubaer_b200", "libmitsubaer_b200.so")
VARIANTS = {"tricubic": "k_stepILi0ELi0ELb0ELb0ELb0E", "trilinear_packed": "k_stepILi1ELi0ELb0ELb0ELb0E"}


def sass(fn_pattern):
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    full = [ln.split()[-1] for ln in names.splitlines() if "Function :" in ln and fn_pattern in ln]
    if not full:
        raise SystemExit("kernel %s not found in %s" % (fn_pattern, LIB))
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", full[0], LIB], capture_output=True, text=True).stdout
    ins = []
    for ln in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return full[0], ins


def main():
    out, listing = {}, []
    for mode, pat in VARIANTS.items():
        name, ins = sass(pat)
        # the loop: the last backward branch and its target
        back = [(a, int(re.search(r"0x([0-9a-f]+)", t).group(1), 16)) for a, t in ins if re.match(r"(@!?U?P\d+ )?BRA(\.U)? ", t) and
                re.search(r"0x([0-9a-f]+)", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
        end, head = max(back, key=lambda b: b[0] - b[1])
        body = [(a, t) for a, t in ins if head <= a <= end]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+ ", "", t).split()[0].split(".")[0] for _, t in body)
        flop = 2 * ops["FFMA"] + ops["FMUL"] + ops["FADD"] + ops["MUFU"]
        out[mode] = {"kernel": name, "loop": "0x%x-0x%x" % (head, end), "instructions_in_loop": len(body), "FFMA": ops["FFMA"], "FMUL": ops["FMUL"],
                     "FADD": ops["FADD"], "MUFU": ops["MUFU"], "TLD4": ops["TLD4"], "LDG": ops["LDG"], "flop_per_step": float(flop),
                     "source": "tools/sass_costs.py on libmitsubaer_b200.so: 2 x FFMA + FMUL + FADD + MUFU in the step loop"}
        listing.append("==== %s (%s): loop 0x%x-0x%x, %d instructions\n" % (mode, name, head, end, len(body)))
        listing += ["  /*%04x*/ %s\n" % (a, t) for a, t in body]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_step_costs.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", "r02_k_step_sass.txt"), "w").writelines(listing)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
