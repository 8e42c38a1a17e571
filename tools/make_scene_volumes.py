#!/usr/bin/env python
"""Writes the .vol files the scenes/*.xml refer to (the reference's MATLAB generators, restated in
mitsubaer_b200/fields.py): BoxRIF_Linear.vol, BoxRIF_Radial.vol, BoxSDRIF_2.vol, BoxDensity.vol, and SphereSDF.vol
(signed distance to a sphere: the `sdf` child volume of a mesh container, MER_SHAPE_SDF)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mitsubaer_b200 import fields  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes"))
    ap.add_argument("--res", type=int, default=64)
    a = ap.parse_args()
    lo_box, hi_box = (-1, -1, -1), (1, 1, 1)
    res = (a.res,) * 3
    lo, hi = fields.padded_bbox(lo_box, hi_box, res)
    os.makedirs(a.out, exist_ok=True)
    fields.write_vol(os.path.join(a.out, "BoxRIF_Linear.vol"), fields.linear_rif(res, lo, hi), lo, hi)
    fields.write_vol(os.path.join(a.out, "BoxRIF_Radial.vol"), fields.radial_rif(res, lo, hi), lo, hi)
    fields.write_vol(os.path.join(a.out, "BoxSDRIF_2.vol"), fields.rif_from_sd(fields.sphere_sdf(res, lo, hi, radius=0.8)), lo, hi)
    fields.write_vol(os.path.join(a.out, "BoxDensity.vol"), fields.sine_density(res, lo_box, hi_box), lo_box, hi_box)
    fields.write_vol(os.path.join(a.out, "SphereSDF.vol"), fields.sphere_sdf(res, lo, hi, radius=0.8).astype(np.float32), lo, hi)
    print("wrote 5 volumes (%d^3) to %s" % (a.res, a.out))


if __name__ == "__main__":
    main()
