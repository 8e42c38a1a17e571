"""Generates tests/golden/*.npz from the REFERENCE's own code (run in the build container,
where /root/reference exists):

  spline_ref_f32.npz / spline_ref_f64.npz
      include/mitsuba/core/basisspline.h compiled verbatim (oracle/ref_spline.cpp ->
      oracle/_ref/libmer_refspline_{f,d}.so): prefiltered coefficients of a seeded random
      20x30x23 grid (the shape of mfiles/Test.m) and value / gradient / Hessian at seeded points.

Usage:  make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import RefSpline  # noqa: E402


def main():
    rng = np.random.default_rng(20201201)
    res = (20, 30, 23)
    data = (1.0 + rng.random((res[2], res[1], res[0]))).astype(np.float32)
    bmin = np.array([-1.0, -2.0, 0.5], np.float32)
    bmax = np.array([1.0, 1.0, 2.0], np.float32)
    pitch = (bmax - bmin) / (np.array(res, np.float32) - 1)
    lo, hi = bmin + 2.05 * pitch, bmax - 2.05 * pitch
    pts = (lo + rng.random((4096, 3)) * (hi - lo)).astype(np.float32)
    # a few exact knots (5-tap case)
    idx = rng.integers(3, [res[0] - 3, res[1] - 3, res[2] - 3], size=(64, 3))
    pts[:64] = (bmin + idx * pitch).astype(np.float32)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        ref = RefSpline(dt).build(data, res, bmin, bmax)
        f, g, H = ref.eval_hessian(pts.astype(dt))
        f2, g2 = ref.eval(pts.astype(dt), 2)
        assert np.array_equal(f, f2) and np.array_equal(g, g2)
        np.savez_compressed(os.path.join(HERE, "spline_ref_%s.npz" % tag), data=data, res=np.array(res), bbox_min=bmin,
                            bbox_max=bmax, coeff=ref.coefficients(), points=pts, value=f, gradient=g, hessian=H)
        print("wrote spline_ref_%s.npz" % tag)


if __name__ == "__main__":
    main()
