"""small end-to-end run for compute-sanitizer (memcheck): prefilter, eval, trace, sampleDistance, render in both modes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mitsubaer_b200 as mer
from common import *
for mode in ("tricubic", "trilinear_packed"):
    data, lo, hi = make_field("radial", (20, 24, 28))
    rif = mer.SplineDataSource(data=data, min=lo, max=hi, mode=mode)
    grid = mer.GridDataSource(data=mer.fields.sine_density((12, 12, 12), BOX_MIN, BOX_MAX), min=BOX_MIN, max=BOX_MAX)
    p = np.concatenate([random_points_in_box(2000, 1) * 1.3, lo[None], hi[None]])   # incl. points outside the limits
    rif.valueAndGradient(p); rif.insideVolumeLimits(p); grid.lookupFloat(p)
    for den in (None, grid):
        props = medium_props(stepsize=2e-2, albedo=0.9, densityScale=6.0)
        med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.7))
        if den is not None: med.addChild("density", den)
        med.configure()
        p0 = random_points_in_box(3000, 2); v0 = random_directions(3000, 3) * 1.5
        med.trace(p0, v0, np.full(3000, 1.5, np.float32)); med.traceTillBoundary(p0[:500], v0[:500])
        if den is None:
            med.sampleDistance(p0, random_directions(3000, 4), 0.0, np.random.default_rng(5).random((3000, 2)))
        film, st = mer.EikonalVolPathIntegrator(stepsPerPass=64, poolPaths=512).render(scene_dict(24, 20, 4), med)
        print(mode, "density" if den is not None else "homogeneous", st["samples"], st["ray_steps"], st["passes"])
print("sanitize run ok")
