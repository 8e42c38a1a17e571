"""ad-hoc GPU timing probe (not a test): trace kernel + small render throughput"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mitsubaer_b200 as mer
from common import *

def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    data, lo, hi = make_field("radial", res)
    for mode in ("tricubic", "trilinear_packed"):
        t = time.time()
        rif = mer.SplineDataSource(data=data, min=lo, max=hi, mode=mode)
        print(mode, "create s", time.time() - t)
        for h in (2e-2, 2e-3):
            med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=h)).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).configure()
            n = 1 << 20
            p0 = random_points_in_box(n, 1, margin=0.01); d0 = random_directions(n, 2)
            v0 = d0 * 1.5
            dist = np.full(n, 200 * h, np.float32)
            med.trace(p0[:1000], v0[:1000], dist[:1000])
            t = time.time(); out = med.trace(p0, v0, dist); dt = time.time() - t
            print("  trace h=%g: %.3f s wall (incl. copies), steps %d -> %.2f Gsteps/s" % (h, dt, out["nsteps"].sum(), out["nsteps"].sum() / dt / 1e9))
        med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=2e-3)).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).configure()
        scene = scene_dict(256, 256, 16, rfilter="box")
        for spp_pass in (512, 2048):
            film, st = mer.EikonalVolPathIntegrator(stepsPerPass=spp_pass).render(scene, med)
            print("  render steps/pass %d: %s -> %.2f Gsteps/s, %.2f Msamples/s" % (spp_pass, st, st["ray_steps"] / st["device_ms"] / 1e6, st["samples"] / st["device_ms"] / 1e3))

main()
