"""ad-hoc diagnostic (not a test): trajectory error growth GPU vs oracle32 vs oracle64"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mitsubaer_b200 as mer
from common import *
from oracle.oracle import Oracle, volume_desc
o32, o64 = Oracle(np.float32), Oracle(np.float64)
for kind in ("sd", "radial", "linear", "smooth"):
    h = 2e-3
    props = medium_props(stepsize=h)
    data, lo, hi = make_field(kind, 48)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).configure()
    d = volume_desc((48,)*3, lo, hi)
    m32 = o32.medium_create(oracle_medium_desc(props), o32.rif_create(d, data))
    m64 = o64.medium_create(oracle_medium_desc(props), o64.rif_create(d, data))
    n = 4000
    p0 = random_points_in_box(n, 21, margin=0.3) ; d0 = random_directions(n, 22)
    v0 = d0 * rif.value(p0)[:, None]
    # gradient accuracy
    f_g, g_g = rif.valueAndGradient(p0); f32, g32 = o32.rif_eval(m32 and o32.rif_create(d, data), p0, 2); f64, g64 = o64.rif_eval(o64.rif_create(d, data), p0.astype(np.float64), 2)
    print(kind, "grad abs err: gpu-64 %.2e  cpu32-64 %.2e  gpu-cpu32 %.2e | max|g| %.3f" % (np.abs(g_g-g64).max(), np.abs(g32-g64).max(), np.abs(g_g-g32).max(), np.abs(g64).max()))
    print(kind, "grad mean signed err: gpu-64 %s cpu32-64 %s" % ((g_g-g64).mean(axis=0), (g32-g64).mean(axis=0)))
    for steps in (10, 100, 300, 1000):
        dist = np.full(n, steps * h * 0.5, np.float32) if False else np.full(n, min(steps * h, 0.6), np.float32)
        g = med.trace(p0, v0, dist); a = o32.trace(m32, p0, v0, dist); b = o64.trace(m64, p0, v0, dist)
        ok = g["success"] & a["success"] & b["success"]
        e_ga = np.abs(g["p"][ok] - a["p"][ok]).max(axis=1); e_gb = np.abs(g["p"][ok] - b["p"][ok]).max(axis=1); e_ab = np.abs(a["p"][ok] - b["p"][ok]).max(axis=1)
        print("  steps %4d (n=%d): |p| gpu-cpu32 max %.2e p99 %.2e med %.2e | gpu-64 max %.2e med %.2e | cpu32-64 max %.2e med %.2e" % (int(dist[0]/h), ok.sum(), e_ga.max(), np.percentile(e_ga, 99), np.median(e_ga), e_gb.max(), np.median(e_gb), e_ab.max(), np.median(e_ab)))
