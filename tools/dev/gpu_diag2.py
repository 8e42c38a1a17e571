import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mitsubaer_b200 as mer
from common import *
from oracle.oracle import Oracle, volume_desc
o32 = Oracle(np.float32)
kind, h = "sd", 2e-3
props = medium_props(stepsize=h)
data, lo, hi = make_field(kind, 48)
rif = mer.SplineDataSource(data=data, min=lo, max=hi)
med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).configure()
d = volume_desc((48,)*3, lo, hi)
orif = o32.rif_create(d, data)
m32 = o32.medium_create(oracle_medium_desc(props), orif)
n = 20000
p0 = random_points_in_box(n, 21, margin=0.02); d0 = random_directions(n, 22)
v0 = d0 * rif.value(p0)[:, None]
dist = (np.random.default_rng(23).random(n) * 2.0).astype(np.float32)
g = med.trace(p0, v0, dist); a = o32.trace(m32, p0, v0, dist)
same = (g["success"] == a["success"]) & (g["nsteps"] == a["nsteps"])
err = np.abs(g["p"] - a["p"]).max(axis=1); err[~same] = 0
idx = np.argsort(-err)[:8]
for i in idx:
    # march the oracle in 50-step chunks and the gpu too, to find where they separate
    print("ray", i, "err %.2e" % err[i], "p0", p0[i], "dist %.3f" % dist[i], "steps", g["nsteps"][i], "success", g["success"][i])
    for frac in (0.1, 0.25, 0.5, 0.75, 1.0):
        dd = np.array([dist[i] * frac], np.float32)
        gg = med.trace(p0[i:i+1], v0[i:i+1], dd); aa = o32.trace(m32, p0[i:i+1], v0[i:i+1], dd)
        pe = aa["p"][0]
        f, gr = rif.valueAndGradient(pe[None]); fo, go = o32.rif_eval(orif, pe[None], 2)
        print("   frac %.2f |dp| %.2e |dv| %.2e  r=%.4f  |p|inf=%.4f  n %.5f gradgpu-gradcpu %.2e" % (frac, np.abs(gg["p"]-aa["p"]).max(), np.abs(gg["v"]-aa["v"]).max(), np.linalg.norm(pe), np.abs(pe).max(), f[0], np.abs(gr-go).max()))
# gradient error distribution near the sphere surface r=0.8
pts = random_directions(200000, 5) * (0.8 + (np.random.default_rng(6).random((200000,1)) - 0.5) * 0.1).astype(np.float32)
f, gr = rif.valueAndGradient(pts); fo, go = o32.rif_eval(orif, pts, 2)
print("near surface: max grad diff %.2e, max val diff %.2e" % (np.abs(gr-go).max(), np.abs(f-fo).max()))
pts = random_points_in_box(200000, 7) * 0.08
f, gr = rif.valueAndGradient(pts); fo, go = o32.rif_eval(orif, pts, 2)
print("near centre: max grad diff %.2e, max val diff %.2e, max |g| %.3f" % (np.abs(gr-go).max(), np.abs(f-fo).max(), np.abs(go).max()))
