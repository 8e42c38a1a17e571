"""ad-hoc probe: where does the time of one render call go"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mitsubaer_b200 as mer
import bench
w = dict(bench.WORKLOADS["C2"]); spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
rif_np, lo, hi, den_np = bench.make_fields(w)
rif = mer.SplineDataSource(data=rif_np, min=lo, max=hi)
grid = mer.GridDataSource(data=den_np, min=bench.BOX_MIN, max=bench.BOX_MAX)
med = mer.HeterogeneousRefractiveMedium(bench.medium_props(w)).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).addChild("density", grid).configure()
scene = bench.scene_dict(w, spp)
film = torch.zeros(w["height"], w["width"], 5, device=dev)
for spass, pool in ((2048, 0),):
    integ = mer.EikonalVolPathIntegrator(maxDepth=64, rrDepth=5, stepsPerPass=spass, poolPaths=pool)
    for rep in range(2):
        film.zero_(); torch.cuda.synchronize(); t0 = time.time()
        st = integ.render_device(scene, med, film.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize(); dt = time.time() - t0
    print("steps/pass %5d pool %7d: wall %.3f s device_ms %.1f passes %d  -> %.2f Gsteps/s (wall) %.2f (events)" % (spass, pool, dt, st["device_ms"], st["passes"], st["ray_steps"]/dt/1e9, st["ray_steps"]/st["device_ms"]/1e6))
