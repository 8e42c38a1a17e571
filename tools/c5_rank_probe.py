"""One rank's share of the C5 strong-scaling frame on ONE GPU (no torchrun): what a rank of an N-GPU run renders, timed alone.
Usage: python tools/c5_rank_probe.py <world> [pool]   (development probe: per-rank rate against samples per pixel per rank)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import mitsubaer_b200 as mer  # noqa: E402

world = int(sys.argv[1])
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = bench.WORKLOADS["C5"]
dev = torch.device("cuda:0")
rif_t, lo, hi, den_t = bench.make_fields(w, xp=torch, device=dev)
rif = mer.SplineDataSource(data_ptr=rif_t.data_ptr(), res=w["rif_res"], min=lo, max=hi, device=0, mode="tricubic")
del rif_t
grid = mer.GridDataSource(data_ptr=den_t.data_ptr(), res=(w["den_res"],) * 3, min=bench.BOX_MIN, max=bench.BOX_MAX, device=0)
del den_t
med = mer.HeterogeneousRefractiveMedium(bench.medium_props(w)).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=w["g"]))
med = med.addChild("density", grid).configure()
integ = mer.EikonalVolPathIntegrator(maxDepth=w["max_depth"], rrDepth=5, poolPaths=pool)
film = torch.zeros(w["height"], w["width"], 5, device=dev)
stream = torch.cuda.current_stream().cuda_stream
integ.render_device(bench.scene_dict(w, max(64 // 8, world)), med, film.data_ptr(), stream=stream, sample_begin=0, sample_stride=world)
torch.cuda.synchronize()
film.zero_()
t0 = time.time()
st = integ.render_device(bench.scene_dict(w, 64), med, film.data_ptr(), stream=stream, sample_begin=0, sample_stride=world)
torch.cuda.synchronize()
print(json.dumps({"world": world, "pool": pool, "wall_s": time.time() - t0, "device_ms": st["device_ms"], "rounds": st["passes"],
                  "G_steps_per_s": st["ray_steps"] / st["device_ms"] / 1e6, "lanes": st["step_lanes_per_sm"], "tail_ms": st["tail_ms"]}))
