#!/usr/bin/env python
"""BASELINE.json config 4: eikonal step-size sweep on a 256^3 RIF — ray-steps/s of the batch stepper
(`mer_medium_trace_device`, kernel k_trace) against the fetch roofline.

2^24 rays (Philox-free: torch generator, seed 20201201) with uniform-random origins in the box and uniform
directions, fixed arc length 1.0 * extent, h in {1e-2, 3e-3, 1e-3, 3e-4, 1e-4} * extent, no scattering.
Rays live in HBM; timing is CUDA events around the kernel on torch's current stream (3 warm-ups, then the
median of 5 runs on fresh copies of the rays).  One JSON line per (mode, h).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mitsubaer_b200 as mer  # noqa: E402
from mitsubaer_b200._abi import check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--modes", default="tricubic,trilinear_packed")
    ap.add_argument("--fractions", default="1e-2,3e-3,1e-3,3e-4,1e-4")
    ap.add_argument("--field", default="radial")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    box_min, box_max = np.array([-1.0] * 3, np.float32), np.array([1.0] * 3, np.float32)
    extent = 2.0
    res = (args.res,) * 3
    lo, hi = mer.fields.padded_bbox(box_min, box_max, res)
    data = mer.fields.radial_rif(res, lo, hi) if args.field == "radial" else mer.fields.linear_rif(res, lo, hi)
    peak = 6555.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    g = torch.Generator(device=dev)
    g.manual_seed(20201201)
    n = args.rays
    p0 = (torch.rand(n, 3, device=dev, generator=g) * 2 - 1) * 0.999
    d0 = torch.randn(n, 3, device=dev, generator=g)
    d0 = d0 / d0.norm(dim=1, keepdim=True)
    for mode in args.modes.split(","):
        rif = mer.SplineDataSource(data=data, min=lo, max=hi, mode=mode)
        n0 = torch.empty(n, device=dev)
        check(lib.mer_rif_eval_device(rif.handle, 0, n, C.c_void_p(p0.data_ptr()), C.c_void_p(n0.data_ptr()), None, None))
        torch.cuda.synchronize()
        v0 = d0 * n0[:, None]
        for frac in [float(x) for x in args.fractions.split(",")]:
            h = frac * extent
            med = mer.HeterogeneousRefractiveMedium(dict(sigmaS=1.0, sigmaA=0.0, stepsize=h, strategy="single",
                                                         shape=("box", box_min, box_max))).addChild("rif", rif).configure()
            dist = torch.full((n,), 1.0 * extent, device=dev)
            nsteps = torch.zeros(n, dtype=torch.int32, device=dev)
            ok = torch.zeros(n, dtype=torch.uint8, device=dev)
            ds = torch.zeros(n, device=dev)
            times = []
            stream = torch.cuda.current_stream().cuda_stream
            for it in range(8):
                p, v = p0.clone(), v0.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                check(lib.mer_medium_trace_device(med.handle, n, C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr()),
                                                  C.c_void_p(dist.data_ptr()), C.c_void_p(ok.data_ptr()),
                                                  C.c_void_p(ds.data_ptr()), None, C.c_void_p(nsteps.data_ptr()),
                                                  C.c_void_p(stream)))
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            steps = int(nsteps.sum(dtype=torch.int64).item())
            bps = 512.0 if mode == "tricubic" else 256.0
            rate = steps / (ms * 1e-3)
            print(json.dumps({"config": "C4", "field": args.field, "res": args.res, "mode": mode, "h_over_extent": frac,
                              "h_over_pitch": h / float((hi[0] - lo[0]) / (args.res - 1)), "rays": n, "ray_steps": steps,
                              "ms": ms, "ray_steps_per_sec": rate, "alg_bytes_per_step": bps,
                              "achieved_GBps": rate * bps / 1e9, "hbm_peak_GBps": peak, "frac_of_hbm_roofline": rate * bps / 1e9 / peak,
                              "grid_bytes": int(np.prod(res)) * 16}), flush=True)


if __name__ == "__main__":
    main()
