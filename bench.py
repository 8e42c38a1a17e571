#!/usr/bin/env python
"""bench.py — throughput of the eikonal volumetric path tracer on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL film reduce

A "step" is one full render of the workload frame.  Default workload = BASELINE.json configs[1] (C2):
radial GRIN RIF 256^3 (tricubic B-spline) + 256^3 density grid (Woodcock tracking along the curved
ray), 512x512 at 256 spp, HG g = 0.9, albedo 0.9, step 1e-3 * extent.  Prints ONE JSON line.

  value      samples/s with the grids resident in HBM and the film left on the device (CUDA events,
             max over ranks); multi-GPU = weak scaling: every rank renders 256 spp of its own sample
             indices (s = rank mod N), films are summed with one NCCL reduce inside the timed region
  e2e        the same metric through the C ABI from HOST buffers: upload of both raw grids from pinned
             memory + GPU prefilter + render + film read-back, all inside the timed region
  roofline   dominant kernel k_render_pass: algorithmic bytes (512 B per tricubic ray step — two 64-tap
             spline evaluations of 4 B coefficients, SURVEY §8d — plus 32 B per Woodcock density lookup)
             over the kernel's CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the oracle port timed on this host's cores on a bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20201201
BOX_MIN = np.array([-1.0, -1.0, -1.0], np.float32)
BOX_MAX = np.array([1.0, 1.0, 1.0], np.float32)

WORKLOADS = {
    # name: RIF kind/res, density res (0 = homogeneous), film, spp, medium
    "C1": dict(desc="linear RIF 226x226x51, homogeneous sigma_t=4 albedo .9, HG g=.9, 256x256 @ 64 spp",
               rif="linear", rif_res=(226, 226, 51), den_res=0, width=256, height=256, spp=64,
               medium=dict(sigmaT=4.0, albedo=0.9, strategy="single"), g=0.9, max_depth=64),
    "C2": dict(desc="radial GRIN RIF 256^3 + 256^3 density grid (scale 8, albedo .9), HG g=.9, 512x512 @ 256 spp",
               rif="radial", rif_res=(256, 256, 256), den_res=256, width=512, height=512, spp=256,
               medium=dict(albedo=0.9, densityScale=8.0), g=0.9, max_depth=64),
    "C3": dict(desc="SD-derived RIF 512^3, homogeneous sigma_t=16 albedo .999, HG g=.9, 1024x1024 @ 1024 spp",
               rif="sd", rif_res=(512, 512, 512), den_res=0, width=1024, height=1024, spp=1024,
               medium=dict(sigmaT=16.0, albedo=0.999, strategy="single"), g=0.9, max_depth=512),
    "C5": dict(desc="radial RIF 1024^3 + 1024^3 density grid, 2048x2048 @ 4096 spp",
               rif="radial", rif_res=(1024, 1024, 1024), den_res=1024, width=2048, height=2048, spp=4096,
               medium=dict(albedo=0.9, densityScale=8.0), g=0.9, max_depth=64),
    # small case for smoke-testing the harness itself
    "tiny": dict(desc="radial RIF 64^3 + 32^3 density, 64x64 @ 8 spp", rif="radial", rif_res=(64, 64, 64), den_res=32,
                 width=64, height=64, spp=8, medium=dict(albedo=0.9, densityScale=8.0), g=0.9, max_depth=64),
}
STEP_FRACTION = 1e-3  # step = 1e-3 * box extent


def make_fields(w, xp=np, **kw):
    from mitsubaer_b200 import fields
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, w["rif_res"])
    if w["rif"] == "linear":
        rif = fields.linear_rif(w["rif_res"], lo, hi, xp=xp, **kw)
    elif w["rif"] == "radial":
        rif = fields.radial_rif(w["rif_res"], lo, hi, xp=xp, **kw)
    else:
        rif = fields.rif_from_sd(fields.sphere_sdf(w["rif_res"], lo, hi, radius=0.8, xp=xp, **kw), xp=xp)
    den = None
    if w["den_res"]:
        den = fields.sine_density((w["den_res"],) * 3, BOX_MIN, BOX_MAX, xp=xp, **kw)
    return rif, lo, hi, den


def medium_props(w):
    p = dict(w["medium"])
    p["stepsize"] = STEP_FRACTION * float(BOX_MAX[0] - BOX_MIN[0])
    p["shape"] = ("box", BOX_MIN, BOX_MAX)
    return p


def scene_dict(w, spp_total):
    return dict(width=w["width"], height=w["height"], sampleCount=spp_total, seed=SEED, origin=(0.0, 0.0, -4.0),
                target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=40.0, rfilter="box", envRadiance=1.0,
                quad=dict(origin=(-0.5, 1.5, -0.5), u=(1.0, 0.0, 0.0), v=(0.0, 0.0, 1.0), radiance=(8.0, 8.0, 8.0)))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed power range
        pw = np.array(power)
        load = pw >= (pw.min() + 0.5 * (pw.max() - pw.min()))
        return {"sm_mhz": float(np.median(np.array(sm)[load])), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(pw.max()), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(w, budget_s=15.0, direct_connections=False):
    """the oracle port (reference algorithm on the host cores) on a bounded sample of the workload:
    same grids, same camera, a centred 1/4-resolution film and as many spp as fit the time budget"""
    from oracle.oracle import Oracle, volume_desc
    from common import oracle_medium_desc, oracle_render_desc
    orc = Oracle(np.float32)
    rif, lo, hi, den = make_fields(w)
    t0 = time.time()
    orif = orc.rif_create(volume_desc(w["rif_res"], lo, hi), rif)
    prefilter_s = time.time() - t0
    oden = orc.grid_create(volume_desc((w["den_res"],) * 3, BOX_MIN, BOX_MAX), den) if den is not None else None
    omed = orc.medium_create(oracle_medium_desc(medium_props(w), w["g"], has_density=den is not None), orif, oden)
    sw, sh = max(w["width"] // 4, 16), max(w["height"] // 4, 16)
    small = dict(w, width=sw, height=sh)
    scene = scene_dict(small, 1)
    nee = dict(direct_connections=direct_connections, props=medium_props(w))
    t0 = time.time()
    _, st = orc.render(omed, oracle_render_desc(scene, max_depth=w["max_depth"], rr_depth=5, **nee))
    cal = time.time() - t0
    spp = int(max(1, min(64, budget_s / max(cal, 1e-3))))
    scene = scene_dict(small, spp)
    t0 = time.time()
    _, st = orc.render(omed, oracle_render_desc(scene, max_depth=w["max_depth"], rr_depth=5, **nee))
    dt = time.time() - t0
    return dict(samples_per_s=st.samples / dt, steps_per_s=st.ray_steps / dt, cores=orc.num_threads(),
                sample="%dx%d film (same camera) at %d spp = %d samples, %.1f s; prefilter %.1f s"
                       % (sw, sh, spp, st.samples, dt, prefilter_s), seconds=dt, samples=int(st.samples))


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    rates, steps_rates, info = [], [], None
    budget = max(3.0, min(15.0, 120.0 / max(K + W, 1)))
    for i in range(W + K):
        r = cpu_arm(w, budget_s=budget, direct_connections=args.direct_connections)
        if i >= W:
            rates.append(r["samples_per_s"])
            steps_rates.append(r["steps_per_s"])
            info = r
    v = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": "samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname + ": " + w["desc"], "note": "CPU oracle port of the reference algorithm; each step = " + info["sample"]},
        "ray_steps_per_sec": float(np.mean(steps_rates)),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, w, wname):
    import torch
    import torch.distributed as dist

    import mitsubaer_b200 as mer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if mer.device_count() == 0:
        raise SystemExit("bench.py: no sm_100 GPU visible and there is no CPU fallback (use --impl reference)")
    K, W = args.steps, args.warmup
    spp = args.spp or w["spp"]
    spp_total = spp * world  # weak scaling: every rank renders `spp` of the spp*N sample indices

    # ---- inputs: host copies in pinned memory (e2e) and the resident handles (value)
    if max(w["rif_res"]) >= 512:
        rif_t, lo, hi, den_t = make_fields(w, xp=torch, device=dev)  # big grids are generated in HBM
        rif_host = den_host = None
    else:
        rif_np, lo, hi, den_np = make_fields(w)
        rif_host = torch.from_numpy(rif_np).pin_memory()
        den_host = torch.from_numpy(den_np).pin_memory() if den_np is not None else None
        rif_t = rif_host.to(dev)
        den_t = den_host.to(dev) if den_host is not None else None
    props = medium_props(w)

    def build_device():
        rif = mer.SplineDataSource(data_ptr=rif_t.data_ptr(), res=w["rif_res"], min=lo, max=hi, device=local, mode=args.mode)
        med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=w["g"]))
        grid = None
        if den_t is not None:
            grid = mer.GridDataSource(data_ptr=den_t.data_ptr(), res=(w["den_res"],) * 3, min=BOX_MIN, max=BOX_MAX, device=local)
            med.addChild("density", grid)
        return med.configure(), rif, grid

    def build_host():
        rif = mer.SplineDataSource(data=rif_host.numpy(), res=w["rif_res"], min=lo, max=hi, device=local, mode=args.mode)
        med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=w["g"]))
        grid = None
        if den_host is not None:
            grid = mer.GridDataSource(data=den_host.numpy(), res=(w["den_res"],) * 3, min=BOX_MIN, max=BOX_MAX, device=local)
            med.addChild("density", grid)
        return med.configure(), rif, grid

    t0 = time.time()
    med, rif, grid = build_device()
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    if den_t is not None:
        del den_t  # the handle keeps its own copy
    integ = mer.EikonalVolPathIntegrator(maxDepth=w["max_depth"], rrDepth=5, poolPaths=args.pool, stepsPerPass=args.steps_per_pass,
                                         directConnections=args.direct_connections)
    scene = scene_dict(w, spp_total)
    film = torch.zeros(w["height"], w["width"], 5, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        st = integ.render_device(scene, med, film.data_ptr(), stream=stream, sample_begin=rank, sample_stride=world)
        if world > 1:
            dist.reduce(film, dst=0)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        film.zero_()
        step_resident()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = mer.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    ev0.record()
    for _ in range(K):
        film.zero_()
        stats.append(step_resident())
    ev1.record()
    barrier()
    launches = mer.kernel_launch_count() - launches0
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(sum(s["samples"] for s in stats)), float(sum(s["ray_steps"] for s in stats)),
                        float(sum(s["scatter_events"] + s["null_collisions"] for s in stats)), float(launches)],
                       device=dev, dtype=torch.float64)
    kern_ms = torch.tensor([float(sum(s["device_ms"] for s in stats))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    samples, ray_steps, lookups, launches = [float(x) for x in tot.tolist()]
    value = samples / (ms * 1e-3)

    # ---- e2e: host buffers in, host film out, through the public plugin API / C ABI
    e2e = None
    if rif_host is not None:
        film_host = torch.zeros(w["height"], w["width"], 5).pin_memory()

        def step_e2e():
            m2, r2, g2 = build_host()  # H2D of the raw grids + GPU prefilter
            film.zero_()
            integ.render_device(scene, m2, film.data_ptr(), stream=stream, sample_begin=rank, sample_stride=world)
            if world > 1:
                dist.reduce(film, dst=0)
            if rank == 0:
                film_host.copy_(film, non_blocking=False)  # D2H read-back of the result
            del m2, r2, g2

        step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(K):
            step_e2e()
        e1.record()
        barrier()
        wall_ms = (time.time() - t0) * 1e3
        ems = torch.tensor([max(e0.elapsed_time(e1), wall_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = rif_host.numel() * 4 + (den_host.numel() * 4 if den_host is not None else 0)
        e2e = {"value": samples / (float(ems.item()) * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(film_host.numel() * 4), "ms_per_step": float(ems.item()) / K}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_render_pass), per launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_step = 512.0 if args.mode == "tricubic" else 256.0
    k_ms = float(kern_ms.item())
    # per-rank figures: algorithmic bytes of this rank's launches / this rank's kernel time
    alg_bytes = (ray_steps * bytes_per_step + lookups * 32.0) / world
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    render_launches = sum(s["kernel_launches"] for s in stats)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(wname + ":" + args.mode)
    roofline = {"bound": "hbm", "kernel": "k_render_pass<%s>" % args.mode, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_ray_step": bytes_per_step, "launches": int(render_launches),
                "avg_launch_ms": k_ms / max(render_launches, 1),
                "alg_bytes_per_launch": alg_bytes / max(render_launches, 1),
                "note": "algorithmic bytes exceed HBM peak when the stencil is served from L1/L2 (grid "
                        "re-use between consecutive steps of a ray); frac > 1 is cache reuse, not skipped work"}

    # secondary ceilings measured with tools/microbench.cu on this pool (FP32 issue rate at ~1.0 kFLOP per
    # tricubic step, 130 FLOP packed; L2 read bandwidth for tables that fit the L2)
    mpath = os.path.join(ROOT, "profiles", "r01_microbench.json")
    if os.path.exists(mpath):
        mb = json.load(open(mpath))
        flop_per_step = 1000.0 if args.mode == "tricubic" else 160.0
        tf = ray_steps / world * flop_per_step / (k_ms * 1e-3) / 1e12
        roofline["fp32"] = {"achieved_tflops": tf, "peak_tflops": mb["fp32_fma_tflops"], "frac": tf / mb["fp32_fma_tflops"],
                            "flop_per_ray_step": flop_per_step, "peak_source": "profiles/r01_microbench.json (FFMA microbenchmark)"}
        roofline["l2_read_gbs_measured"] = mb["l2_read_gbs"]

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample, rank 0 / N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_arm(w, budget_s=15.0, direct_connections=args.direct_connections)
        cpu = {"value": c["samples_per_s"], "unit": "samples/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "ray_steps_per_sec": c["steps_per_s"]}

    out = {
        "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wname + ": " + w["desc"], "rif_mode": args.mode, "spp_per_gpu": spp, "spp_total": spp_total,
                   "stepsize": props["stepsize"], "seed": SEED, "film": "%dx%d box filter" % (w["width"], w["height"]),
                   "l2": "inputs larger than L2 (coeff8 %d MiB + density)" % (int(np.prod(w["rif_res"])) * 32 >> 20)
                   if int(np.prod(w["rif_res"])) * 32 > 126e6 else "grids fit L2; film + path pool rewritten between steps",
                   "parallelism": "sample-index sharding, NCCL film reduce" if world > 1 else "single GPU",
                   "setup_s_upload_prefilter": setup_s},
        "ray_steps_per_sec": ray_steps / (ms * 1e-3),
        "ray_steps_per_sample": ray_steps / max(samples, 1),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
    }
    if args.direct_connections:  # rank-0 figures of the solver kernel (k_nee)
        conn = float(sum(s["connections"] for s in stats))
        out["direct_connections"] = {"connections_per_sec": conn * world / (ms * 1e-3),
                                     "failed_fraction": float(sum(s["connections_failed"] for s in stats)) / max(conn, 1.0),
                                     "hessian_steps_per_sec": float(sum(s["connection_steps"] for s in stats)) * world / (ms * 1e-3),
                                     "hessian_steps_per_connection": float(sum(s["connection_steps"] for s in stats)) / max(conn, 1.0)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MER_WORKLOAD", "C2"), choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="tricubic", choices=["tricubic", "trilinear_packed"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU")
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--steps-per-pass", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--direct-connections", action="store_true",
                    help="next-event estimation along curved connections (SURVEY 8f-1; homogeneous workloads: C1, C3)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_gpu(args, w, args.workload)


if __name__ == "__main__":
    main()
