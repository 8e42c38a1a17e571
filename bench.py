#!/usr/bin/env python
"""bench.py — throughput of the eikonal volumetric path tracer on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL film reduce

A "step" is one full render of the workload frame.  Default workload = BASELINE.json configs[1] (C2):
radial GRIN RIF 256^3 (tricubic B-spline) + 256^3 density grid (Woodcock tracking along the curved
ray), 512x512 at 256 spp, HG g = 0.9, albedo 0.9, step 1e-3 * extent.  Prints ONE JSON line.

  value        samples/s with the grids resident in HBM and the film left on the device (CUDA events, max over
               ranks).  N > 1: `--scaling weak` (default) = every rank renders 256 spp of its own sample indices
               (s = rank mod N); `--scaling strong` = the 256 spp are split over the ranks.  Films are summed with one
               NCCL reduce inside the timed region.
  e2e          the same metric through the C ABI from HOST buffers: upload of both raw grids from pinned memory + GPU
               prefilter + render + film read-back, all inside the timed region
  roofline     the dominant kernel (k_step, the leapfrog stepper), from counters of THIS run: coefficient blocks it
               gathered (256 B each), ray steps, and the CUDA-event time of its launches; against three measured
               ceilings — the texture unit's return path (the unit ncu shows saturated on L2-resident grids), the FP32
               pipe, HBM — with `bound` = the one it is closest to.  `traffic` = DRAM bytes per launch from the ncu
               capture of this configuration (profiles/r02_traffic.json).
  sub_results  c4_sweep: BASELINE configs[3], ray-steps/s of the batch stepper over five step sizes and both RIF modes,
               with the same ceilings; c5_strong: BASELINE configs[4] (1024^3 grids replicated per GPU) with a FIXED
               total sample count split over the N ranks (strong scaling: ms_per_frame at N / at 1 is the speed-up)
  cpu_baseline the oracle port timed on this host's cores on a bounded sample of the same workload
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
C5_STRONG_SPP = 64    # total samples per pixel of the strong-scaling sub-result (fixed work, split over the ranks)


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


def load_json(*parts, default=None):
    p = os.path.join(ROOT, *parts)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return default
    return default


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


# ------------------------------------------------------------------------------------------ ceilings
def ceilings():
    """the measured peaks the roofline block divides by, with where each comes from"""
    peaks = load_json("MEASURED_PEAKS.json")
    if peaks and "hbm_gbs" in peaks:
        hbm, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    else:
        hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    mb = load_json("profiles", "r02_microbench.json", default=None) or load_json("profiles", "r01_microbench.json", default={})
    fp32 = float(mb.get("fp32_fma_tflops", 72.5))
    tex = load_json("profiles", "r02_tex_peak.json", default={})
    return {"hbm_gbs": hbm, "hbm_source": hbm_src,
            "fp32_tflops": fp32, "fp32_source": "profiles/r02_microbench.json (FFMA microbenchmark, tools/microbench.cu)",
            "tex_gbs": float(tex.get("tex_return_gbs", 4820.0)),
            "tex_source": "profiles/r02_tex_peak.json (tools/microbench.cu k_tld4: incoherent 256-byte stencil gathers through the "
                          "texture unit from an L2-resident table, GB/s returned to registers)",
            "lsu_gather_gbs": float(tex.get("lsu_sector_gather_gbs", 8700.0)),
            "l2_read_gbs": float((mb.get("l2_read_gbs") or {}).get("64MiB", 15000.0))}


def step_costs(mode):
    """FP32 operations of one ray step of the step kernel, counted in its SASS (tools/sass_costs.py)"""
    c = load_json("profiles", "r02_step_costs.json", default={})
    return c.get(mode, {"flop_per_step": 520.0 if mode == "tricubic" else 120.0, "source": "default (profiles/r02_step_costs.json missing)"})


def roofline_block(kernel, mode, ray_steps, block_fetches, lookups, kernel_ms, launches, grid_bytes, traffic=None):
    """ray_steps / block_fetches / lookups / kernel_ms: totals of ONE rank over the timed region"""
    pk = ceilings()
    block_bytes = 256.0 if mode == "tricubic" else 128.0
    sc = step_costs(mode)
    secs = max(kernel_ms, 1e-9) * 1e-3
    gather_bytes = block_fetches * block_bytes + lookups * 32.0
    gather_gbs = gather_bytes / secs / 1e9
    tflops = ray_steps * sc["flop_per_step"] / secs / 1e12
    c = {
        "tex": {"achieved": gather_gbs, "peak": pk["tex_gbs"], "unit": "GB/s", "frac": gather_gbs / pk["tex_gbs"], "peak_source": pk["tex_source"],
                "what": "coefficient blocks returned through the texture unit (atlas layout), block_fetches x %d B" % int(block_bytes)},
        "fp32": {"achieved": tflops, "peak": pk["fp32_tflops"], "unit": "TFLOP/s", "frac": tflops / pk["fp32_tflops"], "peak_source": pk["fp32_source"],
                 "flop_per_ray_step": sc["flop_per_step"], "flop_source": sc.get("source", "profiles/r02_step_costs.json")},
        "hbm": {"achieved": gather_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gather_gbs / pk["hbm_gbs"], "peak_source": pk["hbm_source"],
                "what": "the same bytes against HBM: binding only when the coefficient table does not fit the 126 MB L2 (table: %d MiB)" % (grid_bytes >> 20)},
    }
    c["context"] = {"lsu_sector_gather_gbs": pk["lsu_gather_gbs"], "l2_read_gbs_64MiB": pk["l2_read_gbs"],
                    "note": "other measured ceilings of the same box (tools/microbench.cu): 8 x LDG.256 sector gathers from an L2-resident "
                            "table, and coalesced L2 reads; the stepper's gathers go through the texture unit (1x table) because the "
                            "8x sector table does not stay in L2"}
    if mode != "tricubic":
        c.pop("tex")  # the packed mode reads float4 nodes with LDG.128, not through the texture unit
    names = [k for k in c if k != "context" and (k != "hbm" or grid_bytes > 126e6)]
    bound = max(names, key=lambda k: c[k]["frac"])
    b = c[bound]
    return {"kernel": kernel, "bound": bound, "achieved": b["achieved"], "peak": b["peak"], "unit": b["unit"], "frac": b["frac"],
            "traffic": traffic, "peak_source": b["peak_source"], "ceilings": c,
            "ray_steps": ray_steps, "block_fetches": block_fetches, "block_fetches_per_ray_step": block_fetches / max(ray_steps, 1.0),
            "bytes_per_block": block_bytes, "launches": int(launches), "avg_launch_ms": kernel_ms / max(launches, 1),
            "alg_bytes_per_launch": gather_bytes / max(launches, 1),
            "note": "one 64-tap lookup per ray step (the reference's two are at the same point) and the 4x4x4 block is "
                    "re-gathered only when the ray changes cell: algorithmic bytes = blocks gathered x 256 B, counted on "
                    "the device in this run"}


# ------------------------------------------------------------------------------------------ CPU arm
ORACLE_FLAGS = "-O3 -fopenmp -march=x86-64-v3 -ffp-contract=off -fno-math-errno -fomit-frame-pointer (oracle/Makefile; the reference's " \
               "release flags are -O3 -march=nocona -msse2 -mfpmath=sse -funsafe-math-optimizations: BASELINE.md §3)"


def cpu_arm(w, budget_s=15.0, direct_connections=False):
    """the oracle port (reference algorithm on the host cores) on a bounded sample of the workload: same grids, same
    camera, a half-resolution film (>= 64 blocks of 32x32 so that every core has work) and as many spp as fit the budget"""
    os.environ["MER_B200_DEFER_LOAD"] = "1"  # this arm must not map the CUDA library (it only needs fields.py)
    from oracle.oracle import Oracle, volume_desc
    from common import oracle_medium_desc, oracle_render_desc
    nthreads = os.cpu_count() or 1  # explicit: torchrun exports OMP_NUM_THREADS=1
    orc = Oracle(np.float32)
    rif, lo, hi, den = make_fields(w)
    t0 = time.time()
    orif = orc.rif_create(volume_desc(w["rif_res"], lo, hi), rif)
    prefilter_s = time.time() - t0
    oden = orc.grid_create(volume_desc((w["den_res"],) * 3, BOX_MIN, BOX_MAX), den) if den is not None else None
    omed = orc.medium_create(oracle_medium_desc(medium_props(w), w["g"], has_density=den is not None), orif, oden)
    sw, sh = max(w["width"] // 2, 256), max(w["height"] // 2, 256)
    small = dict(w, width=sw, height=sh)
    nee = dict(direct_connections=direct_connections, props=medium_props(w))
    t0 = time.time()
    _, st = orc.render(omed, oracle_render_desc(scene_dict(small, 1), max_depth=w["max_depth"], rr_depth=5, **nee), nthreads=nthreads)
    cal = time.time() - t0
    spp = int(max(1, min(64, budget_s / max(cal, 1e-3))))
    t0 = time.time()
    _, st = orc.render(omed, oracle_render_desc(scene_dict(small, spp), max_depth=w["max_depth"], rr_depth=5, **nee), nthreads=nthreads)
    dt = time.time() - t0
    return dict(samples_per_s=st.samples / dt, steps_per_s=st.ray_steps / dt, cores=nthreads, threads_requested=nthreads,
                blocks=((sw + 31) // 32) * ((sh + 31) // 32),
                sample="%dx%d film (same camera, %d blocks of 32x32) at %d spp = %d samples, %.1f s on %d threads; prefilter %.1f s; flags: %s"
                       % (sw, sh, ((sw + 31) // 32) * ((sh + 31) // 32), spp, st.samples, dt, nthreads, prefilter_s, ORACLE_FLAGS),
                seconds=dt, samples=int(st.samples))


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
        "steps": K, "warmup": W, "ms_per_step": 1e3 * info["seconds"], "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname + ": " + w["desc"], "note": "CPU oracle port of the reference algorithm; each step = " + info["sample"]},
        "ray_steps_per_sec": float(np.mean(steps_rates)),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ sub-results
def c4_sweep(mer, dev, rays=1 << 24, res=256, fractions=(1e-2, 3e-3, 1e-3, 3e-4, 1e-4), modes=("tricubic", "trilinear_packed")):
    """BASELINE configs[3]: ray steps per second of the batch stepper (mer_medium_trace_device, kernel k_trace) on a 256^3
    radial RIF: 2^24 rays, uniform origins in the box, uniform directions, arc length 1.0 * extent, no scattering"""
    import torch
    from mitsubaer_b200._abi import check, lib
    extent = 2.0
    r3 = (res,) * 3
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, r3)
    data = mer.fields.radial_rif(r3, lo, hi)
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    p0 = (torch.rand(rays, 3, device=dev, generator=g) * 2 - 1) * 0.999
    d0 = torch.randn(rays, 3, device=dev, generator=g)
    d0 = d0 / d0.norm(dim=1, keepdim=True)
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    for mode in modes:
        rif = mer.SplineDataSource(data=data, min=lo, max=hi, mode=mode, device=dev.index)
        n0 = torch.empty(rays, device=dev)
        check(lib.mer_rif_eval_device(rif.handle, 0, rays, C.c_void_p(p0.data_ptr()), C.c_void_p(n0.data_ptr()), None, None))
        torch.cuda.synchronize()
        v0 = d0 * n0[:, None]
        for frac in fractions:
            h = frac * extent
            med = mer.HeterogeneousRefractiveMedium(dict(sigmaS=1.0, sigmaA=0.0, stepsize=h, strategy="single",
                                                         shape=("box", BOX_MIN, BOX_MAX))).addChild("rif", rif).configure()
            dist = torch.full((rays,), 1.0 * extent, device=dev)
            nsteps = torch.zeros(rays, dtype=torch.int32, device=dev)
            fetch = torch.zeros(1, dtype=torch.int64, device=dev)
            times = []
            reps = 2 if frac <= 3e-4 else 4  # 1 warm-up + the rest timed; the long runs take seconds each
            for it in range(reps):
                p, v = p0.clone(), v0.clone()
                fetch.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                check(lib.mer_medium_trace_counted_device(med.handle, rays, C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr()),
                                                          C.c_void_p(dist.data_ptr()), None, None, None, C.c_void_p(nsteps.data_ptr()),
                                                          C.c_void_p(fetch.data_ptr()), C.c_void_p(stream)))
                e1.record()
                torch.cuda.synchronize()
                if it >= 1:
                    times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            steps = float(nsteps.sum(dtype=torch.int64).item())
            rl = roofline_block("k_trace<%s>" % mode, mode, steps, float(fetch.item()), 0.0, ms, 1, int(np.prod(r3)) * (4 if mode == "tricubic" else 16))
            out.append({"mode": mode, "h_over_extent": frac, "h_over_pitch": h / float((hi[0] - lo[0]) / (res - 1)), "rays": rays,
                        "ray_steps": steps, "ms": ms, "ray_steps_per_sec": steps / (ms * 1e-3), "bound": rl["bound"], "frac": rl["frac"],
                        "block_fetches_per_ray_step": rl["block_fetches_per_ray_step"],
                        "frac_tex": rl["ceilings"].get("tex", {}).get("frac"), "frac_fp32": rl["ceilings"]["fp32"]["frac"],
                        "gather_GBps": rl["ceilings"]["hbm"]["achieved"]})
            del med
        del rif
    return out


def c5_strong(mer, dist, dev, local, rank, world, spp_total=C5_STRONG_SPP, steps=1, warmup=1):
    """BASELINE configs[4] with FIXED total work: 1024^3 RIF + density generated on the device of every rank, 2048x2048
    at `spp_total` samples per pixel split over the ranks (s = rank mod N), one NCCL reduce of the 80 MiB film"""
    import torch
    w = WORKLOADS["C5"]
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 40 * (1 << 30):
        return {"skipped": "needs 40 GiB of free device memory, %.1f GiB free" % (free / (1 << 30))}
    t0 = time.time()
    rif_t, lo, hi, den_t = make_fields(w, xp=torch, device=dev)
    rif = mer.SplineDataSource(data_ptr=rif_t.data_ptr(), res=w["rif_res"], min=lo, max=hi, device=local, mode="tricubic")
    del rif_t
    grid = mer.GridDataSource(data_ptr=den_t.data_ptr(), res=(w["den_res"],) * 3, min=BOX_MIN, max=BOX_MAX, device=local)
    del den_t
    med = mer.HeterogeneousRefractiveMedium(medium_props(w)).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=w["g"]))
    med = med.addChild("density", grid).configure()
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    integ = mer.EikonalVolPathIntegrator(maxDepth=w["max_depth"], rrDepth=5)
    scene = scene_dict(w, spp_total)
    film = torch.zeros(w["height"], w["width"], 5, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    stats = []

    def frame():
        film.zero_()
        st = integ.render_device(scene, med, film.data_ptr(), stream=stream, sample_begin=rank, sample_stride=world)
        if world > 1:
            dist.reduce(film, dst=0)
        return st

    for _ in range(warmup):  # warm-up on an eighth of the samples: allocations, instruction and TLB caches
        film.zero_()
        integ.render_device(scene_dict(w, max(spp_total // 8, world)), med, film.data_ptr(), stream=stream, sample_begin=rank, sample_stride=world)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        stats.append(frame())
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(sum(s[k] for s in stats)) for k in ("samples", "ray_steps", "block_fetches", "scatter_events", "null_collisions")],
                       device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item()) / steps
    samples, ray_steps = float(tot[0].item()) / steps, float(tot[1].item()) / steps
    s0 = stats[-1]
    rl = roofline_block("k_step<tricubic>", "tricubic", float(s0["ray_steps"]), float(s0["block_fetches"]),
                        float(s0["scatter_events"] + s0["null_collisions"]), float(s0["step_kernel_ms"]), s0["step_launches"],
                        int(np.prod(w["rif_res"])) * 4)
    out = {"workload": "C5: radial RIF 1024^3 + 1024^3 density, 2048x2048, %d spp in total split over %d GPU(s)" % (spp_total, world),
           "scaling": "strong", "n_gpus": world, "spp_total": spp_total, "ms_per_frame": ms, "samples_per_sec": samples / (ms * 1e-3),
           "ray_steps_per_sec": ray_steps / (ms * 1e-3), "setup_s_generate_prefilter": setup_s,
           "rank0": {"device_ms": s0["device_ms"], "step_kernel_ms": s0["step_kernel_ms"], "tail_ms": s0["tail_ms"],
                     "tail_share": s0["tail_ms"] / max(s0["device_ms"], 1e-9), "rounds": s0["passes"],
                     "step_lanes_per_sm": int(s0["step_lanes_per_sm"])},
           "roofline": {k: rl[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "block_fetches_per_ray_step")},
           "frac_hbm": rl["ceilings"]["hbm"]["frac"], "frac_tex": rl["ceilings"]["tex"]["frac"]}
    del med, rif, grid, film
    mer.lib.mer_trim_memory(local)
    return out


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
    if args.scaling == "strong":
        spp_total = spp          # fixed work: the ranks share the sample indices
        spp_rank = (spp - rank + world - 1) // world
    else:
        spp_total = spp * world  # weak scaling: every rank renders `spp` of the spp*N sample indices
        spp_rank = spp

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
    tot = torch.tensor([float(sum(s["samples"] for s in stats)), float(sum(s["ray_steps"] for s in stats)), float(launches)],
                       device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(tmax.item())
    samples, ray_steps, launches = [float(x) for x in tot.tolist()]
    value = samples / (ms * 1e-3)
    r0 = {k: float(sum(s[k] for s in stats)) for k in ("ray_steps", "block_fetches", "scatter_events", "null_collisions", "step_kernel_ms",
                                                        "step_launches", "device_ms", "tail_ms", "passes")}

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

    # ---- sub-results on the configurations the targets are quoted on (all ranks take part in c5_strong)
    sub = {}
    del med, rif, grid
    if not args.no_sub and wname == "C2":
        mer.lib.mer_trim_memory(local)
        torch.cuda.empty_cache()
        try:
            sub["c5_strong"] = c5_strong(mer, dist, dev, local, rank, world)
        except Exception as e:  # noqa: BLE001 -- a sub-result must not take the headline down
            sub["c5_strong"] = {"error": repr(e)}
        if world == 1:
            try:
                sub["c4_sweep"] = c4_sweep(mer, dev)
            except Exception as e:  # noqa: BLE001
                sub["c4_sweep"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_step), rank 0's counters over the timed region
    grid_bytes = int(np.prod(w["rif_res"])) * (4 if args.mode == "tricubic" else 16)
    traffic = (load_json("profiles", "r02_traffic.json", default={}) or {}).get(wname + ":" + args.mode)
    roofline = roofline_block("k_step<%s>" % args.mode, args.mode, r0["ray_steps"], r0["block_fetches"],
                              r0["scatter_events"] + r0["null_collisions"], r0["step_kernel_ms"], r0["step_launches"], grid_bytes, traffic)
    roofline["step_kernel_share_of_device_time"] = r0["step_kernel_ms"] / max(r0["device_ms"], 1e-9)
    roofline["drain_tail_share"] = r0["tail_ms"] / max(r0["device_ms"], 1e-9)
    roofline["step_lanes_per_sm"] = int(stats[-1]["step_lanes_per_sm"])  # 512, or 192 when the table exceeds the L2 and the in-run comparison chose it

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample, rank 0 / N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_arm(w, budget_s=15.0, direct_connections=args.direct_connections)
        cpu = {"value": c["samples_per_s"], "unit": "samples/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "ray_steps_per_sec": c["steps_per_s"]}

    out = {
        "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wname + ": " + w["desc"], "rif_mode": args.mode, "spp_per_gpu": spp_rank, "spp_total": spp_total,
                   "stepsize": props["stepsize"], "seed": SEED, "film": "%dx%d box filter" % (w["width"], w["height"]),
                   "l2": "between steps the film is zeroed and the 310 MB path pool is rewritten (larger than the 126 MB L2); the 64 MiB "
                         "coefficient atlas is meant to stay L2-resident, that is the design" if grid_bytes <= 126e6 else
                         "inputs larger than L2 (coefficient table %d MiB + density)" % (grid_bytes >> 20),
                   "parallelism": "sample-index sharding, NCCL film reduce" if world > 1 else "single GPU",
                   "setup_s_upload_prefilter": setup_s},
        "ray_steps_per_sec": ray_steps / (ms * 1e-3),
        "ray_steps_per_sample": ray_steps / max(samples, 1),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "sub_results": sub,
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank renders the workload's spp (spp*N in total); strong = the spp are split over the ranks")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (per GPU when weak, in total when strong)")
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--steps-per-pass", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip sub_results (the C4 sweep and the C5 strong-scaling frame)")
    ap.add_argument("--direct-connections", action="store_true",
                    help="next-event estimation along curved connections (SURVEY 8f-1; homogeneous workloads: C1, C3)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        os.environ["MER_B200_DEFER_LOAD"] = "1"
        run_reference(args, w, args.workload)
    else:
        run_gpu(args, w, args.workload)


if __name__ == "__main__":
    main()
