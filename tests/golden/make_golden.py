"""Generates tests/golden/*.npz from the REFERENCE's own code (run in the build container,
where /root/reference exists):

  spline_ref_f32.npz / spline_ref_f64.npz
      include/mitsuba/core/basisspline.h compiled verbatim (oracle/ref_spline.cpp ->
      oracle/_ref/libmer_refspline_{f,d}.so): prefiltered coefficients of a seeded random
      20x30x23 grid (the shape of mfiles/Test.m) and value / gradient / Hessian at seeded points.

  phase_ref.npz
      src/phase/hg.cpp (HGPhaseFunction::sample / eval), include/mitsuba/core/frame.h and, from src/libcore/util.cpp,
      coordinateSystem() and fresnelDielectricExt(), compiled verbatim (oracle/ref_phase.cpp ->
      oracle/_ref/libmer_refphase.so): sampled directions, pdfs, frames and Fresnel terms at seeded inputs,
      g in {0.9, -0.3} (data/tests/test_phase.xml:12-21), 0.5, 0 and 1e-5 (the isotropic branch).

  trace_ref.npz
      src/medium/heterogeneousrefractive.cpp (er_step, trace, traceTillBoundary, insideShape = hackForSphere) and the
      SplineDataSource wrappers of src/volume/splinevolume.cpp, compiled verbatim (oracle/ref_trace.cpp ->
      oracle/_ref/libmer_reftrace.so): end states of seeded rays in the reference's hard-coded sphere for four fields
      (the scenes of tests/common.py:ref_sphere_scene; the inputs are regenerated from the seeds, only outputs are stored).

Usage:  make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import RefFilm, RefGrid, RefPhase, RefSpline, RefTrace  # noqa: E402


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


def phase():
    rng = np.random.default_rng(20201202)
    n = 4096
    wi = rng.normal(size=(n, 3))
    wi = (wi / np.linalg.norm(wi, axis=1, keepdims=True)).astype(np.float32)
    wi[:3] = np.eye(3, dtype=np.float32)  # axis-aligned frames (the |a.x| > |a.y| branch point)
    wo_in = rng.normal(size=(n, 3))
    wo_in = (wo_in / np.linalg.norm(wo_in, axis=1, keepdims=True)).astype(np.float32)
    xi = rng.random((n, 2)).astype(np.float32)
    xi[:4] = [[0, 0], [0.99999994, 0.99999994], [0.5, 0], [0, 0.5]]
    ref = RefPhase()
    out = dict(wi=wi, xi=xi, wo_in=wo_in, g=np.array([0.9, -0.3, 0.5, 0.0, 1e-5], np.float32))
    for k, g in enumerate(out["g"]):
        wo, pdf = ref.hg_sample(float(g), wi, xi)
        out["wo_%d" % k], out["pdf_%d" % k] = wo, pdf
        out["eval_%d" % k] = ref.hg_eval(float(g), wi, wo_in)
    out["frame_s"], out["frame_t"] = ref.coordinate_system(wi)
    cos_i = (rng.random(n) * 2 - 1).astype(np.float32)
    eta = (1 + rng.random(n)).astype(np.float32)
    cos_i[:4], eta[:4] = [1.0, -1.0, 0.0, 0.3], [1.5, 1.5, 1.5, 1.0]
    out["cos_i"], out["eta"] = cos_i, eta
    out["fresnel"], out["cos_t"] = ref.fresnel_dielectric_ext(cos_i, eta)
    # MaxExpDist (src/medium/maxexp.h), the "maximum" free-flight strategy
    out["maxexp_sigma_t"] = np.array([[2.5, 3.25, 0.5], [4.0, 3.5, 3.0]], np.float32)
    out["maxexp_u"] = rng.random(n).astype(np.float32)
    out["maxexp_t"] = (rng.random(n) * 6).astype(np.float32)
    for k, st in enumerate(out["maxexp_sigma_t"]):
        out["maxexp_sample_t_%d" % k], out["maxexp_sample_pdf_%d" % k] = ref.maxexp(st, 0, out["maxexp_u"])
        out["maxexp_pdf_%d" % k] = ref.maxexp(st, 1, out["maxexp_t"])
        out["maxexp_cdf_%d" % k] = ref.maxexp(st, 2, out["maxexp_t"])
    # f-3: HSmoothDielectric::sample (src/bsdfs/hdielectric.cpp:244-300), radiance and importance mode
    dd = rng.normal(size=(n, 3))
    dd = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
    nn = rng.normal(size=(n, 3))
    nn = (nn / np.linalg.norm(nn, axis=1, keepdims=True)).astype(np.float32)
    out["hd_d"], out["hd_n"], out["hd_eta"], out["hd_u"] = dd, nn, (1.05 + 0.6 * rng.random(n)).astype(np.float32), rng.random(n).astype(np.float32)
    for mode in (0, 1):
        o, w, es, tr = ref.hdielectric_sample(dd, nn, out["hd_eta"], out["hd_u"], bool(mode))
        out["hd_out_%d" % mode], out["hd_weight_%d" % mode], out["hd_etascale_%d" % mode], out["hd_transmitted_%d" % mode] = o, w, es, tr
    np.savez_compressed(os.path.join(HERE, "phase_ref.npz"), **out)
    print("wrote phase_ref.npz")


TRACE_CASES = (("linear", 2e-3), ("radial", 5e-3), ("sd", 1e-3), ("smooth", 3.3e-3))


def trace():
    sys.path.insert(0, os.path.dirname(HERE))
    from common import ref_sphere_scene
    out = {}
    for kind, h in TRACE_CASES:
        data, lo, hi, p0, d0, dist = ref_sphere_scene(kind, n_rays=1024)
        ref = RefTrace(data, lo, hi, h)
        n0, g0 = ref.value_gradient(p0)
        v0 = (d0 * n0[:, None]).astype(np.float32)
        p, v, ds, opl, ok = ref.trace(p0, v0, dist)
        tp, tv, tds, topl = ref.trace_till_boundary(p0, v0)
        out.update({kind + "_h": np.float32(h), kind + "_n0": n0, kind + "_g0": g0, kind + "_p": p, kind + "_v": v, kind + "_dist_surf": ds,
                    kind + "_opl": opl, kind + "_success": ok, kind + "_tb_p": tp, kind + "_tb_v": tv, kind + "_tb_dist_surf": tds,
                    kind + "_tb_opl": topl, kind + "_data_sum": np.float64(data.astype(np.float64).sum()), kind + "_p0_sum": np.float64(p0.astype(np.float64).sum())})
    # Medium::sampleDistance (:402-568) and evalTransmittance (:393-400); the media are resolved by the restated oracle (weight,
    # sampling density), which test_oracle_cpu.py pins bit for bit against this very library
    from test_oracle_cpu import SAMPLE_DISTANCE_CASES, _sample_distance_scene
    from common import oracle_medium_desc
    from oracle.oracle import Oracle, volume_desc
    orc = Oracle(np.float32)
    for strategy, aggressive in SAMPLE_DISTANCE_CASES:
        props, data, lo, hi, sdf, ro, rd, mint, xi = _sample_distance_scene(orc, strategy, aggressive, n_rays=1024)
        omed = orc.medium_create(oracle_medium_desc(props), orc.rif_create(volume_desc(data.shape[::-1], lo, hi), data))
        w, sd = orc.medium_resolved(omed)
        ref = RefTrace(data, lo, hi, props["stepsize"]).configure(props["sigmaA"], props["sigmaS"], strategy, sd, w, sdf, lo, hi, aggressive)
        r = ref.sample_distance(ro, rd, mint, xi)
        tag = "sd_%s_%d_" % (strategy, int(aggressive))
        out.update({tag + k: v for k, v in r.items()})
        out[tag + "weight"], out[tag + "density"] = np.float32(w), np.float32(sd)
    # a25 minus the solver: er_derivativestep, computefdfBDPT (residual + Jacobian)
    from test_oracle_cpu import CONNECTION_KINDS, _connection_scene
    for kind in CONNECTION_KINDS:
        props, data, lo, hi, sdf, p1, d0, p2, w = _connection_scene(kind, n=512)
        ref = RefTrace(data, lo, hi, props["stepsize"]).configure((0.4,) * 3, (3.6,) * 3, "single", 4.0, 0.9, sdf, lo, hi, False).set_connection(3, 1e-6)
        n0, _ = ref.value_gradient(p1)
        v0 = (d0 * n0[:, None]).astype(np.float32)
        dt = ref.derivative_trace(p1, v0, 40)
        out["conn_%s_n0" % kind] = n0
        out.update({"conn_%s_dt_%s" % (kind, k): v for k, v in dt.items()})
        for sensor in (0, 1):
            B = ref.connection_residual(p1, p2, w, is_sensor=bool(sensor))
            out["conn_%s_error_%d" % (kind, sensor)], out["conn_%s_derror_%d" % (kind, sensor)] = B["error"], B["derror"]
    # a18: GridDataSource::lookupFloat
    from test_oracle_cpu import _grid_scene
    res, data, lo, hi, p = _grid_scene()
    out["grid_lookup"] = RefGrid(data, lo, hi).lookup(p)
    # a23: filter tables and ImageBlock::put
    from test_oracle_cpu import _film_scene
    W, H, pos, values = _film_scene()
    rf = RefFilm()
    for ftype in (0, 1):
        v, r, s_, b = rf.filter_table(ftype)
        out["film_table_%d" % ftype], out["film_radius_%d" % ftype] = v, np.float32(r)
        out["film_put_%d" % ftype] = rf.film_put(ftype, W, H, pos, values)[0]
    # a22: the pinhole sensor's rays
    from test_oracle_cpu import CAMERAS, _camera_samples
    for k, cam in enumerate(CAMERAS):
        out["camera_d_%d" % k] = rf.camera_rays(cam["origin"], cam["target"], cam["up"], cam["fov"], cam["width"], cam["height"], _camera_samples(cam))[1]
    np.savez_compressed(os.path.join(HERE, "trace_ref.npz"), **out)
    print("wrote trace_ref.npz")


def grid_spectrum():
    """GridDataSource::lookupSpectrum (gridvolume.cpp:386-463, float3 :293-329) compiled verbatim, float32 and uint8 payloads"""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_oracle_cpu import _grid_spectrum_scene
    res, rgb, u8, lo, hi, p = _grid_spectrum_scene()
    np.savez_compressed(os.path.join(HERE, "grid_spectrum_ref.npz"), f32=RefGrid(rgb, lo, hi).lookup_spectrum(p),
                        u8=RefGrid(u8, lo, hi).lookup_spectrum(p))
    print("wrote grid_spectrum_ref.npz")


if __name__ == "__main__":
    if "--grid-spectrum-only" in sys.argv:  # the other files are not rewritten
        grid_spectrum()
        sys.exit(0)
    main()
    phase()
    trace()
    grid_spectrum()
