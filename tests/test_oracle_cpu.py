"""CPU tests (no GPU): the oracle against the golden vectors generated from the reference's own
basisspline.h, against the verbatim build in oracle/_ref when present, and its own invariants."""
import os

import numpy as np
import pytest

from common import (BOX_MAX, BOX_MIN, make_field, medium_props, oracle_medium_desc, oracle_render_desc,
                    random_directions, random_points_in_box, scene_dict)
from oracle.oracle import Oracle, RefFilm, RefGrid, RefHeterogeneousMedium, RefPhase, RefSpline, RefTrace, volume_desc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
def test_oracle_spline_bit_exact_vs_reference_golden(tag, dtype):
    g = np.load(os.path.join(GOLDEN, "spline_ref_%s.npz" % tag))
    orc = Oracle(dtype)
    h = orc.rif_create(volume_desc(g["res"], g["bbox_min"], g["bbox_max"]), g["data"])
    assert np.array_equal(orc.rif_coefficients(h, g["data"].size), g["coeff"])
    f, grad, H = orc.rif_eval_hessian(h, g["points"])
    assert np.array_equal(f, g["value"]) and np.array_equal(grad, g["gradient"]) and np.array_equal(H, g["hessian"])
    for what in (0, 1, 2):
        f2, g2 = orc.rif_eval(h, g["points"], what)
        if what != 1:
            assert np.array_equal(f2, g["value"])
        if what != 0:
            assert np.array_equal(g2, g["gradient"])
    orc.rif_destroy(h)


@pytest.mark.skipif(not RefSpline.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_spline_bit_exact_vs_verbatim_reference(dtype):
    rng = np.random.default_rng(5)
    res = (17, 12, 21)
    data = (1 + rng.random((res[2], res[1], res[0]))).astype(np.float32)
    lo, hi = np.array([0, -1, 2], np.float32), np.array([3, 1, 2.5], np.float32)
    ref = RefSpline(dtype).build(data, res, lo, hi)
    orc = Oracle(dtype)
    h = orc.rif_create(volume_desc(res, lo, hi), data)
    assert np.array_equal(ref.coefficients(), orc.rif_coefficients(h, data.size))
    pitch = (hi - lo) / (np.array(res) - 1)
    p = (lo + 2.1 * pitch + rng.random((3000, 3)) * (hi - lo - 4.2 * pitch)).astype(dtype)
    fr, gr, Hr = ref.eval_hessian(p)
    fo, go, Ho = orc.rif_eval_hessian(h, p)
    assert np.array_equal(fr, fo) and np.array_equal(gr, go) and np.array_equal(Hr, Ho)
    orc.rif_destroy(h)


def test_oracle_phase_bit_exact_vs_reference_golden():
    """a15-a17 + fresnelDielectricExt: the restatement against vectors generated from src/phase/hg.cpp, frame.h and
    util.cpp compiled verbatim (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLDEN, "phase_ref.npz"))
    orc = Oracle(np.float32)
    for k, gg in enumerate(g["g"]):
        wo, pdf = orc.hg_sample(float(gg), g["wi"], g["xi"])
        assert np.array_equal(wo, g["wo_%d" % k]) and np.array_equal(pdf, g["pdf_%d" % k])
        assert np.array_equal(orc.hg_eval(float(gg), g["wi"], g["wo_in"]), g["eval_%d" % k])
    s, t = orc.coordinate_system(g["wi"])
    assert np.array_equal(s, g["frame_s"]) and np.array_equal(t, g["frame_t"])
    F, ct = orc.fresnel_dielectric_ext(g["cos_i"], g["eta"])
    assert np.array_equal(F, g["fresnel"]) and np.array_equal(ct, g["cos_t"])
    for k, st in enumerate(g["maxexp_sigma_t"]):  # MaxExpDist, src/medium/maxexp.h
        t, pdf = orc.maxexp(st, 0, g["maxexp_u"])
        assert np.array_equal(t, g["maxexp_sample_t_%d" % k]) and np.array_equal(pdf, g["maxexp_sample_pdf_%d" % k])
        assert np.array_equal(orc.maxexp(st, 1, g["maxexp_t"]), g["maxexp_pdf_%d" % k])
        assert np.array_equal(orc.maxexp(st, 2, g["maxexp_t"]), g["maxexp_cdf_%d" % k])


@pytest.mark.skipif(not RefPhase.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_phase_bit_exact_vs_verbatim_reference():
    rng = np.random.default_rng(11)
    n = 200000
    wi = rng.normal(size=(n, 3))
    wi = (wi / np.linalg.norm(wi, axis=1, keepdims=True)).astype(np.float32)
    wo = rng.normal(size=(n, 3))
    wo = (wo / np.linalg.norm(wo, axis=1, keepdims=True)).astype(np.float32)
    xi = rng.random((n, 2)).astype(np.float32)
    ref, orc = RefPhase(), Oracle(np.float32)
    for gg in (0.9, -0.3, 0.99, 0.0, 5e-5):
        a, pa = orc.hg_sample(gg, wi, xi)
        b, pb = ref.hg_sample(gg, wi, xi)
        assert np.array_equal(a, b) and np.array_equal(pa, pb)
        assert np.array_equal(orc.hg_eval(gg, wi, wo), ref.hg_eval(gg, wi, wo))
    assert all(np.array_equal(x, y) for x, y in zip(orc.coordinate_system(wi), ref.coordinate_system(wi)))
    cos_i = (rng.random(n) * 2 - 1).astype(np.float32)
    eta = (1 + rng.random(n)).astype(np.float32)
    assert all(np.array_equal(x, y) for x, y in zip(orc.fresnel_dielectric_ext(cos_i, eta), ref.fresnel_dielectric_ext(cos_i, eta)))


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("kind,h", [("linear", 2e-3), ("radial", 5e-3), ("sd", 1e-3), ("smooth", 3.3e-3)])
def test_oracle_stepper_bit_exact_vs_verbatim_reference(oracle32, kind, h):
    """SURVEY a5-a9, a11 PINNED: er_step, trace, traceTillBoundary and insideShape of heterogeneousrefractive.cpp and the
    SplineDataSource wrappers of splinevolume.cpp, compiled verbatim (oracle/ref_trace.cpp), against the restatement -
    bit for bit, in the reference's own container (hackForSphere: a hard-coded sphere, strict inequality)."""
    from common import REF_SPHERE_CENTRE, REF_SPHERE_RADIUS, ref_sphere_scene
    data, lo, hi, p0, d0, dist = ref_sphere_scene(kind)
    ref = RefTrace(data, lo, hi, h)
    c, r = ref.container()
    assert np.array_equal(c, REF_SPHERE_CENTRE) and r == np.float32(REF_SPHERE_RADIUS)
    props = medium_props(stepsize=h, shape=("sphere", tuple(float(x) for x in c), float(r)))
    orif = oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    # the lookup wrappers and the containment test
    n0, g0 = ref.value_gradient(p0)
    f1, g1 = oracle32.rif_eval(orif, p0, what=2) if hasattr(oracle32, "rif_eval") else (None, None)
    if f1 is not None:
        assert np.array_equal(np.asarray(f1, np.float32), n0) and np.array_equal(np.asarray(g1, np.float32), g0)
    v0 = (d0 * n0[:, None]).astype(np.float32)
    # trace()
    rp, rv, rds, ropl, rok = ref.trace(p0, v0, dist)
    got = oracle32.trace(omed, p0, v0, dist)
    assert 0.2 < rok.mean() < 0.8  # both outcomes are exercised
    assert np.array_equal(got["success"], rok)
    for name, a, b in (("p", got["p"], rp), ("v", got["v"], rv), ("dist_surf", got["dist_surf"], rds), ("opl", got["opl"], ropl)):
        assert np.array_equal(np.asarray(a, np.float32), b), (name, np.abs(np.asarray(a, np.float64) - b).max())
    # traceTillBoundary()
    tp, tv, tds, topl = ref.trace_till_boundary(p0, v0)
    gotb = oracle32.trace_till_boundary(omed, p0, v0)
    for name, a, b in (("p", gotb["p"], tp), ("v", gotb["v"], tv), ("dist_surf", gotb["dist_surf"], tds), ("opl", gotb["opl"], topl)):
        assert np.array_equal(np.asarray(a, np.float32), b), (name, np.abs(np.asarray(a, np.float64) - b).max())
    assert ref.inside_shape(tp).mean() > 0.99  # stepped back inside (the leapfrog step is not exactly reversible in float)


@pytest.mark.parametrize("kind", ["linear", "radial", "sd", "smooth"])
def test_oracle_stepper_bit_exact_vs_reference_golden(oracle32, kind):
    """the same pin without /root/reference: tests/golden/trace_ref.npz holds what the reference's own er_step / trace /
    traceTillBoundary (compiled verbatim, tests/golden/make_golden.py) returned for these seeded rays"""
    from common import REF_SPHERE_CENTRE, REF_SPHERE_RADIUS, ref_sphere_scene
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    data, lo, hi, p0, d0, dist = ref_sphere_scene(kind, n_rays=1024)
    assert float(data.astype(np.float64).sum()) == float(G[kind + "_data_sum"]) and float(p0.astype(np.float64).sum()) == float(G[kind + "_p0_sum"])
    props = medium_props(stepsize=float(G[kind + "_h"]), shape=("sphere", tuple(float(x) for x in REF_SPHERE_CENTRE), REF_SPHERE_RADIUS))
    orif = oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    f, g = oracle32.rif_eval(orif, p0, what=2)
    assert np.array_equal(np.asarray(f, np.float32), G[kind + "_n0"]) and np.array_equal(np.asarray(g, np.float32), G[kind + "_g0"])
    v0 = (d0 * G[kind + "_n0"][:, None]).astype(np.float32)
    got = oracle32.trace(omed, p0, v0, dist)
    assert np.array_equal(got["success"], G[kind + "_success"])
    for name in ("p", "v", "dist_surf", "opl"):
        assert np.array_equal(np.asarray(got[name], np.float32), G[kind + "_" + name]), name
    gotb = oracle32.trace_till_boundary(omed, p0, v0)
    for name in ("p", "v", "dist_surf", "opl"):
        assert np.array_equal(np.asarray(gotb[name], np.float32), G[kind + "_tb_" + name]), name


SAMPLE_DISTANCE_CASES = [("single", False), ("balance", False), ("manual", False), ("maximum", False), ("single", True)]


def _sample_distance_scene(oracle, strategy, aggressive, n_rays=3000):
    """medium + rays for Medium::sampleDistance in the reference's hard-coded sphere; -> (props, data, lo, hi, sdf, ro, rd, mint, xi)"""
    from common import REF_SPHERE_CENTRE, REF_SPHERE_RADIUS, ref_sphere_scene
    from mitsubaer_b200 import fields
    data, lo, hi, p0, d0, dist = ref_sphere_scene("radial", n_rays=n_rays, seed=5)
    props = medium_props(stepsize=2e-3, sigmaS=(9.0, 12.0, 1.6), sigmaA=(1.0, 1.0, 0.4), strategy=strategy,
                         shape=("sphere", tuple(float(x) for x in REF_SPHERE_CENTRE), REF_SPHERE_RADIUS))
    if strategy == "manual":
        props["samplingDensity"] = 11.0
    sdf = fields.sphere_sdf(data.shape[::-1], lo, hi, centre=tuple(float(x) for x in REF_SPHERE_CENTRE), radius=REF_SPHERE_RADIUS).astype(np.float32) if aggressive else None
    rng = np.random.default_rng(17)
    xi = rng.random((n_rays, 2)).astype(np.float32)
    mint = (0.01 * rng.random(n_rays)).astype(np.float32)
    return props, data, lo, hi, sdf, p0, d0, mint, xi


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("strategy,aggressive", SAMPLE_DISTANCE_CASES)
def test_oracle_sample_distance_bit_exact_vs_verbatim_reference(oracle32, strategy, aggressive):
    """SURVEY a10, a12-a14 PINNED: Medium::sampleDistance of heterogeneousrefractive.cpp (:402-568: free-flight strategies,
    the aggressive-tracing loop over the signed distance, pdfs, transmittance, refRatioSq) and evalTransmittance
    (:393-400), compiled verbatim with src/medium/maxexp.h, against the restatement - bit for bit"""
    props, data, lo, hi, sdf, ro, rd, mint, xi = _sample_distance_scene(oracle32, strategy, aggressive)
    orif = oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    if sdf is not None:
        oracle32.medium_set_sdf(omed, oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), sdf), aggressive=True)
    w, sd = oracle32.medium_resolved(omed)
    ref = RefTrace(data, lo, hi, props["stepsize"]).configure(props["sigmaA"], props["sigmaS"], strategy, sd, w, sdf, lo, hi, aggressive)
    a, b = oracle32.sample_distance(omed, ro, rd, mint, xi), ref.sample_distance(ro, rd, mint, xi)
    assert 0.15 < b["success"].mean() < 0.85
    assert np.array_equal(a["success"], b["success"])
    for key in ("t", "p", "d", "optical_length", "ref_ratio_sq", "transmittance", "pdf_success", "pdf_failure"):
        x, y = np.asarray(a[key], np.float32), b[key]
        assert np.array_equal(x, y), (key, np.abs(x.astype(np.float64) - y).max(), np.mean(x != y))
    rng = np.random.default_rng(3)
    t0, t1 = rng.random(500).astype(np.float32), (1 + 3 * rng.random(500)).astype(np.float32)
    sigma_t = np.asarray(props["sigmaS"], np.float32) + np.asarray(props["sigmaA"], np.float32)
    assert np.array_equal(oracle32.eval_transmittance(sigma_t, t0, t1), ref.eval_transmittance(t0, t1))


@pytest.mark.parametrize("strategy,aggressive", SAMPLE_DISTANCE_CASES)
def test_oracle_sample_distance_bit_exact_vs_reference_golden(oracle32, strategy, aggressive):
    """the same pin without /root/reference: the reference's own sampleDistance outputs in tests/golden/trace_ref.npz"""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    tag = "sd_%s_%d_" % (strategy, int(aggressive))
    props, data, lo, hi, sdf, ro, rd, mint, xi = _sample_distance_scene(oracle32, strategy, aggressive, n_rays=1024)
    orif = oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    if sdf is not None:
        oracle32.medium_set_sdf(omed, oracle32.rif_create(volume_desc(data.shape[::-1], lo, hi), sdf), aggressive=True)
    w, sd = oracle32.medium_resolved(omed)
    assert np.float32(w) == G[tag + "weight"] and np.float32(sd) == G[tag + "density"]
    a = oracle32.sample_distance(omed, ro, rd, mint, xi)
    assert np.array_equal(a["success"], G[tag + "success"])
    for key in ("t", "p", "d", "optical_length", "ref_ratio_sq", "transmittance", "pdf_success", "pdf_failure"):
        assert np.array_equal(np.asarray(a[key], np.float32), G[tag + key]), key


def _connection_scene(kind, n=2000):
    """medium (hdielectric boundary: Snell at the exit, as the reference always does) with an sdf child in the reference's
    hard-coded sphere, and shooting problems p1 -> p2 with launch velocities near the straight line"""
    from common import REF_SPHERE_CENTRE, REF_SPHERE_RADIUS, ref_sphere_scene
    from mitsubaer_b200 import fields
    data, lo, hi, p1, d0, dist = ref_sphere_scene(kind, n_rays=n, seed=3)
    c, r = REF_SPHERE_CENTRE, REF_SPHERE_RADIUS
    sdf = fields.sphere_sdf(data.shape[::-1], lo, hi, centre=tuple(float(x) for x in c), radius=r).astype(np.float32)
    props = medium_props(stepsize=2e-3, shape=("sphere", tuple(float(x) for x in c), r), bsdf="hdielectric")
    rng = np.random.default_rng(1)
    q = rng.normal(size=(n, 3))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    p2 = (c + q * (r * (0.3 + 1.2 * rng.random((n, 1))))).astype(np.float32)  # inside and outside the container
    w = p2 - p1
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    w = (w + 0.05 * rng.normal(size=w.shape)).astype(np.float32)
    w = (w / np.linalg.norm(w, axis=1, keepdims=True) * 1.7).astype(np.float32)  # |v0| != n: the renormalisation chain rule
    return props, data, lo, hi, sdf, p1, d0, p2, w


CONNECTION_KINDS = ["linear", "radial", "sd", "smooth"]


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("kind", CONNECTION_KINDS)
def test_oracle_connection_bit_exact_vs_verbatim_reference(oracle32, kind):
    """SURVEY a25 minus the solver, PINNED: er_derivativestep (:798-814), computefdfBDPT (:816-939: the shooting problem's
    residual AND its Jacobian, closest approach by bisection, Snell exit through boundaryVelocityDerivative :1057-1074) and
    computePathLengthsTillClosestP2 (:941-1030) with boundaryVelocity (:1036-1051), compiled verbatim, against the
    restatement - bit for bit.  (What stays unpinned is Ceres' BFGS itself.)"""
    props, data, lo, hi, sdf, p1, d0, p2, w = _connection_scene(kind)
    d = volume_desc(data.shape[::-1], lo, hi)
    omed = oracle32.medium_create(oracle_medium_desc(props), oracle32.rif_create(d, data))
    oracle32.medium_set_sdf(omed, oracle32.rif_create(d, sdf), aggressive=False)
    ref = RefTrace(data, lo, hi, props["stepsize"]).configure((0.4,) * 3, (3.6,) * 3, "single", 4.0, 0.9, sdf, lo, hi, False).set_connection(3, 1e-6)
    n0, _ = ref.value_gradient(p1)
    v0 = (d0 * n0[:, None]).astype(np.float32)
    a, b = oracle32.derivative_trace(omed, p1, v0, 40), ref.derivative_trace(p1, v0, 40)
    for key in ("p", "v", "dpdv0", "dvdv0"):
        assert np.array_equal(np.asarray(a[key], np.float32), b[key]), key
    for sensor in (False, True):
        A, B = oracle32.connection_residual(omed, p1, p2, w, is_sensor=sensor), ref.connection_residual(p1, p2, w, is_sensor=sensor)
        assert set(np.unique(A["status"])) >= {0, 1}  # closest approach inside the medium, and exits through the boundary
        assert np.array_equal(np.asarray(A["error"], np.float32), B["error"])
        assert np.array_equal(np.asarray(A["derror"], np.float32), B["derror"])
    # the connection's lengths along a solved direction: the restated solver's result fed to the reference's re-trace
    sel = slice(0, 300)
    r = oracle32.connect(omed, p1[sel], p2[sel], (p2[sel] - p1[sel]) / np.linalg.norm(p2[sel] - p1[sel], axis=1, keepdims=True), tol2=1e-6, precision=3, seed=5, start_mode=2)
    ok = r["success"].astype(bool)
    assert ok.mean() > 0.5
    L = ref.path_lengths(p1[sel][ok], p2[sel][ok], np.asarray(r["dir_to_p2"], np.float32)[ok])
    assert L["success"].all()
    assert np.array_equal(L["rev_dir"], np.asarray(r["rev_dir"], np.float32)[ok])
    assert np.array_equal(L["optical_dist"], np.asarray(r["optical_dist"], np.float32)[ok]) and np.array_equal(L["dist"], np.asarray(r["dist"], np.float32)[ok])


@pytest.mark.parametrize("kind", CONNECTION_KINDS)
def test_oracle_connection_bit_exact_vs_reference_golden(oracle32, kind):
    """the same pin without /root/reference (tests/golden/trace_ref.npz, key prefix conn_)"""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    props, data, lo, hi, sdf, p1, d0, p2, w = _connection_scene(kind, n=512)
    d = volume_desc(data.shape[::-1], lo, hi)
    omed = oracle32.medium_create(oracle_medium_desc(props), oracle32.rif_create(d, data))
    oracle32.medium_set_sdf(omed, oracle32.rif_create(d, sdf), aggressive=False)
    v0 = (d0 * G["conn_%s_n0" % kind][:, None]).astype(np.float32)
    a = oracle32.derivative_trace(omed, p1, v0, 40)
    for key in ("p", "v", "dpdv0", "dvdv0"):
        assert np.array_equal(np.asarray(a[key], np.float32), G["conn_%s_dt_%s" % (kind, key)]), key
    for sensor in (0, 1):
        A = oracle32.connection_residual(omed, p1, p2, w, is_sensor=bool(sensor))
        assert np.array_equal(np.asarray(A["error"], np.float32), G["conn_%s_error_%d" % (kind, sensor)])
        assert np.array_equal(np.asarray(A["derror"], np.float32), G["conn_%s_derror_%d" % (kind, sensor)])


def _grid_scene():
    rng = np.random.default_rng(31)
    res = (23, 17, 29)
    data = rng.random((res[2], res[1], res[0])).astype(np.float32)
    lo, hi = np.array([-1.0, -0.5, 0.25], np.float32), np.array([1.5, 0.75, 2.0], np.float32)
    p = (lo - 0.1 + rng.random((6000, 3)) * (hi - lo + 0.2)).astype(np.float32)  # some outside: lookupFloat returns 0 there
    p[:8] = [[lo[0], lo[1], lo[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], lo[2]], [hi[0], lo[1], hi[2]], [0, 0, 1], [1.5, 0, 1], [0, 0.75, 1], [0, 0, 2]]
    return res, data, lo, hi, p


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
def test_oracle_grid_lookup_bit_exact_vs_verbatim_reference(oracle32):
    """SURVEY a18 PINNED: GridDataSource::lookupFloat (gridvolume.cpp:337-388) with Transform::scale / translate / operator*
    (transform.cpp) compiled verbatim against the restated density lookup - bit for bit, faces and corners of the box included"""
    res, data, lo, hi, p = _grid_scene()
    ref = RefGrid(data, lo, hi).lookup(p)
    g = oracle32.grid_create(volume_desc(res, lo, hi), data)
    got = oracle32.grid_lookup(g, p)
    assert (ref == 0).sum() > 100 and (ref > 0).sum() > 3000
    assert np.array_equal(got, ref), (np.abs(got - ref).max(), np.mean(got != ref))


def test_oracle_grid_lookup_bit_exact_vs_reference_golden(oracle32):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    res, data, lo, hi, p = _grid_scene()
    g = oracle32.grid_create(volume_desc(res, lo, hi), data)
    assert np.array_equal(oracle32.grid_lookup(g, p), G["grid_lookup"])


def _grid_spectrum_scene():
    rng = np.random.default_rng(37)
    res = (19, 13, 21)
    rgb = rng.random((res[2], res[1], res[0], 3)).astype(np.float32)
    u8 = rng.integers(0, 256, (res[2], res[1], res[0], 3), dtype=np.uint8)
    lo, hi = np.array([-1.0, -0.5, 0.25], np.float32), np.array([1.5, 0.75, 2.0], np.float32)
    p = (lo - 0.1 + rng.random((6000, 3)) * (hi - lo + 0.2)).astype(np.float32)  # some outside: lookupSpectrum returns 0 there
    p[:8] = [[lo[0], lo[1], lo[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], lo[2]], [hi[0], lo[1], hi[2]], [0, 0, 1], [1.5, 0, 1], [0, 0.75, 1], [0, 0, 2]]
    return res, rgb, u8, lo, hi, p


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
def test_oracle_grid_spectrum_lookup_bit_exact_vs_verbatim_reference(oracle32):
    """GridDataSource::lookupSpectrum (gridvolume.cpp:386-463) with its float3 helper (:293-329) compiled verbatim against the
    restated 3-channel lookup - bit for bit, float32 and uint8 payloads (the latter through m_densityMap = i / 255)"""
    res, rgb, u8, lo, hi, p = _grid_spectrum_scene()
    for payload, as_float in ((rgb, rgb), (u8, u8.astype(np.float32) / np.float32(255.0))):
        ref = RefGrid(payload, lo, hi).lookup_spectrum(p)
        g = oracle32.grid_create_spectrum(volume_desc(res, lo, hi), as_float)
        got = oracle32.grid_lookup_spectrum(g, p)
        assert (ref.sum(axis=1) == 0).sum() > 100 and (ref.min(axis=1) > 0).sum() > 3000
        assert np.array_equal(got, ref), (np.abs(got - ref).max(), np.mean(got != ref))


def test_oracle_grid_spectrum_lookup_bit_exact_vs_reference_golden(oracle32):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "grid_spectrum_ref.npz"))
    res, rgb, u8, lo, hi, p = _grid_spectrum_scene()
    g = oracle32.grid_create_spectrum(volume_desc(res, lo, hi), rgb)
    assert np.array_equal(oracle32.grid_lookup_spectrum(g, p), G["f32"])
    g = oracle32.grid_create_spectrum(volume_desc(res, lo, hi), u8.astype(np.float32) / np.float32(255.0))
    assert np.array_equal(oracle32.grid_lookup_spectrum(g, p), G["u8"])


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
def test_oracle_mi_weight_bit_exact_vs_verbatim_reference(oracle32):
    """SURVEY a21's only arithmetic of its own: VolumetricPathTracer::miWeight (volpath.cpp:430-433) compiled verbatim against the
    power heuristic the restated MIS connections use - bit for bit over 12 decades of densities"""
    rng = np.random.default_rng(43)
    a = (10.0 ** rng.uniform(-6, 6, 100000)).astype(np.float32)
    b = (10.0 ** rng.uniform(-6, 6, 100000)).astype(np.float32)
    b[:100] = 0.0  # a delta on the other side: weight 1
    ref = RefFilm().mi_weight(a, b)
    got = oracle32.mi_weight(a, b)
    assert np.all(ref[:100] == 1.0) and np.array_equal(got, ref)


def _film_scene():
    rng = np.random.default_rng(41)
    W, H, n = 37, 29, 20000
    pos = (rng.random((n, 2)) * [W, H]).astype(np.float32)
    pos[:6] = [[0, 0], [W, H], [0.5, 0.5], [W - 1e-3, 1e-3], [12.0, 7.0], [12.5, 7.5]]  # corners, pixel centres and edges
    values = rng.random((n, 5)).astype(np.float32)
    values[:, 4] = 1.0
    values[100, 1], values[200, 3] = np.nan, np.inf  # rejected like imageblock.h:147-152
    return W, H, pos, values


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("ftype", [0, 1])
def test_oracle_filter_and_film_put_bit_exact_vs_verbatim_reference(oracle32, ftype):
    """SURVEY a23 PINNED: ReconstructionFilter::configure (rfilter.cpp:37-55), evalDiscretized (rfilter.h), the box and
    gaussian kernels and ImageBlock::put (imageblock.h:124-206) compiled verbatim against the restatement - bit for bit"""
    ref = RefFilm()
    v, r, s, b = ref.filter_table(ftype)
    ov, orr, os_ = oracle32.filter_table(ftype)
    assert np.array_equal(ov, v) and orr == r and os_ == s and b == int(np.ceil(r - 0.5))
    W, H, pos, values = _film_scene()
    film, ok = ref.film_put(ftype, W, H, pos, values)
    ofilm, ook = oracle32.film_put(ftype, W, H, pos, values)
    assert (~ok).sum() == 2 and np.array_equal(ook, ok)
    assert np.array_equal(ofilm, film), np.abs(ofilm - film).max()


@pytest.mark.parametrize("ftype", [0, 1])
def test_oracle_filter_and_film_put_bit_exact_vs_reference_golden(oracle32, ftype):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    ov, orr, os_ = oracle32.filter_table(ftype)
    assert np.array_equal(ov, G["film_table_%d" % ftype]) and np.float32(orr) == G["film_radius_%d" % ftype]
    W, H, pos, values = _film_scene()
    ofilm, ook = oracle32.film_put(ftype, W, H, pos, values)
    assert np.array_equal(ofilm, G["film_put_%d" % ftype])


def _woodcock_scene(n=3000):
    from mitsubaer_b200 import fields
    res = (24, 20, 28)
    dens = fields.sine_density(res, BOX_MIN, BOX_MAX)
    o = (random_points_in_box(n, 131) * 1.6).astype(np.float32)  # some origins outside the box
    dd = random_directions(n, 132)
    dd[:50, 0] = 0.0  # rays parallel to a slab
    mint = np.zeros(n, np.float32)
    maxt = (np.random.default_rng(133).random(n) * 3 + 0.1).astype(np.float32)
    return res, dens, o, dd, mint, maxt


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("scale", [4.0, 20.0])
def test_oracle_straight_woodcock_bit_exact_vs_verbatim_reference(oracle32, scale):
    """SURVEY a19 PINNED: HeterogeneousMedium::sampleDistance / evalTransmittance (heterogeneous.cpp:546-672, Woodcock tracking)
    with lookupDensity and AABB::rayIntersect, compiled verbatim over the verbatim GridDataSource, fed the oracle's own Philox
    draws through a replaying sampler - success, t, sigma_s and the 2-sample transmittance estimate bit for bit"""
    res, dens, o, dd, mint, maxt = _woodcock_scene()
    n, K = o.shape[0], 512
    d = volume_desc(res, BOX_MIN, BOX_MAX)
    ogrid = oracle32.grid_create(d, dens)
    ref = RefHeterogeneousMedium(RefGrid(dens, BOX_MIN, BOX_MAX), BOX_MIN, BOX_MAX, scale, scale * 1.0, "woodcock")
    xi = np.stack([oracle32.philox(77, i, K) for i in range(n)])
    rok, rt, rden = oracle32.grid_sample_distance(ogrid, d, scale, o, dd, mint, maxt, 77)
    ok, t, ss, T = ref.sample_distance(o, dd, mint, maxt, xi)
    assert ok.sum() > n // 10 and (~ok).sum() > n // 10
    assert np.array_equal(ok, rok) and np.array_equal(t[ok], rt[ok])
    assert np.array_equal(ss[ok][:, 0], (np.float32(0.9) * rden[ok]).astype(np.float32))  # sigmaS = albedo * densityAtT (:641)
    xi = np.stack([oracle32.philox(78, i, K) for i in range(n)])
    assert np.array_equal(ref.eval_transmittance(o, dd, mint, maxt, xi), oracle32.grid_eval_transmittance(ogrid, d, scale, o, dd, mint, maxt, 78))


CAMERAS = [dict(origin=(0.0, 0.0, -4.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=40.0, width=96, height=64),
           dict(origin=(2.5, 1.25, -3.0), target=(0.1, -0.2, 0.3), up=(0.1, 1.0, 0.2), fov=28.0, width=50, height=80)]


def _camera_samples(cam):
    rng = np.random.default_rng(9)
    sp = (rng.random((5000, 2)) * [cam["width"], cam["height"]]).astype(np.float32)
    sp[:4] = [[0, 0], [cam["width"], cam["height"]], [cam["width"] / 2, cam["height"] / 2], [0.5, cam["height"] - 0.5]]
    return sp


@pytest.mark.parametrize("k", [0, 1])
def test_oracle_camera_against_reference_golden(oracle32, k):
    cam = CAMERAS[k]
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_ref.npz"))
    scene = scene_dict(cam["width"], cam["height"], 4)
    scene.update(origin=cam["origin"], target=cam["target"], up=cam["up"], fov=cam["fov"])
    assert np.abs(oracle32.camera_ray(oracle_render_desc(scene), _camera_samples(cam)) - G["camera_d_%d" % k]).max() <= 4e-7


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
@pytest.mark.parametrize("cam", CAMERAS)
def test_oracle_camera_against_verbatim_reference(oracle32, cam):
    """SURVEY a22 (pinhole sensor): PerspectiveCameraImpl's m_cameraToSample composition with the reference's own 4x4
    inversion, sampleRay (perspective.cpp:128-156, 247-269), Transform::lookAt / perspective, compiled verbatim.  The oracle
    and the kernels evaluate a CLOSED FORM of m_sampleToCamera, so this pin is to rounding (<= 4e-7 on unit directions),
    not bit for bit."""
    scene = scene_dict(cam["width"], cam["height"], 4)
    scene.update(origin=cam["origin"], target=cam["target"], up=cam["up"], fov=cam["fov"])
    sp = _camera_samples(cam)
    o, d = RefFilm().camera_rays(cam["origin"], cam["target"], cam["up"], cam["fov"], cam["width"], cam["height"], sp)
    got = oracle32.camera_ray(oracle_render_desc(scene), sp)
    assert np.abs(o - np.asarray(cam["origin"], np.float32)).max() <= 1e-6
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() <= 1e-6
    assert np.abs(got - d).max() <= 4e-7, np.abs(got - d).max()


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
def test_vol_file_is_what_the_reference_loader_reads(oracle32, tmp_path):
    """a .vol v3 file written by this repo's writer (mer_vol_write, the format of mfiles/writeGridToVol.m) is read by
    the reference's OWN loader - SplineDataSource::loadFromFile (splinevolume.cpp:204-317: header, bounding box, payload offset,
    float -> FLOAT copy, prefilter), compiled verbatim - and gives, bit for bit, the spline built from the array in memory"""
    from mitsubaer_b200 import fields
    rng = np.random.default_rng(51)
    res = (19, 23, 17)
    data = (1.0 + rng.random((res[2], res[1], res[0]))).astype(np.float32)
    lo, hi = np.array([-1.25, 0.5, -0.75], np.float32), np.array([0.5, 2.0, 1.0], np.float32)
    path = tmp_path / "rif.vol"
    fields.write_vol(path, data, lo, hi)
    raw = path.read_bytes()
    assert raw[:4] == b"VOL\x03" and len(raw) == 48 + data.size * 4
    pitch = (hi - lo) / (np.array(res, np.float32) - 1)
    p = (lo + 2.1 * pitch + rng.random((4000, 3)) * (hi - lo - 4.2 * pitch)).astype(np.float32)
    from_file, from_array = RefTrace(path), RefTrace(data, lo, hi, 1e-3)
    f1, g1 = from_file.value_gradient(p)
    f2, g2 = from_array.value_gradient(p)
    assert np.array_equal(f1, f2) and np.array_equal(g1, g2)
    assert np.array_equal(from_file.inside_limits(p), from_array.inside_limits(p))
    orif = oracle32.rif_create(volume_desc(res, lo, hi), data)
    f3, g3 = oracle32.rif_eval(orif, p, what=2)
    assert np.array_equal(np.asarray(f3, np.float32), f1) and np.array_equal(np.asarray(g3, np.float32), g1)
    d2, lo2, hi2 = fields.read_vol(path)  # and this repo's reader returns what was written
    assert np.array_equal(d2, data) and np.array_equal(lo2, lo) and np.array_equal(hi2, hi)


@pytest.mark.skipif(not RefTrace.available(), reason="oracle/_ref/libmer_reftrace.so not built (needs /root/reference)")
def test_oracle_medium_property_resolution_vs_verbatim_reference(oracle32):
    """the constructor of HeterogeneousRefractiveMedium (:238-293: automatic mediumSamplingWeight = max albedo, at least 0.5;
    strategy single = the channel with the smallest sigma_t; manual; balance; maximum), compiled verbatim, against the
    restated resolution (what plugins.HeterogeneousRefractiveMedium.configure and mer_medium_create do too)"""
    data, lo, hi = make_field("linear", 16)
    ref = RefTrace(data, lo, hi, 1e-2)
    orif = oracle32.rif_create(volume_desc((16,) * 3, lo, hi), data)
    cases = [dict(sigmaS=(3.6, 3.6, 3.6), sigmaA=(0.4, 0.4, 0.4), strategy="single"),
             dict(sigmaS=(2.0, 3.0, 4.0), sigmaA=(0.5, 0.25, 0.1), strategy="single"),
             dict(sigmaS=(0.1, 0.2, 0.05), sigmaA=(1.0, 1.0, 1.0), strategy="balance"),        # albedo < 0.5 -> weight 0.5
             dict(sigmaS=(0.0, 0.0, 0.0), sigmaA=(1.0, 2.0, 3.0), strategy="single"),          # no scattering -> weight 0
             dict(sigmaS=(2.0, 0.0, 1.0), sigmaA=(0.5, 0.0, 3.0), strategy="balance"),         # a channel with sigma_t = 0
             dict(sigmaS=(2.0, 3.0, 4.0), sigmaA=(0.5, 0.25, 0.1), strategy="manual", samplingDensity=2.5),
             dict(sigmaS=(2.0, 3.0, 0.4), sigmaA=(0.5, 0.25, 0.1), strategy="maximum"),
             dict(sigmaS=(2.0, 3.0, 4.0), sigmaA=(0.5, 0.25, 0.1), strategy="single", mediumSamplingWeight=0.8)]
    for c in cases:
        props = medium_props(stepsize=1e-2, **c)
        omed = oracle32.medium_create(oracle_medium_desc(props), orif)
        w, sd = oracle32.medium_resolved(omed)
        rw, rsd, rst = ref.resolve(c["sigmaA"], c["sigmaS"], c["strategy"], c.get("mediumSamplingWeight", -1.0), c.get("samplingDensity", 0.0))
        assert np.float32(w) == np.float32(rw), (c, w, rw)
        if c["strategy"] in ("single", "manual"):
            assert np.float32(sd) == np.float32(rsd), (c, sd, rsd)
        assert rst == {"balance": 0, "single": 1, "manual": 2, "maximum": 3}[c["strategy"]]
        oracle32.medium_destroy(omed)


@pytest.mark.skipif(not RefPhase.available(), reason="oracle/_ref/libmer_refphase.so not built (needs /root/reference)")
@pytest.mark.parametrize("importance", [False, True])
def test_oracle_hdielectric_sample_vs_verbatim_reference(oracle32, importance):
    """SURVEY f-3 PINNED: HSmoothDielectric::sample with reflect / refract / getEtaInvEta (src/bsdfs/hdielectric.cpp:86-125,
    244-300), compiled verbatim and driven through Frame::toLocal / toWorld like Intersection does, against the restated
    world-space form: the same event (reflection or refraction, total internal reflection included), weight and relative
    index exactly, the direction to rounding (the reference goes through a local frame, the restatement does not)"""
    rng = np.random.default_rng(61)
    n = 20000
    d = random_directions(n, 62)
    N = random_directions(n, 63)
    eta = (1.05 + 0.6 * rng.random(n)).astype(np.float32)
    u = rng.random(n).astype(np.float32)
    ro, rw, res, rtr = RefPhase().hdielectric_sample(d, N, eta, u, importance)
    go, gw, ges, gtr = oracle32.hdielectric_sample(d, N, eta, u, importance)
    assert 0.05 < rtr.mean() < 0.95 and np.array_equal(gtr, rtr)
    assert np.array_equal(gw, rw) and np.array_equal(ges, res)
    assert np.abs(go - ro).max() <= 1e-6
    # total internal reflection from inside: always reflected
    inside = np.einsum("ij,ij->i", d, N) > 0
    tir = inside & (1 - np.einsum("ij,ij->i", d, N) ** 2 > 1 / eta.astype(np.float64) ** 2 + 1e-6)
    assert tir.sum() > 100 and not rtr[tir].any()


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_hdielectric_sample_vs_reference_golden(oracle32, mode):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "phase_ref.npz"))
    go, gw, ges, gtr = oracle32.hdielectric_sample(G["hd_d"], G["hd_n"], G["hd_eta"], G["hd_u"], bool(mode))
    assert np.array_equal(gtr, G["hd_transmitted_%d" % mode]) and np.array_equal(gw, G["hd_weight_%d" % mode])
    assert np.array_equal(ges, G["hd_etascale_%d" % mode]) and np.abs(go - G["hd_out_%d" % mode]).max() <= 1e-6


def test_spline_interpolates_data_at_nodes(oracle64):
    """the prefilter makes the cubic B-spline INTERPOLATE the samples (that is what build3d is for)"""
    res = (16, 14, 12)
    data, lo, hi = make_field("random", res, seed=2)
    h = oracle64.rif_create(volume_desc(res, lo, hi), data)
    idx = np.stack(np.meshgrid(np.arange(2, 14), np.arange(2, 12), np.arange(2, 10), indexing="ij"), -1).reshape(-1, 3)
    p = lo.astype(np.float64) + idx * ((hi - lo).astype(np.float64) / (np.array(res) - 1))
    f, _ = oracle64.rif_eval(h, p, 0)
    assert np.max(np.abs(f - data[idx[:, 2], idx[:, 1], idx[:, 0]])) < 1e-6  # float32 bbox -> node positions
    oracle64.rif_destroy(h)


def test_leapfrog_invariants(oracle64):
    """known answers of heterogeneousrefractive.cpp:653-661: constant gradient => v_x, v_z constant and
    v_y linear in arc length; radial field => (p x v) conserved by the discrete scheme"""
    h, steps, n = 2e-3, 300, 500
    p0 = (random_points_in_box(n, 1, margin=0.3) * 0.5).astype(np.float64)
    d0 = random_directions(n, 2).astype(np.float64)
    props = medium_props(stepsize=h)
    for kind in ("linear", "radial"):
        data, lo, hi = make_field(kind, 40)
        orif = oracle64.rif_create(volume_desc((40,) * 3, lo, hi), data)
        omed = oracle64.medium_create(oracle_medium_desc(props), orif)
        n0, g0 = oracle64.rif_eval(orif, p0, 2)
        v0 = d0 * n0[:, None]
        out = oracle64.trace(omed, p0, v0, np.full(n, h * steps))
        ok = out["success"]
        assert ok.sum() > n // 2
        if kind == "linear":
            a = np.median(g0[:, 1])
            assert np.max(np.abs(out["v"][ok][:, [0, 2]] - v0[ok][:, [0, 2]])) < 1e-5
            assert np.max(np.abs(out["v"][ok][:, 1] - (v0[ok][:, 1] + a * out["dist_surf"][ok]))) < 1e-5
        else:
            # exact for a radial n; the 40^3 spline of it is radial only to interpolation error
            assert np.max(np.abs(np.cross(out["p"][ok], out["v"][ok]) - np.cross(p0[ok], v0[ok]))) < 5e-5
        # int(dist/h) full steps + ONE remainder step, even when the remainder is ~0 (quirk 3)
        assert np.all(out["nsteps"][ok] == int(np.float64(h * steps) / np.float64(np.float32(h))) + 1)  # stepsize is a Float property
        oracle64.medium_destroy(omed)
        oracle64.rif_destroy(orif)


def test_trace_quirks(oracle32):
    """appendix A quirks 3-5: truncated step count + remainder step, step back with -h, distSurf one h short"""
    data, lo, hi = make_field("radial", 32)
    props = medium_props(stepsize=1e-2)
    orif = oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    p0 = np.array([[0.0, 0.0, 0.0]], np.float32)
    v0 = np.array([[0.0, 0.0, 1.0]], np.float32) * oracle32.rif_eval(orif, p0, 0)[0][:, None]
    inside = oracle32.trace(omed, p0, v0, np.array([0.255], np.float32))
    assert inside["success"][0] and inside["nsteps"][0] == 26 and abs(inside["dist_surf"][0] - 0.255) < 1e-6
    out = oracle32.trace(omed, p0, v0, np.array([5.0], np.float32))
    assert not out["success"][0]
    k = out["nsteps"][0]  # steps until outside (k-1 forward incl. the exiting one) + 1 back
    assert abs(out["dist_surf"][0] - (k - 2) * 1e-2) < 1e-4 and out["p"][0, 2] <= 1.0
    tb = oracle32.trace_till_boundary(omed, p0, v0)
    assert tb["nsteps"][0] == k and abs(tb["dist_surf"][0] - (k - 3) * 1e-2) < 1e-4  # one h short (:761)
    assert np.allclose(tb["p"], out["p"], atol=1e-6)


def test_sample_distance_semantics(oracle32):
    data, lo, hi = make_field("linear", 24)
    props = medium_props(stepsize=1e-2, strategy="balance", sigmaS=(2.0, 3.0, 4.0), sigmaA=(0.5, 0.25, 0.1))
    orif = oracle32.rif_create(volume_desc((24,) * 3, lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props), orif)
    w, sd = oracle32.medium_resolved(omed)
    assert w == pytest.approx(max(2 / 2.5, 3 / 3.25, 4 / 4.1))  # max albedo (>= 0.5), :239-255
    n = 4000
    o = random_points_in_box(n, 3, margin=0.1)
    d = random_directions(n, 4)
    xi = np.random.default_rng(5).random((n, 2)).astype(np.float32)
    r = oracle32.sample_distance(omed, o, d, 0.0, xi)
    # xi >= w  => no medium interaction is sampled: traceTillBoundary, failure (:447-450, :495-498)
    assert not r["success"][xi[:, 0] >= w].any()
    s = r["success"]
    ch = np.minimum((xi[:, 1] * 3).astype(int), 2)
    sig_t = np.array([2.5, 3.25, 4.1], np.float32)
    expect_t = -np.log(1 - (xi[:, 0] / np.float32(w)).astype(np.float64)) / sig_t[ch]
    assert np.allclose(r["t"][s], expect_t[s], rtol=1e-5)
    # balance pdfs (:540-548) and transmittance (:557) at the geometric distance
    T = np.exp(-sig_t[None, :] * r["t"][:, None])
    assert np.allclose(r["transmittance"], T, rtol=1e-5)
    assert np.allclose(r["pdf_success"], w * (sig_t * T).mean(axis=1), rtol=1e-5)
    assert np.allclose(r["pdf_failure"], w * T.mean(axis=1) + (1 - w), rtol=1e-5)
    # v = n * d: |mRec.d| ~ n at the end point (quirk 2); refRatioSq = (n_end/n_start)^2 (quirk 7)
    n_end = oracle32.rif_eval(orif, r["p"], 0)[0]
    n_start = oracle32.rif_eval(orif, o, 0)[0]
    assert np.allclose(np.linalg.norm(r["d"], axis=1), n_end, rtol=2e-3)
    assert np.allclose(r["ref_ratio_sq"], (n_end / n_start) ** 2, rtol=1e-5)
    # start outside insideVolumeLimits (quirk 8)
    far = oracle32.sample_distance(omed, np.array([[5.0, 0, 0]]), np.array([[1.0, 0, 0]]), 0.0, np.array([[0.1, 0.1]]))
    assert not far["success"][0] and np.all(far["transmittance"] == 0) and far["pdf_failure"][0] == 1


def test_hg_normalisation_and_mean_cosine(oracle32):
    for g in (0.9, -0.3, 0.0):
        n = 200000
        wi = np.repeat(np.array([[0.0, 0.0, 1.0]], np.float32), n, 0)
        xi = np.random.default_rng(6).random((n, 2)).astype(np.float32)
        wo, pdf = oracle32.hg_sample(g, wi, xi)
        # wi points back along the incoming ray: mean cosine of the scattering angle is -<wi, wo>
        assert abs(np.mean(-(wi * wo).sum(axis=1)) - g) < 5e-3
        assert np.allclose(pdf, oracle32.hg_eval(g, wi, wo))
        mu = np.linspace(-1, 1, 20001)
        dirs = np.stack([np.sqrt(1 - mu ** 2), 0 * mu, mu], 1).astype(np.float32)
        val = oracle32.hg_eval(g, np.repeat(wi[:1], mu.size, 0), dirs)
        assert abs(np.trapezoid(val, mu) * 2 * np.pi - 1) < 1e-3


def test_filter_tables_and_film_put(oracle32):
    vals, r, s = oracle32.filter_table(1)
    assert r == 2.0 and s == pytest.approx(31 / 2.0) and vals[31] == 0
    assert abs(vals[:31].sum() * 2 * r / 31 - 1) < 1e-6  # rfilter.cpp:51-54 normalisation
    vals, r, s = oracle32.filter_table(0)
    assert r == pytest.approx(0.5 + 1e-5) and np.allclose(vals[:31], vals[0])


def test_philox_known_answer(oracle32):
    """Random123 known-answer test for philox4x32-10: counter = key = 0"""
    def as_u24(words):
        return np.array([(w >> 8) / 16777216.0 for w in words], np.float32)
    assert np.array_equal(oracle32.philox(0, 0, 4), as_u24([0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]))
    a, b = oracle32.philox(20201201, 77, 64), oracle32.philox(20201201, 78, 64)
    assert not np.array_equal(a, b) and np.all((a >= 0) & (a < 1))


def test_render_white_furnace_and_sharding(oracle32):
    res = 16
    from mitsubaer_b200 import fields
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    data = np.full((res,) * 3, 1.33, np.float32)
    props = medium_props(stepsize=5e-2, sigmaS=4.0, sigmaA=0.0)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), data)
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), orif)
    scene = scene_dict(16, 12, 4, rfilter="box", quad=False)
    film, st = oracle32.render(omed, oracle_render_desc(scene, rr_depth=1000))
    rgb = oracle32.film_develop(film)
    assert st.samples == 16 * 12 * 4 and np.allclose(rgb, 1.0, atol=1e-4)
    parts = [oracle32.render(omed, oracle_render_desc(scene, rr_depth=1000, sample_begin=r, sample_stride=3)) for r in range(3)]
    assert sum(p[1].samples for p in parts) == st.samples and sum(p[1].ray_steps for p in parts) == st.ray_steps
    assert np.allclose(sum(p[0] for p in parts), film, rtol=1e-5, atol=1e-6)


def test_render_with_an_albedo_volume(oracle32):
    """the `albedo` child (heterogeneous.cpp:262-268, :646-649) in the oracle's Woodcock walk: a constant RGB grid is the
    constant albedo (to the rounding of the trilinear weights), a per-channel grid scales each channel's single-scattered part
    and leaves the path geometry (events, steps) alone"""
    res = 16
    from mitsubaer_b200 import fields
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), fields.radial_rif((res,) * 3, lo, hi))
    dens = oracle32.grid_create(volume_desc((12,) * 3, BOX_MIN, BOX_MAX), fields.sine_density((12,) * 3, BOX_MIN, BOX_MAX))
    props = medium_props(stepsize=5e-2, albedo=(0.9, 0.6, 0.3), densityScale=4.0)
    scene = scene_dict(12, 12, 4, rfilter="box")
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5, has_density=True), orif, dens)
    film0, st0 = oracle32.render(omed, oracle_render_desc(scene))
    const = np.broadcast_to(np.array((0.9, 0.6, 0.3), np.float32), (9, 8, 7, 3)).copy()
    oracle32.medium_set_albedo_grid(omed, oracle32.grid_create_spectrum(volume_desc((7, 8, 9), BOX_MIN, BOX_MAX), const))
    film1, st1 = oracle32.render(omed, oracle_render_desc(scene))
    assert st1.ray_steps == st0.ray_steps and st1.scatter_events == st0.scatter_events and st0.scatter_events > 100
    assert np.allclose(film1, film0, rtol=2e-5, atol=1e-6)
    red = const.copy()
    red[..., 1:] = 0.0  # no green or blue scattering at all: those channels keep only unscattered light
    oracle32.medium_set_albedo_grid(omed, oracle32.grid_create_spectrum(volume_desc((7, 8, 9), BOX_MIN, BOX_MAX), red))
    film2, st2 = oracle32.render(omed, oracle_render_desc(scene))
    assert st2.ray_steps == st0.ray_steps
    assert np.allclose(film2[..., 0], film0[..., 0], rtol=2e-5, atol=1e-6)
    assert film2[..., 1].sum() < film0[..., 1].sum() and np.all(film2[..., 1] <= film0[..., 1] + 1e-6)
    oracle32.medium_set_albedo_grid(omed, None)
    assert np.array_equal(oracle32.render(omed, oracle_render_desc(scene))[0], film0)


def test_hessian_and_derivative_step_consistency(oracle64):
    """a25 building blocks of the oracle: the spline Hessian is the derivative of its gradient, and the Jacobians
    carried by er_derivativestep (:798-814) are the derivatives of the end state w.r.t. the launch velocity"""
    data, lo, hi = make_field("smooth", 36)
    orif = oracle64.rif_create(volume_desc((36,) * 3, lo, hi), data)
    p = random_points_in_box(200, 1, margin=0.4).astype(np.float64)
    f, g, H = oracle64.rif_eval_hessian(orif, p)
    eps = 1e-5
    for j in range(3):
        dp = np.zeros(3)
        dp[j] = eps
        gp, gm = oracle64.rif_eval(orif, p + dp, 1)[1], oracle64.rif_eval(orif, p - dp, 1)[1]
        assert np.abs((gp - gm) / (2 * eps) - H[:, :, j]).max() < 1e-4
    med = oracle64.medium_create(oracle_medium_desc(medium_props(stepsize=5e-3)), orif)
    d0 = random_directions(200, 2).astype(np.float64)
    v0 = d0 * oracle64.rif_eval(orif, p, 0)[0][:, None]
    r = oracle64.derivative_trace(med, p, v0, 50)
    plain = oracle64.trace(med, p, v0, np.full(200, 50 * np.float64(np.float32(5e-3))))
    assert np.abs(r["p"] - plain["p"]).max() < 1e-9  # same trajectory as er_step (the last step of trace() is a ~0 remainder)
    for j in range(3):
        dv = np.zeros(3)
        dv[j] = 1e-6
        rp, rm = oracle64.derivative_trace(med, p, v0 + dv, 50), oracle64.derivative_trace(med, p, v0 - dv, 50)
        # transported Jacobians agree with differences of the discrete map to O(h)
        assert np.abs((rp["p"] - rm["p"]) / 2e-6 - r["dpdv0"][:, :, j]).max() < 2e-3
        assert np.abs((rp["v"] - rm["v"]) / 2e-6 - r["dvdv0"][:, :, j]).max() < 2e-3


def test_curved_connection_recovers_known_ray(oracle64):
    """makeDirectConnections with the LM minimiser: p2 is the end point of a known eikonal ray"""
    data, lo, hi = make_field("smooth", 36)
    orif = oracle64.rif_create(volume_desc((36,) * 3, lo, hi), data)
    med = oracle64.medium_create(oracle_medium_desc(medium_props(stepsize=5e-3, strategy="single")), orif)
    n = 150
    p1 = (random_points_in_box(n, 3) * 0.3).astype(np.float64)
    d0 = random_directions(n, 4).astype(np.float64)
    v0 = d0 * oracle64.rif_eval(orif, p1, 0)[0][:, None]
    truth = oracle64.trace(med, p1, v0, np.full(n, 0.4))
    assert truth["success"].all()
    seeds = random_directions(n, 5) * 0.3 + d0
    seeds /= np.linalg.norm(seeds, axis=1, keepdims=True)
    r = oracle64.connect(med, p1, truth["p"], seeds, seed=9)
    ok = r["success"]
    assert ok.mean() > 0.8
    dirn = r["dir_to_p2"] / np.linalg.norm(r["dir_to_p2"], axis=1, keepdims=True)
    assert np.abs(dirn - d0)[ok].max() < 6e-3 and np.abs(r["dist"] - 0.4)[ok].max() < 2e-3
    assert np.abs(r["optical_dist"] - truth["opl"])[ok].max() < 4e-3
    res = oracle64.connection_residual(med, p1[ok], truth["p"][ok], r["dir_to_p2"][ok])
    assert (0.5 * (res["error"] ** 2).sum(axis=1) < 1e-6).all() and (res["status"] == 0).all()


def test_oracle_hdielectric_boundary(oracle32):
    """next-row 3: the oracle's Fresnel container (hdielectric.cpp:244-300).  (i) white furnace: eta^2 factors cancel;
    (ii) a clear slab of index n reflects 2R/(1+R) of a source on the camera side (incoherent multiple reflection)."""
    from common import BOX_MAX, BOX_MIN, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import fields
    res, n = 16, 1.5
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), np.full((res,) * 3, n, np.float32))
    med = oracle32.medium_create(oracle_medium_desc(medium_props(stepsize=5e-2, sigmaS=2.0, sigmaA=0.0, bsdf="hdielectric"), 0.3), orif)
    film, st = oracle32.render(med, oracle_render_desc(scene_dict(16, 16, 16, rfilter="box", quad=False), rr_depth=1000))
    assert np.allclose(oracle32.film_develop(film), 1.0, atol=3e-4)
    assert st.boundary_exits > st.samples * 0.5  # internal reflections are counted as surface events

    clear = oracle32.medium_create(oracle_medium_desc(medium_props(stepsize=5e-2, sigmaS=0.0, sigmaA=0.0, mediumSamplingWeight=0.0,
                                                                   bsdf="hdielectric"), 0.0), orif)
    scene = scene_dict(4, 4, 2048, rfilter="box", quad=False)
    scene.update(fov=2.0, envRadiance=0.0, quad=dict(origin=(-50.0, -50.0, -6.0), u=(100.0, 0.0, 0.0), v=(0.0, 100.0, 0.0), radiance=(1.0, 1.0, 1.0)))
    film, st = oracle32.render(clear, oracle_render_desc(scene, rr_depth=1000))
    R = ((n - 1) / (n + 1)) ** 2
    expect = 2 * R / (1 + R)
    got = oracle32.film_develop(film).mean()
    assert abs(got - expect) < 4 * np.sqrt(expect * (1 - expect) / (16 * 2048)) + 1e-3, (got, expect)


def test_oracle_direct_connections_agree_with_random_walk(oracle32):
    """next-row 1: the oracle's next-event estimation has the expectation of its random walk (coarse CPU-sized check;
    the tight statistical version runs on the GPU) and its connections converge from the straight first guess"""
    from common import BOX_MAX, BOX_MIN, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import fields
    res = 16
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), fields.linear_rif((res,) * 3, lo, hi))
    props = medium_props(stepsize=1e-2, sigmaS=1.5, sigmaA=0.5)
    med = oracle32.medium_create(oracle_medium_desc(props, 0.5), orif)
    scene = scene_dict(64, 64, 12, rfilter="box", quad=True)
    scene["envRadiance"] = 0.0
    means, conn = {}, {}
    for nee in (False, "mis", True):
        film, st = oracle32.render(med, oracle_render_desc(scene, direct_connections=nee, props=props))
        means[nee] = oracle32.film_develop(film)[..., 0].mean()
        conn[nee] = st.connections
    assert st.connections > 30000 and st.connections_failed < 0.01 * st.connections
    assert st.connection_steps / st.connections < 4 * (2.5 / 1e-2)  # ~2-3 residual evaluations of ~200 steps each
    assert abs(means[True] / means[False] - 1 + 0.75 * 2.0 * 1e-2) < 0.05, means  # walk noise ~1.5 % at 49k samples
    # volpath's power heuristic between the two (volpath.cpp:120-147, 164-173, 430-433): same expectation, and the emitter
    # hits of connected chains asked for their weight (one more solve each)
    assert conn["mis"] > conn[True]
    assert min(means[True], means[False]) * 0.97 < means["mis"] < max(means[True], means[False]) * 1.03, means


def test_oracle_transient_film(oracle32):
    """next-row 2 (film half): path-length resolved film of the fork (film.cpp:56-78, bdpt_proc.cpp:147-176, 446-449)"""
    from common import BOX_MAX, BOX_MIN, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import fields
    res = 16
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), np.full((res,) * 3, 1.5, np.float32))
    # (i) a clear slab of index 1.5 in front of a bright wall: camera -> box 3, inside 2 * 1.5, box -> wall 2: length 8
    clear = oracle32.medium_create(oracle_medium_desc(medium_props(stepsize=1e-2, sigmaS=0.0, sigmaA=0.0, mediumSamplingWeight=0.0), 0.0), orif)
    scene = scene_dict(8, 8, 8, rfilter="box", quad=False)
    scene.update(fov=2.0, envRadiance=0.0, quad=dict(origin=(-50.0, -50.0, 3.0), u=(100.0, 0.0, 0.0), v=(0.0, 100.0, 0.0), radiance=(1.0, 1.0, 1.0)),
                 transient=dict(minBound=7.0, maxBound=9.0, binWidth=0.125))
    film, st = oracle32.render(clear, oracle_render_desc(scene))
    assert film.shape == (8, 8, 3 * 16 + 2)
    frames = oracle32.film_develop(film)[..., 0].mean(axis=(0, 1))
    # the walk stops up to 2 h short of the far face (exit quirk), i.e. lengths in (8 - 0.03, 8 + 1e-3]
    assert abs(frames[7] + frames[8] - 1.0) < 1e-5 and frames[:7].sum() == 0 and frames[9:].sum() == 0
    scene["transient"]["calibrated"] = True  # the camera segment (3) is not counted
    scene["transient"].update(minBound=4.0, maxBound=6.0)
    frames = oracle32.film_develop(oracle32.render(clear, oracle_render_desc(scene))[0])[..., 0].mean(axis=(0, 1))
    assert abs(frames[7] + frames[8] - 1.0) < 1e-5
    # (ii) scattering medium + direct connections: the frames add up to the steady-state image of the same samples
    lin = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), fields.linear_rif((res,) * 3, lo, hi))
    props = medium_props(stepsize=2e-2, sigmaS=1.5, sigmaA=0.5)
    med = oracle32.medium_create(oracle_medium_desc(props, 0.5), lin)
    for nee in (False, True):
        scene = scene_dict(16, 16, 8, rfilter="gaussian")
        scene["envRadiance"] = 0.0
        steady, _ = oracle32.render(med, oracle_render_desc(scene, direct_connections=nee, props=props))
        scene["transient"] = dict(minBound=0.0, maxBound=64.0, binWidth=0.5)
        trans, _ = oracle32.render(med, oracle_render_desc(scene, direct_connections=nee, props=props))
        rgb = trans[..., :-2].reshape(16, 16, -1, 3)
        assert np.allclose(rgb.sum(axis=2), steady[..., :3], rtol=1e-5, atol=1e-6)
        assert np.array_equal(trans[..., -2:], steady[..., 3:])
        assert (rgb.sum(axis=(0, 1, 3)) > 0).sum() >= 5  # light arrives spread over several frames


def test_oracle_radiance_scaling_conventions(oracle32):
    """DESIGN.md 6c observation: a lossless GRIN medium in an hdielectric container under a uniform environment is a white
    furnace.  With the reciprocal of the fork's refRatioSq ("physical") every path's factors telescope to 1 (up to the
    O(h) gap between the last interior point and the surface); with the fork's own factor they give (n_exit/n_entry)^4."""
    from common import make_field, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    data, lo, hi = make_field("linear", 32)
    orif = oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data)
    out = {}
    for scaling in ("physical", "reference"):
        props = medium_props(stepsize=1e-2, sigmaS=2.0, sigmaA=0.0, bsdf="hdielectric", radianceScaling=scaling)
        med = oracle32.medium_create(oracle_medium_desc(props, 0.3), orif)
        film, _ = oracle32.render(med, oracle_render_desc(scene_dict(24, 24, 16, rfilter="box", quad=False), rr_depth=1000))
        out[scaling] = oracle32.film_develop(film)
    assert np.abs(out["physical"] - 1.0).max() < 3e-3
    assert np.abs(out["reference"] - 1.0).max() > 0.1


def test_oracle_light_tracing_agrees_with_camera_tracer(oracle32):
    """next-row 2, walk half: the oracle's emitter-side walk with sensor connections estimates the image its camera walk
    does (constant index, quad outside the field of view so that all light scatters first; coarse CPU-sized check)"""
    from common import BOX_MAX, BOX_MIN, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import fields
    res = 16
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), np.full((res,) * 3, 1.5, np.float32))
    props = medium_props(stepsize=1e-2, sigmaS=1.5, sigmaA=0.5)
    med = oracle32.medium_create(oracle_medium_desc(props, 0.5), orif)
    scene = scene_dict(64, 64, 6, rfilter="box")
    scene.update(fov=30.0, envRadiance=0.0)
    cam, st_c = oracle32.render(med, oracle_render_desc(scene, direct_connections=True, props=props))
    lit, st_l = oracle32.render(med, oracle_render_desc(scene, light_tracing=True, props=props))
    a, b = oracle32.film_develop(cam)[..., 0], oracle32.film_develop(lit)[..., 0]
    assert np.allclose(lit[..., -1], 1.0) and st_l.connections > 5000
    assert abs(b.mean() / a.mean() - 1) < 0.05, (a.mean(), b.mean())
    ab, bb = a.reshape(8, 8, 8, 8).mean(axis=(1, 3)), b.reshape(8, 8, 8, 8).mean(axis=(1, 3))
    assert np.corrcoef(ab.ravel(), bb.ravel())[0, 1] > 0.98


@pytest.mark.parametrize("modulation", ["sine", "square", "hamiltonian"])
def test_oracle_cw_tof_modulation(oracle32, modulation):
    """continuous-wave ToF film (PathLengthSampler::correlationFunction, src/librender/pathlengthsampler.cpp:66-96): the
    clear-slab scene puts all light at optical length 8 (minus up to 0.03 for the walk's short exit), so every pixel
    is the correlation function at that length"""
    from common import BOX_MAX, BOX_MIN, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    from mitsubaer_b200 import fields
    res = 16
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    orif = oracle32.rif_create(volume_desc((res,) * 3, lo, hi), np.full((res,) * 3, 1.5, np.float32))
    clear = oracle32.medium_create(oracle_medium_desc(medium_props(stepsize=1e-2, sigmaS=0.0, sigmaA=0.0, mediumSamplingWeight=0.0), 0.0), orif)
    lam, phase = 5.0, 60.0
    scene = scene_dict(8, 8, 8, rfilter="box", quad=False)
    scene.update(fov=2.0, envRadiance=0.0, quad=dict(origin=(-50.0, -50.0, 3.0), u=(100.0, 0.0, 0.0), v=(0.0, 100.0, 0.0), radiance=(1.0, 1.0, 1.0)),
                 transient=dict(minBound=0.0, maxBound=100.0, binWidth=1.0, modulation=modulation, **{"lambda": lam, "phase": phase}))
    film, _ = oracle32.render(clear, oracle_render_desc(scene))
    assert film.shape == (8, 8, 5)  # one frame under a modulation (film.cpp:76-78)
    got = oracle32.film_develop(film)[..., 0].mean()
    pl = 8.0 - 0.015 + np.deg2rad(phase) * lam / (2 * np.pi)
    expect = {"sine": np.cos(pl * 2 * np.pi / lam), "square": 4 / lam * (abs(pl % lam - lam / 2) - lam / 4),
              "hamiltonian": (lambda t: 6 * t / lam if t < lam / 6 else 1.0 if t < lam / 2 else 1 - (t - lam / 2) * 6 / lam if t < 2 * lam / 3 else 0.0)(pl % lam)}[modulation]
    assert abs(got - expect) < 0.03, (modulation, got, expect)


def test_oracle_float_vs_double_on_the_widened_path(oracle32, oracle64):
    """R9 for the "next" rows: the float restatement (Mitsuba's Float, what the GPU is gated against) and the double one
    (the authors' FLOATDEBUG arithmetic for the eikonal math) walk the same Philox streams; direct connections, the Fresnel
    boundary and the transient film give the same image up to the few paths that rounding sends different ways"""
    from common import make_field, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict
    data, lo, hi = make_field("linear", 24)
    props = medium_props(stepsize=2e-2, sigmaS=1.5, sigmaA=0.5, bsdf="hdielectric")
    scene = scene_dict(24, 24, 6, rfilter="box")
    scene.update(envRadiance=0.0, transient=dict(minBound=4.0, maxBound=36.0, binWidth=1.0))
    films = {}
    for name, orc in (("f32", oracle32), ("f64", oracle64)):
        med = orc.medium_create(oracle_medium_desc(props, 0.5), orc.rif_create(volume_desc((24,) * 3, lo, hi), data))
        films[name], st = orc.render(med, oracle_render_desc(scene, direct_connections=True, props=props))
        assert st.connections > 2000 and st.connections_failed < 0.3 * st.connections
    a, b = films["f32"][..., :-2].reshape(24, 24, 32, 3), films["f64"][..., :-2].reshape(24, 24, 32, 3)
    assert abs(a.sum() / b.sum() - 1) < 0.02
    pa, pb = a.sum(axis=(0, 1, 3)), b.sum(axis=(0, 1, 3))  # time profile
    assert np.allclose(pa, pb, rtol=0.1, atol=0.02 * pb.max())
    ia, ib = a.sum(axis=(2, 3)), b.sum(axis=(2, 3))
    assert np.mean(np.abs(ia - ib) <= 0.02 * ib + 1e-3 * ib.max()) > 0.9
