"""GPU parity, rows a20-a24: the wavefront integrator vs the CPU oracle renderer.

Both walk the SAME Philox4x32-10 streams (key = seed, counter = (pixel, sample, draw)), so apart from
FP32 rounding a sample's path is identical; the image gates are those of SURVEY §8d (iv) and in addition
a much tighter direct comparison that the shared streams make possible."""
import numpy as np
import pytest

import mitsubaer_b200 as mer
from common import (BOX_MAX, BOX_MIN, make_field, medium_props, oracle_medium_desc, oracle_render_desc, scene_dict)
from oracle.oracle import volume_desc

pytestmark = pytest.mark.gpu


def setup(oracle, kind="radial", res=40, props=None, g=0.9, density=False, mode="tricubic"):
    props = props or medium_props(stepsize=1e-2)
    data, lo, hi = make_field(kind, res)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi, mode=mode)
    med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=g))
    orif = oracle.rif_create(volume_desc(rif.getResolution(), lo, hi), data)
    oden = None
    keep = [rif]
    if density:
        dres = (32, 32, 32)
        dens = mer.fields.sine_density(dres, BOX_MIN, BOX_MAX)
        grid = mer.GridDataSource(data=dens, min=BOX_MIN, max=BOX_MAX)
        med.addChild("density", grid)
        oden = oracle.grid_create(volume_desc(dres, BOX_MIN, BOX_MAX), dens)
        keep.append(grid)
    med.configure()
    omed = oracle.medium_create(oracle_medium_desc(props, g, has_density=density), orif, oden)
    return med, omed, keep


def image_gates(film_gpu, film_cpu, spp):
    """§8d (iv): per-pixel k-sigma and relMSE at equal spp.  With shared RNG streams the images are
    far closer than the statistical bound; both are asserted."""
    w_g, w_c = film_gpu[..., 4], film_cpu[..., 4]
    assert np.allclose(w_g, w_c, rtol=1e-5, atol=1e-5)  # filter weights: same sample positions
    rgb_g = film_gpu[..., :3] / np.maximum(w_g, 1e-9)[..., None]
    rgb_c = film_cpu[..., :3] / np.maximum(w_c, 1e-9)[..., None]
    diff = np.abs(rgb_g - rgb_c)
    # statistical gate: a pixel estimate at `spp` samples has std <= ~mean, so k*sigma/sqrt(N) with k=4
    sigma = np.maximum(rgb_c, 0.05)
    frac_ok = np.mean(diff <= 4 * sigma * np.sqrt(2.0 / spp))
    assert frac_ok >= 0.999, frac_ok
    relmse = np.mean(diff ** 2) / np.mean(rgb_c ** 2)
    return float(relmse), float(np.mean(diff <= 1e-3 * np.maximum(rgb_c, 1.0)))


@pytest.mark.parametrize("kind,strategy,filt", [("linear", "single", "gaussian"), ("radial", "balance", "box"),
                                                ("sd", "single", "box"), ("radial", "maximum", "box")])
def test_homogeneous_image_matches_oracle(oracle32, kind, strategy, filt):
    props = medium_props(stepsize=1e-2, strategy=strategy, sigmaS=(3.6, 3.0, 2.4), sigmaA=(0.4, 0.5, 0.6))
    med, omed, keep = setup(oracle32, kind, 40, props)
    scene = scene_dict(48, 40, 8, rfilter=filt)
    integ = mer.EikonalVolPathIntegrator(maxDepth=-1, rrDepth=5, stepsPerPass=256, poolPaths=4096)
    film, stats = integ.render(scene, med)
    ofilm, ostats = oracle32.render(omed, oracle_render_desc(scene))
    assert stats["samples"] == 48 * 40 * 8 == ostats.samples
    assert stats["nonfinite_dropped"] == 0
    relmse, frac_tight = image_gates(film, ofilm, 8)
    # same streams => almost every pixel agrees to 1e-3; event counts agree to a fraction of a percent
    assert frac_tight > 0.97 and relmse < 1e-3, (relmse, frac_tight)
    for k, ok in (("ray_steps", ostats.ray_steps), ("scatter_events", ostats.scatter_events),
                  ("boundary_exits", ostats.boundary_exits)):
        assert abs(stats[k] - ok) <= 0.005 * ok + 5, (k, stats[k], ok)
    assert stats["passes"] > 3  # the wavefront really ran in several compacted passes


def test_woodcock_density_image_matches_oracle(oracle32):
    props = medium_props(stepsize=1e-2, albedo=(0.9, 0.8, 0.7), densityScale=8.0)
    med, omed, keep = setup(oracle32, "radial", 40, props, density=True)
    scene = scene_dict(40, 40, 8, rfilter="box")
    film, stats = mer.EikonalVolPathIntegrator(stepsPerPass=300, poolPaths=2048).render(scene, med)
    ofilm, ostats = oracle32.render(omed, oracle_render_desc(scene))
    relmse, frac_tight = image_gates(film, ofilm, 8)
    assert frac_tight > 0.97 and relmse < 1e-3, (relmse, frac_tight)
    assert stats["null_collisions"] > 0
    for k, ok in (("ray_steps", ostats.ray_steps), ("scatter_events", ostats.scatter_events),
                  ("null_collisions", ostats.null_collisions)):
        assert abs(stats[k] - ok) <= 0.005 * ok + 5, (k, stats[k], ok)


def _albedo_field(res):
    """a smooth RGB albedo in [0.2, 1] with different structure per channel, [z, y, x, 3]"""
    x, y, z = [np.linspace(-1, 1, n, dtype=np.float32) for n in res]
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    return np.stack([0.6 + 0.4 * np.sin(3 * X), 0.6 + 0.4 * np.cos(2 * Y + 1), 0.6 + 0.4 * np.sin(2 * Z + X)], axis=-1).astype(np.float32)


def test_albedo_grid_image_matches_oracle(oracle32, tmp_path):
    """<volume name="albedo"> (heterogeneous.cpp:262-268): scattering events of the Woodcock walk weigh the path by
    lookupSpectrum(p) (:646-649).  The image and the event counts follow the oracle on shared streams; a constant albedo
    grid renders the constant-albedo image; a 3-channel .vol file gives the same medium as the array"""
    props = medium_props(stepsize=1e-2, albedo=(0.9, 0.8, 0.7), densityScale=8.0)
    med, omed, keep = setup(oracle32, "radial", 40, props, density=True)
    scene = scene_dict(40, 40, 8, rfilter="box")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=300, poolPaths=2048)
    film_const, stats_const = integ.render(scene, med)

    ares = (24, 20, 28)
    rgb = _albedo_field(ares)
    agrid = mer.GridDataSource(data=rgb, min=BOX_MIN, max=BOX_MAX)
    assert agrid.supportsSpectrumLookups() and not agrid.supportsFloatLookups()
    med2, omed2, keep2 = setup(oracle32, "radial", 40, props, density=True)
    med2.setAlbedoVolume(agrid)
    oracle32.medium_set_albedo_grid(omed2, oracle32.grid_create_spectrum(volume_desc(ares, BOX_MIN, BOX_MAX), rgb))
    film, stats = integ.render(scene, med2)
    ofilm, ostats = oracle32.render(omed2, oracle_render_desc(scene))
    relmse, frac_tight = image_gates(film, ofilm, 8)
    assert frac_tight > 0.97 and relmse < 1e-3, (relmse, frac_tight)
    for k, ok in (("ray_steps", ostats.ray_steps), ("scatter_events", ostats.scatter_events), ("null_collisions", ostats.null_collisions)):
        assert abs(stats[k] - ok) <= 0.005 * ok + 5, (k, stats[k], ok)
    assert np.abs(film[..., :3] - film_const[..., :3]).max() > 1e-3  # the albedo grid does change the image

    # through addChild("albedo") + configure(), from a 3-channel .vol file
    mer.fields.write_vol(tmp_path / "albedo.vol", rgb, BOX_MIN, BOX_MAX)
    fgrid = mer.GridDataSource(filename=str(tmp_path / "albedo.vol"))
    assert fgrid.channels == 3
    data, lo, hi = make_field("radial", 40)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    dens = mer.GridDataSource(data=mer.fields.sine_density((32, 32, 32), BOX_MIN, BOX_MAX), min=BOX_MIN, max=BOX_MAX)
    med3 = (mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9))
            .addChild("density", dens).addChild("albedo", fgrid).configure())
    film3, _ = integ.render(scene, med3)
    assert np.allclose(film3, film, rtol=2e-5, atol=2e-5)  # the same paths; splats are atomic adds in a free order

    # a constant grid is the constant albedo: lerps of equal values are exact up to (1 - f) + f rounding, so compare closely
    cgrid = mer.GridDataSource(data=np.broadcast_to(np.array((0.9, 0.8, 0.7), np.float32), ares[::-1] + (3,)).copy(), min=BOX_MIN, max=BOX_MAX)
    med2.setAlbedoVolume(cgrid)
    film_c, _ = integ.render(scene, med2)
    assert np.allclose(film_c, film_const, rtol=5e-5, atol=2e-5)
    med2.setAlbedoVolume(None)  # detached: the constant again
    film_d, stats_d = integ.render(scene, med2)
    assert np.allclose(film_d, film_const, rtol=2e-5, atol=2e-5) and stats_d["ray_steps"] == stats_const["ray_steps"]

    # validation: an albedo volume needs spectrum lookups and a density volume (heterogeneous.cpp:229-236, 266-268)
    with pytest.raises(mer.MerError, match="spectrum lookups"):
        med2.setAlbedoVolume(keep2[1])
    plain, _, _ = setup(oracle32, "radial", 40, medium_props(stepsize=1e-2))
    with pytest.raises(mer.MerError, match="density volume"):
        plain.setAlbedoVolume(agrid)
    with pytest.raises(mer.MerError, match="one channel"):
        mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("density", agrid).configure()


def test_max_depth_and_rr(oracle32):
    props = medium_props(stepsize=2e-2, sigmaS=6.0, sigmaA=0.0)
    med, omed, keep = setup(oracle32, "radial", 32, props, g=0.0)
    scene = scene_dict(32, 32, 4, rfilter="box", quad=False)
    for md, rr in ((3, 5), (8, 2), (-1, 1)):
        film, stats = mer.EikonalVolPathIntegrator(maxDepth=md, rrDepth=rr, stepsPerPass=128, poolPaths=1024).render(scene, med)
        ofilm, ostats = oracle32.render(omed, oracle_render_desc(scene, max_depth=md, rr_depth=rr))
        relmse, frac_tight = image_gates(film, ofilm, 4)
        assert frac_tight > 0.95, (md, rr, frac_tight)
        assert abs(stats["scatter_events"] - ostats.scatter_events) <= 0.01 * ostats.scatter_events + 5


def test_energy_conservation_white_furnace():
    """no absorption, unit environment, constant index => every pixel is exactly 1: with the default
    mediumSamplingWeight (= albedo = 1, heterogeneousrefractive.cpp:239-255) the weight sigma_s T / pdf of
    every edge is exactly 1, as are refRatioSq and the RR factor"""
    res = 24
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    data = np.full((res, res, res), 1.33, np.float32)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=2e-2, sigmaS=4.0, sigmaA=0.0))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.5)).configure()
    scene = scene_dict(32, 32, 4, rfilter="box", quad=False)
    film, stats = mer.EikonalVolPathIntegrator(rrDepth=1000, stepsPerPass=200, poolPaths=2048).render(scene, med)
    rgb = mer.develop(film)
    assert np.allclose(rgb, 1.0, atol=2e-4), (rgb.min(), rgb.max())
    assert np.allclose(film[..., 4], film[0, 0, 4])


def test_sample_sharding_adds_up(oracle32):
    """§8e: sample s -> GPU s mod G; partial films add to the single-GPU film (FP32 summation order only)"""
    med, omed, keep = setup(oracle32, "radial", 32, medium_props(stepsize=2e-2))
    scene = scene_dict(32, 24, 8, rfilter="gaussian")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=128, poolPaths=1024)
    full, st = integ.render(scene, med)
    parts = [integ.render(scene, med, sample_begin=r, sample_stride=3) for r in range(3)]
    total = sum(p[0] for p in parts)
    assert sum(p[1]["samples"] for p in parts) == st["samples"]
    assert sum(p[1]["ray_steps"] for p in parts) == st["ray_steps"]  # identical paths, just regrouped
    assert np.allclose(total, full, rtol=1e-5, atol=1e-5)
    # pool/pass sizes are scheduling knobs only
    other, st2 = mer.EikonalVolPathIntegrator(stepsPerPass=37, poolPaths=128).render(scene, med)
    assert st2["ray_steps"] == st["ray_steps"] and np.allclose(other, full, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("env", [{"MER_STEP_CTAS": "5", "MER_STEP_TPB": "32"}, {"MER_STEP_CTAS": "2", "MER_STEP_TPB": "96"}, {"MER_STEP_TUNE": "1"}])
def test_step_kernel_configurations_render_the_same_paths(oracle32, env, monkeypatch):
    """the step kernel's lanes per SM (the narrow configuration the in-run tuner picks for tables far larger than the L2) are
    scheduling only: same paths, same step and block counts, same film"""
    med, omed, keep = setup(oracle32, "radial", 32, medium_props(stepsize=2e-2))
    scene = scene_dict(48, 40, 16, rfilter="gaussian")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=16, poolPaths=4096)  # short visits: many rounds, so the tuner probes
    full, st = integ.render(scene, med)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    other, st2 = integ.render(scene, med)
    assert st["step_lanes_per_sm"] == 512 and (st2["step_lanes_per_sm"] == int(env.get("MER_STEP_CTAS", 0)) * int(env.get("MER_STEP_TPB", 0)) or "MER_STEP_TUNE" in env)
    for k in ("samples", "ray_steps", "block_fetches", "scatter_events", "boundary_exits"):
        assert st2[k] == st[k], k
    assert np.allclose(other, full, rtol=1e-5, atol=1e-5)


def test_render_multi_two_shards_on_one_device(oracle32):
    """mer_render_multi (the scheduler behind Integrator::render, integrator.cpp:95-127 / renderproc.cpp:142-148): two
    shards of the sample indices rendered by two host threads and reduced — here both on the one GPU a test box has,
    which takes the peer-copy + add path of the reduce — give the single-call film"""
    med, omed, keep = setup(oracle32, "radial", 32, medium_props(stepsize=2e-2))
    scene = scene_dict(32, 24, 8, rfilter="gaussian")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=128, poolPaths=1024)
    full, st = integ.render(scene, med)
    both, st2 = integ.render_multi(scene, [med, med])
    assert st2["samples"] == st["samples"] and st2["ray_steps"] == st["ray_steps"]
    assert np.allclose(both, full, rtol=1e-5, atol=1e-5)
    one, st1 = integ.render_multi(scene, [med])
    assert st1["ray_steps"] == st["ray_steps"] and np.allclose(one, full, rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(mer.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("nccl", ["1", "0"])
def test_render_multi_two_gpus(oracle32, nccl, monkeypatch):
    """one call drives two GPUs (grids replicated, sample indices interleaved, ncclReduce of the film — or, with
    MER_NCCL=0, peer copies) and the film matches the one-GPU film to 1e-5"""
    import torch  # noqa: F401  the library reduces with the NCCL the process already has; without this import it would load the
    # system's (older) libnccl.so.2, and a torch imported by a LATER test would be handed that one (same SONAME)
    monkeypatch.setenv("MER_NCCL", nccl)
    props = medium_props(stepsize=2e-2)
    data, lo, hi = make_field("radial", 32)
    media, keep = [], []
    for dev in (0, 1):
        rif = mer.SplineDataSource(data=data, min=lo, max=hi, device=dev)
        media.append(mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).configure())
        keep.append(rif)
    scene = scene_dict(40, 32, 16, rfilter="gaussian")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=128, poolPaths=2048)
    full, st = integ.render(scene, media[0])
    both, st2 = integ.render_multi(scene, media)
    assert st2["samples"] == st["samples"] and st2["ray_steps"] == st["ray_steps"]
    assert np.allclose(both, full, rtol=1e-5, atol=1e-5)


def test_packed_mode_renders_close_to_tricubic(oracle32):
    props = medium_props(stepsize=1e-2)
    med_c, _, k1 = setup(oracle32, "radial", 64, props, mode="tricubic")
    med_p, _, k2 = setup(oracle32, "radial", 64, props, mode="trilinear_packed")
    scene = scene_dict(32, 32, 64, rfilter="box")
    integ = mer.EikonalVolPathIntegrator(stepsPerPass=256, poolPaths=8192)
    a = mer.develop(integ.render(scene, med_c)[0])
    b = mer.develop(integ.render(scene, med_p)[0])
    # different interpolant (R1): images agree statistically, not bitwise
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02


# ---------------------------------------------------------------------------------------------------------
# next-row 3: index-mismatched container, <bsdf type="hdielectric"> (src/bsdfs/hdielectric.cpp)

@pytest.mark.parametrize("kind,shape", [("radial", ("box", BOX_MIN, BOX_MAX)), ("sd", ("sphere", (0.0, 0.0, 0.0), 0.95))])
def test_hdielectric_boundary_image_matches_oracle(oracle32, kind, shape):
    props = medium_props(stepsize=1e-2, sigmaS=(1.8, 1.5, 1.2), sigmaA=(0.2, 0.25, 0.3), bsdf="hdielectric", shape=shape)
    med, omed, keep = setup(oracle32, kind, 40, props, g=0.6)
    scene = scene_dict(48, 40, 8, rfilter="box")
    integ = mer.EikonalVolPathIntegrator(maxDepth=-1, rrDepth=5, stepsPerPass=256, poolPaths=4096)
    film, stats = integ.render(scene, med)
    ofilm, ostats = oracle32.render(omed, oracle_render_desc(scene))
    assert stats["samples"] == ostats.samples and stats["nonfinite_dropped"] == 0
    relmse, frac_tight = image_gates(film, ofilm, 8)
    assert frac_tight > 0.95 and relmse < 2e-3, (relmse, frac_tight)
    for k, ok in (("ray_steps", ostats.ray_steps), ("scatter_events", ostats.scatter_events),
                  ("boundary_exits", ostats.boundary_exits)):
        assert abs(stats[k] - ok) <= 0.01 * ok + 5, (k, stats[k], ok)
    # the boundary matters: internal reflection sends paths back in, so there are more surface events than samples
    matched = mer.HeterogeneousRefractiveMedium(dict(props, bsdf="null")).addChild("rif", keep[0]).addChild("", mer.HGPhaseFunction(g=0.6))
    matched.configure()
    _, st0 = integ.render(scene, matched)
    assert stats["boundary_exits"] > 1.02 * st0["boundary_exits"]


def test_hdielectric_white_furnace():
    """constant index 1.5 behind a Fresnel boundary, no absorption, unit environment: reflection has weight 1 and the
    1/eta^2 of entering cancels the eta^2 of leaving, so every pixel is exactly 1 whatever the path does"""
    res = 24
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    rif = mer.SplineDataSource(data=np.full((res,) * 3, 1.5, np.float32), min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=2e-2, sigmaS=2.0, sigmaA=0.0, bsdf="hdielectric"))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.3)).configure()
    scene = scene_dict(32, 32, 8, rfilter="box", quad=False)
    film, stats = mer.EikonalVolPathIntegrator(rrDepth=1000, stepsPerPass=200, poolPaths=2048).render(scene, med)
    rgb = mer.develop(film)
    assert np.allclose(rgb, 1.0, atol=3e-4), (rgb.min(), rgb.max())


def test_hdielectric_fresnel_reflectance():
    """a non-scattering slab of constant index seen head-on against a black environment with a bright quad BEHIND THE
    CAMERA: the only light a pixel can receive is the mirror reflection off the front face plus the (multiply
    reflected) light re-emerging towards the camera side, all of which see the quad; a camera ray that transmits through
    both faces sees black.  Expected pixel value = R_total = 2R/(1+R) for normal incidence, R = ((n-1)/(n+1))^2."""
    n = 1.5
    res = 16
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    rif = mer.SplineDataSource(data=np.full((res,) * 3, n, np.float32), min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=5e-2, sigmaS=0.0, sigmaA=0.0, mediumSamplingWeight=0.0,
                                                         bsdf="hdielectric"))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.0)).configure()
    scene = scene_dict(8, 8, 4096, rfilter="box", quad=False)
    scene.update(fov=2.0, envRadiance=0.0, quad=dict(origin=(-50.0, -50.0, -6.0), u=(100.0, 0.0, 0.0), v=(0.0, 100.0, 0.0), radiance=(1.0, 1.0, 1.0)))
    film, stats = mer.EikonalVolPathIntegrator(rrDepth=1000, stepsPerPass=64, poolPaths=4096).render(scene, med)
    rgb = mer.develop(film)
    R = ((n - 1) / (n + 1)) ** 2
    expect = 2 * R / (1 + R)
    assert abs(rgb.mean() - expect) < 4 * np.sqrt(expect * (1 - expect) / (64 * 4096)) + 1e-3, (rgb.mean(), expect)


# ---------------------------------------------------------------------------------------------------------
# next-row 1: next-event estimation along curved connections inside the bounce loop

def _nee_medium(kind, h, bsdf, res=32, g=0.5, sigmaS=1.5, sigmaA=0.5, shape=("box", BOX_MIN, BOX_MAX)):
    if kind == "const":
        lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
        data = np.full((res,) * 3, 1.5, np.float32)
    else:
        data, lo, hi = make_field(kind, res)
    props = medium_props(stepsize=h, sigmaS=sigmaS, sigmaA=sigmaA, bsdf=bsdf, shape=shape)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=g)).configure()
    return med, rif, props, data, lo, hi


@pytest.mark.parametrize("kind,bsdf", [("linear", "null"), ("radial", "hdielectric")])
def test_direct_connections_match_oracle(oracle32, kind, bsdf):
    """the wavefront (request queue + k_nee) against the oracle's inline restatement of the same estimator: the walk and
    the solver draw from the same Philox streams, so connections, failures and pixels agree one by one"""
    med, rif, props, data, lo, hi = _nee_medium(kind, 1e-2, bsdf)
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data))
    scene = scene_dict(32, 32, 8, rfilter="box")
    scene["envRadiance"] = 0.25
    film, st = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections=True, poolPaths=512, stepsPerPass=64).render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections=True, props=props))
    assert st["connections"] > 5000 and abs(st["connections"] - ost.connections) <= 0.005 * ost.connections
    assert abs(st["connections_failed"] - ost.connections_failed) <= 0.02 * ost.connections + 3
    assert abs(st["connection_steps"] - ost.connection_steps) <= 0.01 * ost.connection_steps
    assert st["nonfinite_dropped"] == 0 and st["kernel_launches"] > st["passes"]  # k_nee ran between the passes
    a, b = mer.develop(film), oracle32.film_develop(ofilm)
    assert np.mean(np.abs(a - b) <= 2e-3 * np.maximum(b, 1.0)) > 0.97
    assert abs(a.mean() - b.mean()) <= 2e-3 * b.mean()


@pytest.mark.parametrize("kind,bsdf,shape", [("const", "null", ("box", BOX_MIN, BOX_MAX)), ("linear", "null", ("box", BOX_MIN, BOX_MAX)),
                                             ("sd", "hdielectric", ("box", BOX_MIN, BOX_MAX)),
                                             ("radial", "null", ("sphere", (0.0, 0.0, 0.0), 0.9))])
def test_direct_connections_agree_with_random_walk(kind, bsdf, shape):
    """the physics check SURVEY 8f-1 asks for: with and without direct connections the image has the same expectation.
    The quad's light is estimated either by hitting it (walk) or by solving the curved connection from every scattering
    vertex (and then no longer counted on hits).  The walk's exit edges are ~0.75 h short (heterogeneousrefractive.cpp
    :742-776 quirk), which makes it brighter by about 0.75 sigma_t h; that shift is allowed for."""
    h, sigma_t = 2.5e-3, 2.0
    med = _nee_medium(kind, h, bsdf, shape=shape)[0]
    mean = {}
    for nee in (False, True):
        vals = []
        for seed in range(1, 9):
            scene = scene_dict(128, 128, 16, rfilter="box", seed=seed)
            scene["envRadiance"] = 0.0
            film, st = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections=nee).render(scene, med)
            vals.append(float(mer.develop(film)[..., 0].mean()))
        mean[nee] = (np.mean(vals), np.std(vals, ddof=1) / np.sqrt(len(vals)))
    assert st["connections"] > 1e5 and st["connections_failed"] < (0.02 if bsdf == "null" else 0.2) * st["connections"]
    ratio = mean[True][0] / mean[False][0] - 1 + 0.75 * sigma_t * h
    sigma = np.hypot(mean[True][1], mean[False][1]) / mean[False][0]
    assert abs(ratio) < 4 * sigma + 4e-3, (kind, bsdf, ratio, sigma, mean)
    assert mean[True][1] < mean[False][1]  # and it is the lower-variance estimator


@pytest.mark.parametrize("kind,bsdf", [("linear", "null"), ("radial", "hdielectric")])
def test_mis_connections_match_oracle(oracle32, kind, bsdf):
    """direct_connections = 2 against the oracle's restatement of volpath's power heuristic (volpath.cpp:120-147, 164-173,
    430-433) on shared Philox streams: the walk, the connections AND the weight requests of the emitter hits (solved for
    the hit point, second half of the vertex's NEE stream) agree one by one"""
    med, rif, props, data, lo, hi = _nee_medium(kind, 1e-2, bsdf)
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data))
    scene = scene_dict(32, 32, 8, rfilter="box")
    scene["envRadiance"] = 0.25
    film, st = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections="mis", poolPaths=512, stepsPerPass=64).render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections="mis", props=props))
    _, ost1 = oracle32.render(omed, oracle_render_desc(scene, direct_connections=True, props=props))
    assert ost.connections > ost1.connections + 20  # the hit requests are there
    assert abs(st["connections"] - ost.connections) <= 0.005 * ost.connections
    assert abs(st["connections_failed"] - ost.connections_failed) <= 0.02 * ost.connections + 3
    assert abs(st["connection_steps"] - ost.connection_steps) <= 0.01 * ost.connection_steps
    assert st["nonfinite_dropped"] == 0
    a, b = mer.develop(film), oracle32.film_develop(ofilm)
    assert np.mean(np.abs(a - b) <= 2e-3 * np.maximum(b, 1.0)) > 0.97
    assert abs(a.mean() - b.mean()) <= 2e-3 * b.mean()


@pytest.mark.parametrize("kind,bsdf,shape", [("linear", "null", ("box", BOX_MIN, BOX_MAX)), ("sd", "hdielectric", ("box", BOX_MIN, BOX_MAX))])
def test_mis_connections_agree_with_both_estimators(kind, bsdf, shape):
    """direct_connections = 2: volpath's power heuristic (volpath.cpp:120-147, 164-173, 430-433) between the curved
    connection and phase sampling.  A weighted combination of two unbiased estimators of the quad's light, so its
    expectation is theirs: against the un-MIS'd connections (same 0.75 sigma_t h exit-length quirk on the hit part only
    where hits count) and against the plain walk; and the hit requests really ran (more requests than scattering
    vertices ask for)."""
    h, sigma_t = 2.5e-3, 2.0
    med = _nee_medium(kind, h, bsdf, shape=shape)[0]
    mean, conn = {}, {}
    for mode in (False, True, "mis"):
        vals = []
        for seed in range(1, 9):
            scene = scene_dict(128, 128, 16, rfilter="box", seed=seed)
            scene["envRadiance"] = 0.0
            film, st = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections=mode).render(scene, med)
            vals.append(float(mer.develop(film)[..., 0].mean()))
        mean[mode] = (np.mean(vals), np.std(vals, ddof=1) / np.sqrt(len(vals)))
        conn[mode] = st["connections"]
    assert conn["mis"] > conn[True] > 1e5  # the emitter hits of covered chains asked for their weight
    lo_, hi_ = sorted([mean[False][0], mean[True][0]])
    slack = 4 * max(mean[False][1], mean[True][1], mean["mis"][1]) + 4e-3 * hi_
    assert lo_ - slack <= mean["mis"][0] <= hi_ + slack, (kind, bsdf, mean)
    assert mean["mis"][1] < mean[False][1]  # far less noisy than finding the quad by hitting it


def test_direct_connections_queue_and_validation(oracle32):
    med, rif, props, data, lo, hi = _nee_medium("linear", 1e-2, "null")
    scene = scene_dict(32, 32, 8, rfilter="box")
    big, st_big = mer.EikonalVolPathIntegrator(directConnections=True).render(scene, med)
    # 128 path slots => a 256-entry request queue that overflows in every pass: vertices wait for the next pass
    small, st_small = mer.EikonalVolPathIntegrator(directConnections=True, poolPaths=128, stepsPerPass=512).render(scene, med)
    assert st_small["connections"] == st_big["connections"] and st_small["ray_steps"] == st_big["ray_steps"]
    assert st_small["passes"] > st_big["passes"] and st_small["connections_failed"] == st_big["connections_failed"]
    assert np.allclose(small, big, rtol=1e-5, atol=1e-5 * big.max())  # same requests whatever the queue order: FP32 add order only
    again, _ = mer.EikonalVolPathIntegrator(directConnections=True, poolPaths=128, stepsPerPass=512).render(scene, med)
    assert np.allclose(again, small, rtol=1e-5, atol=1e-5 * big.max())
    # the random first guess of the reference is available and statistically equivalent, only more expensive
    rnd, st_rnd = mer.EikonalVolPathIntegrator(directConnections=True, connectionStart="random").render(scene, med)
    assert st_rnd["connection_steps"] > 1.5 * st_big["connection_steps"]
    assert abs(mer.develop(rnd).mean() / mer.develop(big).mean() - 1) < 0.1
    with pytest.raises(mer.MerError, match="quad"):
        mer.EikonalVolPathIntegrator(directConnections=True).render(scene_dict(16, 16, 1, quad=False), med)
    packed = mer.SplineDataSource(data=data, min=lo, max=hi, mode="trilinear_packed")
    pm = mer.HeterogeneousRefractiveMedium(props).addChild("rif", packed).configure()
    with pytest.raises(mer.MerError, match="tricubic"):
        mer.EikonalVolPathIntegrator(directConnections=True).render(scene, pm)


def test_direct_connections_through_density_grid(oracle32):
    """Woodcock medium (density grid) + direct connections: the connection's transmittance is exp(-optical depth) along the
    curve.  Parity with the oracle on shared streams, and the same expectation as the walk without connections."""
    data, lo, hi = make_field("linear", 32)
    dres = (24,) * 3
    dens = mer.fields.sine_density(dres, BOX_MIN, BOX_MAX)
    props = medium_props(stepsize=5e-3, albedo=0.8, densityScale=3.0)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    grid = mer.GridDataSource(data=dens, min=BOX_MIN, max=BOX_MAX)
    med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("density", grid).addChild("", mer.HGPhaseFunction(g=0.5)).configure()
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5, has_density=True), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data),
                                  oracle32.grid_create(volume_desc(dres, BOX_MIN, BOX_MAX), dens))
    scene = scene_dict(32, 32, 4, rfilter="box")
    scene["envRadiance"] = 0.0
    film, st = mer.EikonalVolPathIntegrator(directConnections=True, poolPaths=2048, stepsPerPass=256).render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections=True, props=props))
    assert st["connections"] > 2000 and abs(st["connections"] - ost.connections) <= 0.01 * ost.connections and st["null_collisions"] > 0
    a, b = mer.develop(film), oracle32.film_develop(ofilm)
    assert np.mean(np.abs(a - b) <= 2e-3 * np.maximum(b, 1.0)) > 0.97 and abs(a.mean() - b.mean()) <= 3e-3 * b.mean()
    mean = {}
    for nee in (False, True):
        vals = []
        for seed in range(1, 9):
            sc = scene_dict(128, 128, 16, rfilter="box", seed=seed)
            sc["envRadiance"] = 0.0
            vals.append(float(mer.develop(mer.EikonalVolPathIntegrator(directConnections=nee).render(sc, med)[0])[..., 0].mean()))
        mean[nee] = (np.mean(vals), np.std(vals, ddof=1) / np.sqrt(len(vals)))
    ratio = mean[True][0] / mean[False][0] - 1
    sigma = np.hypot(mean[True][1], mean[False][1]) / mean[False][0]
    assert abs(ratio) < 4 * sigma + 4e-3, (ratio, sigma, mean)


# ---------------------------------------------------------------------------------------------------------
# next-row 2, film half: transient (path-length resolved) film

@pytest.mark.parametrize("kind,bsdf,nee", [("linear", "null", False), ("radial", "hdielectric", True)])
def test_transient_film_matches_oracle(oracle32, kind, bsdf, nee):
    med, rif, props, data, lo, hi = _nee_medium(kind, 1e-2, bsdf)
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data))
    scene = scene_dict(32, 24, 8, rfilter="gaussian")
    scene["envRadiance"] = 0.0
    integ = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections=nee, poolPaths=1024, stepsPerPass=128)
    steady, st0 = integ.render(scene, med)
    scene["transient"] = dict(minBound=4.0, maxBound=36.0, binWidth=0.5)
    film, st = integ.render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections=nee, props=props))
    assert film.shape == ofilm.shape == (24, 32, 3 * 64 + 2)
    assert st["ray_steps"] == st0["ray_steps"] and st["samples"] == ost.samples
    # the frames add up to the steady-state film of the same samples (next to nothing is longer than 36: internal reflections)
    rgb = film[..., :-2].reshape(24, 32, 64, 3)
    assert np.allclose(rgb.sum(axis=2), steady[..., :3], rtol=2e-3, atol=1e-5)
    assert rgb.sum() <= steady[..., :3].sum() * (1 + 1e-5)
    assert np.allclose(film[..., -2:], steady[..., 3:], rtol=1e-5, atol=1e-6)
    # against the oracle: the time profile of the whole image, and frame by frame where there is signal
    prof_g, prof_c = rgb.sum(axis=(0, 1, 3)), ofilm[..., :-2].reshape(24, 32, 64, 3).sum(axis=(0, 1, 3))
    assert (prof_c > 0).sum() >= 8
    assert np.allclose(prof_g, prof_c, rtol=0.02, atol=2e-3 * prof_c.max())
    a, b = mer.develop(film), oracle32.film_develop(ofilm)
    assert a.shape == b.shape == (24, 32, 64, 3)
    assert np.mean(np.abs(a - b) <= 2e-3 * np.maximum(b, 1.0)) > 0.995


def test_transient_film_slab_time_of_flight():
    """a clear slab of index 1.5 in front of a bright wall: camera -> slab 3, inside 2 x 1.5, slab -> wall 2.  All light
    arrives at optical length 8 (the walk stops up to 2 h short of the far face); `calibrated` drops the camera segment."""
    res = 16
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    rif = mer.SplineDataSource(data=np.full((res,) * 3, 1.5, np.float32), min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=1e-2, sigmaS=0.0, sigmaA=0.0, mediumSamplingWeight=0.0))
    med.addChild("rif", rif).configure()
    scene = scene_dict(8, 8, 64, rfilter="box", quad=False)
    scene.update(fov=2.0, envRadiance=0.0, quad=dict(origin=(-50.0, -50.0, 3.0), u=(100.0, 0.0, 0.0), v=(0.0, 100.0, 0.0), radiance=(1.0, 1.0, 1.0)),
                 transient=dict(minBound=7.0, maxBound=9.0, binWidth=0.125))
    frames = mer.develop(mer.EikonalVolPathIntegrator().render(scene, med)[0])[..., 0].mean(axis=(0, 1))
    assert frames.shape == (16,) and abs(frames[7] + frames[8] - 1.0) < 1e-4 and frames[:7].sum() == 0 and frames[9:].sum() == 0
    scene["transient"].update(minBound=4.0, maxBound=6.0, calibrated=True)
    frames = mer.develop(mer.EikonalVolPathIntegrator().render(scene, med)[0])[..., 0].mean(axis=(0, 1))
    assert abs(frames[7] + frames[8] - 1.0) < 1e-4
    # steady state through the same call: one frame, five channels
    del scene["transient"]
    film, _ = mer.EikonalVolPathIntegrator().render(scene, med)
    assert film.shape == (8, 8, 5) and np.allclose(mer.develop(film), 1.0, atol=1e-4)


def test_radiance_scaling_conventions(oracle32):
    """the fork's refRatioSq (default, parity) against its reciprocal (`radianceScaling="physical"`): only the latter makes a
    lossless GRIN medium behind a Fresnel boundary a white furnace; both match the oracle"""
    data, lo, hi = make_field("linear", 32)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    orif = oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data)
    scene = scene_dict(32, 32, 16, rfilter="box", quad=False)
    for scaling in ("physical", "reference"):
        props = medium_props(stepsize=1e-2, sigmaS=2.0, sigmaA=0.0, bsdf="hdielectric", radianceScaling=scaling)
        med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.3)).configure()
        rgb = mer.develop(mer.EikonalVolPathIntegrator(rrDepth=1000).render(scene, med)[0])
        omed = oracle32.medium_create(oracle_medium_desc(props, 0.3), orif)
        ref = oracle32.film_develop(oracle32.render(omed, oracle_render_desc(scene, rr_depth=1000))[0])
        assert np.mean(np.abs(rgb - ref) <= 2e-3 * np.maximum(ref, 1.0)) > 0.97
        if scaling == "physical":
            assert np.abs(rgb - 1.0).max() < 3e-3
        else:
            assert np.abs(rgb - 1.0).max() > 0.1


# ---------------------------------------------------------------------------------------------------------
# next-row 2, walk half: light tracing = emitter-side walk + t = 1 sensor connections

def _hidden_quad_scene(w, h, spp, seed=20201201, **kw):
    """fov 30 deg: the quad above the box is outside the camera's view, so every bit of light scatters in the medium first
    (light tracing towards a pinhole cannot sample an emitter that is seen directly)"""
    scene = scene_dict(w, h, spp, rfilter="box", seed=seed)
    scene.update(fov=30.0, envRadiance=0.0, **kw)
    return scene


@pytest.mark.parametrize("kind,bsdf,emitter", [("linear", "hdielectric", "quad"), ("const", "null", "collimated")])
def test_light_tracing_matches_oracle(oracle32, kind, bsdf, emitter):
    med, rif, props, data, lo, hi = _nee_medium(kind, 1e-2, bsdf)
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data))
    scene = _hidden_quad_scene(32, 32, 8, transient=dict(minBound=2.0, maxBound=18.0, binWidth=0.5))
    if emitter == "collimated":
        scene["emitter"] = dict(type="collimated", origin=(0.2, 0.1, -3.0), direction=(0.0, 0.0, 1.0), power=(30.0, 20.0, 10.0))
    # every path of a collimated beam starts on the same ray and lands on the same few pixels, so the handful of walks that
    # FP32 rounding sends different ways (0.06 % of them) would dominate a multiple-scattering comparison: single scattering there
    md = 4 if emitter == "collimated" else -1
    film, st = mer.EikonalVolPathIntegrator(maxDepth=md, rrDepth=5, lightTracing=True, poolPaths=1024, stepsPerPass=128).render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, max_depth=md, light_tracing=True, props=props))
    assert film.shape == ofilm.shape == (32, 32, 3 * 32 + 2)
    assert st["samples"] == ost.samples == 32 * 32 * 8 and st["connections"] > 3000
    assert abs(st["connections"] - ost.connections) <= 0.01 * ost.connections
    assert np.allclose(film[..., -1], 1.0) and np.allclose(ofilm[..., -1], 1.0)  # unit weight: develop() returns the splats' sum
    a, b = film[..., :-2].reshape(32, 32, 32, 3), ofilm[..., :-2].reshape(32, 32, 32, 3)
    assert abs(a.sum() / b.sum() - 1) < 5e-3
    # time profile and image (summed over frames) of the same light paths
    assert np.allclose(a.sum(axis=(0, 1, 3)), b.sum(axis=(0, 1, 3)), rtol=0.03, atol=3e-3 * b.sum(axis=(0, 1, 3)).max())
    ia, ib = a.sum(axis=(2, 3)), b.sum(axis=(2, 3))
    assert np.mean(np.abs(ia - ib) <= 0.02 * ib + 1e-3 * ib.max()) > 0.95


@pytest.mark.parametrize("kind,bsdf,scaling", [("const", "null", "reference"), ("linear", "hdielectric", "physical")])
def test_light_tracing_agrees_with_camera_tracer(kind, bsdf, scaling):
    """the two estimators of the same image: camera walk (+ direct connections to the quad) and emitter walk (+ sensor
    connections).  Importance transport carries no index factors, so it is the physically scaled camera estimator
    (radianceScaling="physical") that it has to agree with once the index varies."""
    if kind == "const":
        lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (32,) * 3)
        data = np.full((32,) * 3, 1.5, np.float32)
    else:
        data, lo, hi = make_field(kind, 32)
    props = medium_props(stepsize=5e-3, sigmaS=1.5, sigmaA=0.5, bsdf=bsdf, radianceScaling=scaling)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.5)).configure()
    imgs = {}
    for name, kw in (("camera", dict(directConnections=True)), ("light", dict(lightTracing=True))):
        runs = []
        for seed in range(1, 5):
            film, st = mer.EikonalVolPathIntegrator(rrDepth=5, **kw).render(_hidden_quad_scene(64, 64, 64, seed=seed), med)
            runs.append(mer.develop(film)[..., 0])
        imgs[name] = np.array(runs)
    # blocks of 8x8 pixels, 4 independent runs each: means and their standard errors
    blk = {k: v.reshape(4, 8, 8, 8, 8).mean(axis=(2, 4)) for k, v in imgs.items()}
    mc, ml = blk["camera"].mean(axis=0), blk["light"].mean(axis=0)
    se = np.sqrt(blk["camera"].var(axis=0, ddof=1) / 4 + blk["light"].var(axis=0, ddof=1) / 4)
    assert abs(ml.mean() / mc.mean() - 1) < 0.015, (ml.mean(), mc.mean())
    bright = mc > 0.05 * mc.max()
    assert np.all(np.abs(ml - mc)[bright] <= 5 * se[bright] + 0.03 * mc[bright])
    assert np.corrcoef(mc.ravel(), ml.ravel())[0, 1] > 0.995


def test_light_tracing_collimated_beam_time_of_flight():
    """the fork's showcase configuration (collimated beam into a scattering refractive medium, transient film), with a
    closed-form bound: a beam along +z from z = -3 into a slab of index 1.5 behind z = -1, camera at z = -4.  Light that
    scatters at depth s has travelled 2 + 1.5 s and returns over at least 1.5 s + 3: nothing arrives before 5, and
    `calibratedTransient` (sensor connection left out) starts at 2."""
    res = 16
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, (res,) * 3)
    rif = mer.SplineDataSource(data=np.full((res,) * 3, 1.5, np.float32), min=lo, max=hi)
    med = mer.HeterogeneousRefractiveMedium(medium_props(stepsize=1e-2, sigmaS=1.0, sigmaA=0.2))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.0)).configure()
    scene = scene_dict(32, 32, 16, rfilter="box", quad=False)
    scene.update(envRadiance=0.0, emitter=dict(type="collimated", origin=(0.0, 0.0, -3.0), direction=(0.0, 0.0, 1.0), power=100.0),
                 transient=dict(minBound=0.0, maxBound=32.0, binWidth=0.25))
    integ = mer.EikonalVolPathIntegrator(lightTracing=True)
    film, st = integ.render(scene, med)
    prof = film[..., :-2].reshape(32, 32, 128, 3).sum(axis=(0, 1, 3))
    first = int(np.nonzero(prof > 0)[0][0])
    assert first == 20 and prof[20:24].sum() > 0.2 * prof.sum()  # 5 / 0.25; single scattering near the front face dominates
    scene["transient"]["calibrated"] = True
    cal = integ.render(scene, med)[0][..., :-2].reshape(32, 32, 128, 3).sum(axis=(0, 1, 3))
    assert int(np.nonzero(cal > 0)[0][0]) == 8  # 2 / 0.25
    del scene["transient"]
    steady = integ.render(scene, med)[0]
    assert steady.shape == (32, 32, 5) and abs(prof.sum() / steady[..., :3].sum() - 1) < 1e-3
    # the beam is seen as a streak through the centre of the image
    img = steady[..., 0]
    assert img[12:20, 12:20].sum() > 0.5 * img.sum()


def test_sample_sharding_adds_up_in_every_mode():
    """§8e for the widened path: sample s -> GPU s mod G also shards direct connections, light paths and transient frames;
    the partial films (each with its share of the light-traced unit weight) add up to the single-GPU film"""
    med = _nee_medium("linear", 1e-2, "hdielectric")[0]
    for kw, scene in ((dict(directConnections=True), _hidden_quad_scene(24, 24, 6, transient=dict(minBound=4.0, maxBound=20.0, binWidth=1.0))),
                      (dict(lightTracing=True), _hidden_quad_scene(24, 24, 6, transient=dict(minBound=2.0, maxBound=18.0, binWidth=1.0))),
                      (dict(lightTracing=True), _hidden_quad_scene(24, 24, 7))):
        integ = mer.EikonalVolPathIntegrator(rrDepth=5, poolPaths=512, stepsPerPass=128, **kw)
        full, st = integ.render(scene, med)
        parts = [integ.render(scene, med, sample_begin=r, sample_stride=3) for r in range(3)]
        total = sum(p[0] for p in parts)
        for k in ("samples", "ray_steps", "connections", "connections_failed"):
            assert sum(p[1][k] for p in parts) == st[k], k
        assert np.allclose(total, full, rtol=1e-5, atol=1e-5 * np.abs(full).max())


# ---------------------------------------------------------------------------------------------------------
# next-row 4: containment by a signed-distance grid

@pytest.mark.parametrize("bsdf,nee", [("null", False), ("hdielectric", True)])
def test_sdf_container_matches_oracle_and_the_analytic_sphere(oracle32, bsdf, nee):
    """shape ("sdf", bbox): inside <=> sdf(p) < 0, entry by sphere tracing, normals from the gradient, and a stepper that
    skips the containment lookup while the last one guarantees room.  With the signed distance of a sphere on a grid the
    result must be the analytic sphere's (up to the 1e-4 the entry point sits inside the surface), and the oracle's."""
    res = 48
    data, lo, hi = make_field("sd", res)
    sdf_data = mer.fields.sphere_sdf((res,) * 3, lo, hi, radius=0.8).astype(np.float32)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    sdf = mer.SplineDataSource(data=sdf_data, min=lo, max=hi)
    scene = scene_dict(40, 40, 8, rfilter="box")
    integ = mer.EikonalVolPathIntegrator(rrDepth=5, directConnections=nee, poolPaths=2048, stepsPerPass=128)
    out = {}
    for name, shape in (("sdf", ("sdf", BOX_MIN, BOX_MAX)), ("sphere", ("sphere", (0.0, 0.0, 0.0), 0.8))):
        props = medium_props(stepsize=5e-3, sigmaS=2.0, sigmaA=0.5, bsdf=bsdf, shape=shape)
        med = mer.HeterogeneousRefractiveMedium(props).addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.5))
        if name == "sdf":
            med.addChild("sdf", sdf)
        med.configure()
        film, st = integ.render(scene, med)
        out[name] = (mer.develop(film), st, props)
    a, b = out["sdf"][0], out["sphere"][0]
    for k in ("ray_steps", "scatter_events", "boundary_exits"):
        assert abs(out["sdf"][1][k] - out["sphere"][1][k]) <= 0.01 * out["sphere"][1][k], k
    assert abs(a.mean() / b.mean() - 1) < 0.01 and np.mean(np.abs(a - b) <= 0.02 * np.maximum(b, 1.0)) > 0.9
    # oracle with the same shape type
    d = volume_desc((res,) * 3, lo, hi)
    omed = oracle32.medium_create(oracle_medium_desc(out["sdf"][2], 0.5), oracle32.rif_create(d, data))
    oracle32.medium_set_sdf(omed, oracle32.rif_create(d, sdf_data), False)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections=nee, props=out["sdf"][2]))
    ref = oracle32.film_develop(ofilm)
    assert abs(out["sdf"][1]["ray_steps"] - ost.ray_steps) <= 0.005 * ost.ray_steps
    assert np.mean(np.abs(a - ref) <= 2e-3 * np.maximum(ref, 1.0)) > 0.95 and abs(a.mean() / ref.mean() - 1) < 3e-3
    with pytest.raises(mer.MerError, match="sdf"):
        mer.HeterogeneousRefractiveMedium(medium_props(shape=("sdf", BOX_MIN, BOX_MAX))).addChild("rif", rif).configure()


@pytest.mark.parametrize("modulation,mode", [("sine", "camera"), ("square", "camera+nee"), ("hamiltonian", "light")])
def test_cw_tof_modulation_matches_oracle(oracle32, modulation, mode):
    """continuous-wave ToF: contributions times PathLengthSampler::correlationFunction(path length), one frame
    (src/librender/pathlengthsampler.cpp:66-96, film.cpp:76-78) — camera walk, its direct connections, and light tracing"""
    med, rif, props, data, lo, hi = _nee_medium("linear", 1e-2, "null")
    omed = oracle32.medium_create(oracle_medium_desc(props, 0.5), oracle32.rif_create(volume_desc((32,) * 3, lo, hi), data))
    scene = _hidden_quad_scene(32, 32, 8, transient=dict(minBound=0.0, maxBound=50.0, binWidth=1.0, modulation=modulation, **{"lambda": 3.0, "phase": 30.0}))
    kw = dict(directConnections=mode == "camera+nee", lightTracing=mode == "light")
    film, st = mer.EikonalVolPathIntegrator(rrDepth=5, poolPaths=1024, stepsPerPass=128, **kw).render(scene, med)
    ofilm, ost = oracle32.render(omed, oracle_render_desc(scene, direct_connections=kw["directConnections"], light_tracing=kw["lightTracing"], props=props))
    assert film.shape == ofilm.shape == (32, 32, 5) and st["samples"] == ost.samples
    a, b = mer.develop(film), oracle32.film_develop(ofilm)
    assert np.abs(b).max() > 0
    if modulation != "hamiltonian":  # sine and square codes have both signs, the Hamiltonian code lives in [0, 1]
        assert b.min() < 0 < b.max()
    scale = np.abs(b).max()
    assert np.mean(np.abs(a - b) <= 0.02 * np.abs(b) + 2e-3 * scale) > 0.95
    assert abs(a.sum() - b.sum()) <= 0.01 * np.abs(b).sum()
    # and the unmodulated steady-state image of the same samples bounds it: |correlation| <= 1
    del scene["transient"]
    steady = mer.develop(mer.EikonalVolPathIntegrator(rrDepth=5, poolPaths=1024, stepsPerPass=128, **kw).render(scene, med)[0])
    assert np.all(np.abs(a) <= steady * (1 + 1e-4) + 1e-6)
