"""The compiled C++ host side (host/): Mitsuba scene XML -> plugin objects -> C ABI.
CPU tests use --dry-run (everything but device handles); the GPU test renders the scene with the CLI and
compares the film with the Python mirror driving the same C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

import mitsubaer_b200 as mer
from common import BOX_MAX, BOX_MIN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "mer_render")
SCENE = os.path.join(ROOT, "scenes", "eikonal_box.xml")


@pytest.fixture(scope="module")
def volumes(tmp_path_factory):
    d = tmp_path_factory.mktemp("vols")
    res = (40, 40, 40)
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, res)
    mer.fields.write_vol(d / "rif.vol", mer.fields.radial_rif(res, lo, hi), lo, hi)
    mer.fields.write_vol(d / "den.vol", mer.fields.sine_density((24, 24, 24), BOX_MIN, BOX_MAX), BOX_MIN, BOX_MAX)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "host")], check=True)
    return d, lo, hi


def run(args, **kw):
    return subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600, **kw)


def test_scene_xml_resolves_like_mitsuba(volumes):
    d, lo, hi = volumes
    out = run([SCENE, "-D", "rif=%s" % (d / "rif.vol"), "-Dspp=7", "-D", "g=0.5", "--dry-run"])
    assert out.returncode == 0, out.stderr
    s = json.loads(out.stdout)
    assert s["integrator"] == "ervolpath" and (s["width"], s["height"], s["spp"]) == (128, 128, 7)  # <default> + -D
    assert s["max_depth"] == 64 and s["rr_depth"] == 5 and s["filter"] == 1 and s["fov"] == 40
    assert s["cam_origin"] == [0, 0, -4] and s["cam_target"] == [0, 0, -3] and s["cam_up"] == [0, 1, 0]  # <lookat>
    assert s["env"] == [1, 1, 1] and s["has_quad"] == 1 and s["quad_radiance"] == [8, 6, 4]
    # rectangle [-1,1]^2 -> scale .5 -> rotate 90 deg about x -> translate y 1.5 (elements apply in document order)
    assert np.allclose(s["quad_origin"], [-0.5, 1.5, -0.5]) and np.allclose(s["quad_u"], [1, 0, 0]) and np.allclose(s["quad_v"], [0, 0, 1], atol=1e-12)
    m = s["medium"]
    assert np.allclose(m["sigma_s"], 3.6) and np.allclose(m["sigma_a"], 0.4) and m["strategy"] == 1 and m["hg_g"] == 0.5
    assert m["boundary"] == 0  # <bsdf type="null"/> on the container
    out = run([SCENE, "-D", "rif=%s" % (d / "rif.vol"), "-D", "bsdf=hdielectric", "--dry-run"])
    assert out.returncode == 0 and json.loads(out.stdout)["medium"]["boundary"] == 1, out.stderr
    out = run([SCENE, "-D", "rif=%s" % (d / "rif.vol"), "-D", "bsdf=dielectric", "--dry-run"])
    assert out.returncode == 1 and "hdielectric" in out.stderr  # a fixed-IOR boundary is not on this path
    assert m["stepsize"] == pytest.approx(0.002) and m["weight"] == -1 and m["shape_type"] == 0 and m["shape"] == [-1, -1, -1, 1, 1, 1]
    assert m["rif_res"] == [40, 40, 40] and np.allclose(m["rif_bbox"], list(lo) + list(hi))


def test_scene_xml_error_behaviour(volumes, tmp_path):
    d, lo, hi = volumes
    rif = "rif=%s" % (d / "rif.vol")
    out = run([SCENE, "--dry-run"])
    assert out.returncode == 1 and "undefined parameter" in out.stderr  # scenehandler.cpp:210-219
    text = open(SCENE).read()

    def variant(name, old, new):
        assert old in text
        p = tmp_path / name
        p.write_text(text.replace(old, new))
        return run([str(p), "-D", rif, "--dry-run"])

    out = variant("a.xml", '<volume name="rif" type="splinevolume">', '<volume name="sdf" type="splinevolume">')
    assert out.returncode == 1 and "No RIF specified!" in out.stderr
    out = variant("b.xml", '<string name="strategy" value="single"/>', '<string name="strategy" value="bogus"/>')
    assert out.returncode == 1 and "unknown sampling strategy" in out.stderr
    out = variant("c.xml", '<float name="g" value="$g"/>', '<float name="g" value="1.5"/>')
    assert out.returncode == 1 and "interval (-1, 1)" in out.stderr
    out = variant("d.xml", '<integrator type="ervolpath">', '<integrator type="bdpt">')
    assert out.returncode == 1 and "bdpt" in out.stderr
    out = variant("e.xml", '<float name="stepsize" value="$stepsize"/>', '<float name="stepsize" value="$stepsize"/><float name="stepsiz" value="1"/>')
    assert out.returncode == 1 and "unused property" in out.stderr
    out = variant("f.xml", '<integer name="maxDepth" value="64"/>', '<integer name="maxDepth" value="0"/>')
    assert out.returncode == 1 and "maxDepth" in out.stderr
    out = variant("g.xml", "</scene>", "")
    assert out.returncode == 1 and "XML parse error" in out.stderr
    out = run([SCENE, "-D", "rif=/nonexistent.vol", "--dry-run"])
    assert out.returncode == 1 and "cannot open volume file" in out.stderr


def test_density_child_and_sphere_container(volumes, tmp_path):
    d, lo, hi = volumes
    text = open(SCENE).read()
    text = text.replace('<volume name="rif" type="splinevolume">', '<volume name="density" type="gridvolume"><string name="filename" value="%s"/></volume>\n'
                        '<float name="scale" value="8"/><spectrum name="albedo" value="0.9, 0.8, 0.7"/>\n<volume name="rif" type="splinevolume">' % (d / "den.vol"))
    text = text.replace('<shape type="cube">', '<shape type="sphere"><point name="center" x="0.1" y="0" z="0"/><float name="radius" value="0.7"/>')
    p = tmp_path / "s.xml"
    p.write_text(text)
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 0, out.stderr
    m = json.loads(out.stdout)["medium"]
    assert m["has_density"] == 1 and m["density_scale"] == 8 and np.allclose(m["albedo"], [0.9, 0.8, 0.7])
    assert m["shape_type"] == 1 and np.allclose(m["shape"][:4], [0.1, 0, 0, 0.7])
    assert np.allclose(m["sigma_s"], 3.6 * 8)  # `scale` multiplies sigmaS / sigmaA (medium/materials.h)


def _albedo_scene(volumes, tmp_path):
    d, lo, hi = volumes
    x = np.linspace(-1, 1, 20, dtype=np.float32)
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    rgb = np.stack([0.6 + 0.4 * np.sin(3 * X), 0.6 + 0.4 * np.cos(2 * Y), 0.6 + 0.4 * np.sin(2 * Z)], axis=-1).astype(np.float32)
    mer.fields.write_vol(tmp_path / "albedo.vol", rgb, BOX_MIN, BOX_MAX)
    text = open(SCENE).read()
    text = text.replace('<volume name="rif" type="splinevolume">', '<volume name="density" type="gridvolume"><string name="filename" value="%s"/></volume>\n'
                        '<volume name="albedo" type="gridvolume"><string name="filename" value="%s"/></volume>\n'
                        '<float name="scale" value="8"/><spectrum name="albedo" value="0.9, 0.8, 0.7"/>\n<volume name="rif" type="splinevolume">'
                        % (d / "den.vol", tmp_path / "albedo.vol"))
    p = tmp_path / "albedo.xml"
    p.write_text(text)
    return p, rgb


def test_albedo_volume_child(volumes, tmp_path):
    """<volume name="albedo" type="gridvolume"> (heterogeneous.cpp:262-268): a 3-channel .vol; a density file in its place is refused"""
    d, lo, hi = volumes
    p, rgb = _albedo_scene(volumes, tmp_path)
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 0, out.stderr
    m = json.loads(out.stdout)["medium"]
    assert m["has_density"] == 1 and m["has_albedo_volume"] == 1
    bad = tmp_path / "bad.xml"
    bad.write_text(p.read_text().replace(str(tmp_path / "albedo.vol"), str(d / "den.vol")))
    out = run([str(bad), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 1 and "spectrum lookups" in out.stderr


@pytest.mark.gpu
def test_cli_render_with_albedo_volume_matches_python_mirror(volumes, tmp_path):
    d, lo, hi = volumes
    p, rgb = _albedo_scene(volumes, tmp_path)
    film_path = tmp_path / "film.bin"
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "-D", "spp=8", "-D", "width=48", "-D", "height=40", "-D", "stepsize=0.01", "--film", str(film_path)])
    assert out.returncode == 0, out.stderr
    stats = json.loads(out.stdout)
    film = np.fromfile(film_path, np.float32).reshape(40, 48, 5)
    rif = mer.SplineDataSource(filename=str(d / "rif.vol"))
    med = mer.HeterogeneousRefractiveMedium(sigmaS=3.6, sigmaA=0.4, scale=8.0, albedo=(0.9, 0.8, 0.7), stepsize=0.01, strategy="single", shape=("box", BOX_MIN, BOX_MAX))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).addChild("density", mer.GridDataSource(filename=str(d / "den.vol")))
    med.addChild("albedo", mer.GridDataSource(data=rgb, min=BOX_MIN, max=BOX_MAX)).configure()
    scene = dict(width=48, height=40, sampleCount=8, seed=20201201, origin=(0, 0, -4), target=(0, 0, -3), up=(0, 1, 0), fov=40.0,
                 rfilter="gaussian", envRadiance=1.0, quad=dict(origin=(-0.5, 1.5, -0.5), u=(1, 0, 0), v=(0, 0, 1), radiance=(8, 6, 4)))
    ref, st = mer.EikonalVolPathIntegrator(maxDepth=64, rrDepth=5).render(scene, med)
    assert stats["samples"] == st["samples"] == 48 * 40 * 8 and stats["ray_steps"] == st["ray_steps"] and st["null_collisions"] > 0
    assert np.allclose(film, ref, rtol=1e-5, atol=1e-5)  # same paths; only the order of the float atomics differs


@pytest.mark.gpu
def test_cli_render_matches_python_mirror(volumes, tmp_path):
    d, lo, hi = volumes
    film_path, pfm = tmp_path / "film.bin", tmp_path / "out.pfm"
    out = run([SCENE, "-D", "rif=%s" % (d / "rif.vol"), "-D", "spp=8", "-D", "width=64", "-D", "height=48", "-D", "stepsize=0.01",
               "-o", str(pfm), "--film", str(film_path)])
    assert out.returncode == 0, out.stderr
    stats = json.loads(out.stdout)
    film = np.fromfile(film_path, np.float32).reshape(48, 64, 5)
    rif = mer.SplineDataSource(filename=str(d / "rif.vol"))
    med = mer.HeterogeneousRefractiveMedium(sigmaS=3.6, sigmaA=0.4, stepsize=0.01, strategy="single", shape=("box", BOX_MIN, BOX_MAX))
    med.addChild("rif", rif).addChild("", mer.HGPhaseFunction(g=0.9)).configure()
    scene = dict(width=64, height=48, sampleCount=8, seed=20201201, origin=(0, 0, -4), target=(0, 0, -3), up=(0, 1, 0), fov=40.0,
                 rfilter="gaussian", envRadiance=1.0, quad=dict(origin=(-0.5, 1.5, -0.5), u=(1, 0, 0), v=(0, 0, 1), radiance=(8, 6, 4)))
    ref, st = mer.EikonalVolPathIntegrator(maxDepth=64, rrDepth=5).render(scene, med)
    assert stats["samples"] == st["samples"] == 64 * 48 * 8 and stats["ray_steps"] == st["ray_steps"]
    assert np.allclose(film, ref, rtol=1e-5, atol=1e-5)  # same paths; only the order of the float atomics differs
    hdr = open(pfm, "rb").read(16)
    assert hdr.startswith(b"PF\n64 48\n-1.0\n")


BEAM = os.path.join(ROOT, "scenes", "eikonal_beam_transient.xml")


def test_beam_scene_resolves(volumes):
    """scenes/eikonal_beam_transient.xml: collimated emitter (toWorld lookat -> origin, direction), lightTracing, transient film"""
    d, lo, hi = volumes
    out = run([BEAM, "-D", "rif=%s" % (d / "rif.vol"), "-D", "tRes=0.5", "--dry-run"])
    assert out.returncode == 0, out.stderr
    s = json.loads(out.stdout)
    assert s["light_tracing"] == 1 and s["emitter_type"] == 1 and s["has_quad"] == 0 and s["env"] == [0, 0, 0]
    assert s["frames"] == 16 and s["min_bound"] == 4 and s["bin_width"] == 0.5  # ceil((12 - 4) / 0.5), film.cpp:73
    assert s["medium"]["boundary"] == 1
    text = open(BEAM).read()
    p = d / "beam_camera.xml"
    p.write_text(text.replace('<boolean name="lightTracing" value="true"/>', ""))
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 1 and "lightTracing" in out.stderr  # a delta beam cannot be found by a camera walk


@pytest.mark.gpu
def test_cli_renders_beam_transient(volumes, tmp_path):
    d, lo, hi = volumes
    pfm, film_path = tmp_path / "beam.pfm", tmp_path / "beam.bin"
    out = run([BEAM, "-D", "rif=%s" % (d / "rif.vol"), "-D", "spp=4", "-D", "width=48", "-D", "height=48", "-D", "stepsize=0.01", "-D", "tRes=0.5",
               "-o", str(pfm), "--film", str(film_path)])
    assert out.returncode == 0, out.stderr
    film = np.fromfile(film_path, np.float32).reshape(48, 48, 3 * 16 + 2)
    assert np.allclose(film[..., -1], 1.0) and film[..., :-2].sum() > 0
    assert os.path.exists(tmp_path / "beam_0000.pfm") and os.path.exists(tmp_path / "beam_0015.pfm")
    prof = film[..., :-2].reshape(48, 48, 16, 3).sum(axis=(0, 1, 3))
    # beam enters at x = -1 after 2 units; the earliest return to the camera (z = -4) is a few units later: early frames are empty
    assert prof[0] == 0 and (prof > 0).sum() >= 4


def test_mesh_container_through_sdf_volume(volumes, tmp_path):
    """<shape type="obj"> cannot be read here, but with an <volume name="sdf"> child on the medium the signed-distance grid is
    the container; without it the scene is refused like any shape this path cannot hold"""
    d, lo, hi = volumes
    mer.fields.write_vol(d / "sdf.vol", mer.fields.sphere_sdf((40, 40, 40), lo, hi, radius=0.8).astype(np.float32), lo, hi)
    text = open(SCENE).read().replace('<shape type="cube">', '<shape type="obj"><string name="filename" value="bunny.obj"/>')
    p = tmp_path / "mesh.xml"
    p.write_text(text)
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 1 and "sdf" in out.stderr
    p.write_text(text.replace('<volume name="rif" type="splinevolume">',
                              '<volume name="sdf" type="splinevolume"><string name="filename" value="%s"/></volume>\n<volume name="rif" type="splinevolume">' % (d / "sdf.vol")))
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "--dry-run"])
    assert out.returncode == 0, out.stderr
    m = json.loads(out.stdout)["medium"]
    assert m["shape_type"] == 2 and np.allclose(m["shape"], list(lo) + list(hi))


def test_cw_tof_film_resolves(volumes, tmp_path):
    """<film> modulation / lambda / phase (PathLengthSampler, src/librender/pathlengthsampler.cpp:6-35): one frame, the code and
    its parameters in the descriptor; unknown codes are refused"""
    d, lo, hi = volumes
    text = open(BEAM).read().replace('<float name="binWidth" value="$tRes"/>',
                                     '<float name="binWidth" value="$tRes"/>\n<string name="modulation" value="$mod"/>\n'
                                     '<float name="lambda" value="3.5"/>\n<float name="phase" value="90"/>')
    p = tmp_path / "cw.xml"
    p.write_text(text)
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "-D", "mod=Sine", "--dry-run"])
    assert out.returncode == 0, out.stderr
    s = json.loads(out.stdout)
    assert s["frames"] == 0 and s["light_tracing"] == 1  # a modulated film has one frame (film.cpp:76-78)
    assert s["modulation"] == 1 and s["lambda"] == 3.5 and s["phase"] == 90
    out = run([str(p), "-D", "rif=%s" % (d / "rif.vol"), "-D", "mod=mseq", "--dry-run"])
    assert out.returncode == 1 and "modulation" in out.stderr

@pytest.mark.parametrize("tag", ["SPLINEVOLUME", "GRIDVOLUME", "HG", "HETEROGENEOUSREFRACTIVE", "ERVOLPATH"])
def test_mitsuba_binding_compiles_against_the_header_stub(tag):
    """integration/mitsuba_plugins.cpp (the file a MitsubaER maintainer builds once per plugin tag) is syntax-checked
    against integration/mitsuba_stub — declarations of exactly the Mitsuba members it uses — and against the real
    include/mitsubaer_b200.h: a change of the C ABI that the binding does not follow fails here"""
    import subprocess
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Wno-unused-function", "-DMER_PLUGIN_" + tag,
                        "-I" + os.path.join(ROOT, "integration", "mitsuba_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "mitsuba_plugins.cpp")], capture_output=True, text=True)
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr[-2000:]
