"""GPU parity, rows a1-a6 + a18 of SURVEY.md §8(a): prefilter, tricubic value/gradient, volume limits,
density lookup — CUDA path (through the C ABI) vs the CPU oracle on the same inputs."""
import os

import numpy as np
import pytest

import mitsubaer_b200 as mer
from common import BOX_MAX, BOX_MIN, make_field, random_points_in_box
from oracle.oracle import volume_desc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def relerr(a, b, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return np.max(np.abs(a - b)) / s


@pytest.mark.parametrize("kind,res", [("random", (20, 30, 23)), ("radial", (48, 48, 48)), ("linear", (33, 40, 17)),
                                      ("sd", (64, 64, 64))])
def test_prefilter_matches_build3d(oracle32, kind, res):
    data, lo, hi = make_field(kind, res, seed=3)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    h = oracle32.rif_create(volume_desc(res, lo, hi), data)
    c_ref = oracle32.rif_coefficients(h, data.size)
    c_gpu = rif.coefficients().reshape(-1)
    # tolerance 1e-6 relative; in practice bit-exact except for double-pow ulps
    assert relerr(c_gpu, c_ref) <= 1e-6
    assert np.mean(c_gpu == c_ref) > 0.99
    oracle32.rif_destroy(h)


@pytest.mark.parametrize("kind", ["random", "radial", "linear", "sd", "smooth"])
def test_value_and_gradient_1e5_relative(oracle32, oracle64, kind):
    res = (40, 36, 44)
    data, lo, hi = make_field(kind, res, seed=5)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    d = volume_desc(res, lo, hi)
    h32, h64 = oracle32.rif_create(d, data), oracle64.rif_create(d, data)
    p = random_points_in_box(200000, seed=11)
    f_gpu, g_gpu = rif.valueAndGradient(p)
    f32, g32 = oracle32.rif_eval(h32, p, 2)
    f64, g64 = oracle64.rif_eval(h64, p.astype(np.float64), 2)
    # parity gate (i) of SURVEY §8d.  Value: max rel err <= 1e-5.
    assert relerr(f_gpu, f32) <= 1e-5 and relerr(f_gpu, f64) <= 1e-5
    # Gradient: a difference quotient of n over one voxel, so single precision can only hold it to
    # eps * max|n| * dxres in absolute terms (the reference's own float build is that far from its
    # double build).  Gates: (a) within 1e-5 of the operand scale max|n|*dxres of the reference float
    # arithmetic, (b) at least as close to the FLOAT=double reference as the reference's float build is.
    dxres = float(np.max((np.array(res) - 1) / (hi - lo)))
    operand = float(np.max(np.abs(data))) * dxres
    e_gpu32, e_gpu64, e_cpu = np.abs(g_gpu - g32).max(), np.abs(g_gpu - g64).max(), np.abs(g32 - g64).max()
    print("gradient abs err [%s]: gpu-vs-float %.2e gpu-vs-double %.2e reference float-vs-double %.2e (max|g| %.3f)"
          % (kind, e_gpu32, e_gpu64, e_cpu, np.abs(g64).max()))
    assert e_gpu32 <= 1e-5 * operand
    assert e_gpu64 <= 1.25 * e_cpu + 1e-7
    # value() and gradient() alone agree with valueAndGradient()
    assert np.array_equal(rif.value(p[:1000]), f_gpu[:1000])
    assert np.array_equal(rif.gradient(p[:1000]), g_gpu[:1000])
    oracle32.rif_destroy(h32)
    oracle64.rif_destroy(h64)


def test_exact_knots_and_limits(oracle32):
    res = (24, 24, 24)
    data, lo, hi = make_field("random", res, seed=9)
    rif = mer.SplineDataSource(data=data, min=lo, max=hi)
    h = oracle32.rif_create(volume_desc(res, lo, hi), data)
    # points exactly on grid nodes (the reference's 5-tap case with zero end weights)
    idx = np.stack(np.meshgrid(np.arange(3, 21), np.arange(3, 21), np.arange(3, 21), indexing="ij"), -1).reshape(-1, 3)
    pitch = (hi - lo) / (np.array(res, np.float32) - 1)
    p = (lo + idx * pitch).astype(np.float32)
    f_gpu, g_gpu = rif.valueAndGradient(p)
    f_ref, g_ref = oracle32.rif_eval(h, p, 2)
    operand = float(np.max(np.abs(data))) * float(np.max((np.array(res) - 1) / (hi - lo)))
    assert relerr(f_gpu, f_ref) <= 1e-5 and np.abs(g_gpu - g_ref).max() <= 1e-5 * operand
    # insideVolumeLimits: bit-exact on points straddling the margin
    q = np.concatenate([random_points_in_box(5000, 1) * 1.6, p[:200] + 1e-6, (lo + 2 * pitch)[None], (hi - 2 * pitch)[None]])
    assert np.array_equal(rif.insideVolumeLimits(q), oracle32.rif_inside_limits(h, q.astype(np.float32)))
    oracle32.rif_destroy(h)


def test_volume_transform(oracle32):
    res = (20, 22, 24)
    data, lo, hi = make_field("smooth", res)
    th = 0.4
    to_world = np.array([[np.cos(th), -np.sin(th), 0, 0.1], [np.sin(th), np.cos(th), 0, -0.2], [0, 0, 1, 0.05], [0, 0, 0, 1]])
    rif = mer.SplineDataSource(data=data, min=lo, max=hi, toWorld=to_world)
    w2v = np.linalg.inv(to_world)[:3, :]
    h = oracle32.rif_create(volume_desc(res, lo, hi, w2v), data)
    p = random_points_in_box(20000, 2, margin=0.3)
    f_gpu, g_gpu = rif.valueAndGradient(p)
    f_ref, g_ref = oracle32.rif_eval(h, p, 2)
    operand = float(np.max(np.abs(data))) * float(np.max((np.array(res) - 1) / (hi - lo)))
    assert relerr(f_gpu, f_ref) <= 1e-5 and np.abs(g_gpu - g_ref).max() <= 1e-5 * operand
    assert np.array_equal(rif.insideVolumeLimits(p), oracle32.rif_inside_limits(h, p))
    oracle32.rif_destroy(h)


def test_golden_reference_spline():
    """vectors generated from the reference's own basisspline.h (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLDEN, "spline_ref_f32.npz"))
    rif = mer.SplineDataSource(data=g["data"], min=g["bbox_min"], max=g["bbox_max"])
    assert relerr(rif.coefficients().reshape(-1), g["coeff"]) <= 1e-6
    f, grad = rif.valueAndGradient(g["points"])
    assert relerr(f, g["value"]) <= 1e-5
    operand = float(np.max(np.abs(g["data"]))) * float(np.max((g["res"] - 1) / (g["bbox_max"] - g["bbox_min"])))
    assert np.abs(grad - g["gradient"]).max() <= 1e-5 * operand


def test_packed_trilinear_mode(oracle32):
    res = (40, 40, 40)
    data, lo, hi = make_field("linear", res)
    fast = mer.SplineDataSource(data=data, min=lo, max=hi, mode="trilinear_packed")
    cubic = mer.SplineDataSource(data=data, min=lo, max=hi, mode="tricubic")
    p = random_points_in_box(50000, 4, margin=0.2)
    ff, gf = fast.valueAndGradient(p)
    fc, gc = cubic.valueAndGradient(p)
    # a linear RIF is where both interpolants are exact in the interior (the mirror-boundary
    # prefilter is not exactly linear-preserving near the grid faces): agreement to FP error
    assert relerr(ff, fc) <= 2e-6 and np.max(np.abs(gf - gc)) <= 5e-5
    # on a curved field they differ by the trilinear interpolation error only, O(pitch^2)
    data, lo, hi = make_field("radial", res)
    fast = mer.SplineDataSource(data=data, min=lo, max=hi, mode="trilinear_packed")
    cubic = mer.SplineDataSource(data=data, min=lo, max=hi, mode="tricubic")
    ff, gf = fast.valueAndGradient(p)
    fc, gc = cubic.valueAndGradient(p)
    pitch = float((hi[0] - lo[0]) / (res[0] - 1))
    assert np.max(np.abs(ff - fc)) <= 2.0 * pitch ** 2 and np.max(np.abs(gf - gc)) <= 2.0 * pitch ** 2
    # exactly at the nodes the packed grid stores the spline itself
    idx = np.stack(np.meshgrid(np.arange(4, 36, 3), np.arange(4, 36, 3), np.arange(4, 36, 3), indexing="ij"), -1).reshape(-1, 3)
    nodes = (lo + idx * ((hi - lo) / (np.array(res, np.float32) - 1))).astype(np.float32)
    fn, gn = fast.valueAndGradient(nodes)
    cn, dn = cubic.valueAndGradient(nodes)
    assert relerr(fn, cn) <= 1e-5 and np.max(np.abs(gn - dn)) <= 1e-4


@pytest.mark.parametrize("res", [(16, 20, 12), (64, 64, 64)])
def test_density_lookup_bit_exact(oracle32, res):
    rng = np.random.default_rng(7)
    data = rng.random((res[2], res[1], res[0])).astype(np.float32)
    grid = mer.GridDataSource(data=data, min=BOX_MIN, max=BOX_MAX)
    h = oracle32.grid_create(volume_desc(res, BOX_MIN, BOX_MAX), data)
    p = np.concatenate([random_points_in_box(100000, 8) * 1.1, BOX_MIN[None], BOX_MAX[None], np.zeros((1, 3), np.float32)])
    got, ref = grid.lookupFloat(p), oracle32.grid_lookup(h, p)
    assert np.array_equal(got, ref)  # integer/index work + individually rounded lerps: bit-exact
    oracle32.grid_destroy(h)


def test_vol_roundtrip_and_file_loading(tmp_path, oracle32):
    """mfiles/Test.m:1-16 (rng(1), 20x30x23 round trip through writeGridToVol/readVolToGrid)"""
    rng = np.random.default_rng(1)
    data = (1 + rng.random((23, 30, 20))).astype(np.float32)
    lo, hi = np.array([-1, -2, 0.5], np.float32), np.array([1, 1, 2], np.float32)
    path = tmp_path / "t.vol"
    mer.fields.write_vol(path, data, lo, hi)
    back, blo, bhi = mer.fields.read_vol(path)
    assert np.array_equal(back, data) and np.array_equal(blo, lo) and np.array_equal(bhi, hi)
    raw = open(path, "rb").read()
    assert raw[:4] == b"VOL\x03" and len(raw) == 48 + data.size * 4
    a = mer.SplineDataSource(filename=str(path))
    b = mer.SplineDataSource(data=data, min=lo, max=hi)
    assert np.array_equal(a.coefficients(), b.coefficients())
    with pytest.raises(mer.MerError):
        mer.SplineDataSource(filename=str(tmp_path / "missing.vol"))
    bad = tmp_path / "bad.vol"
    bad.write_bytes(b"VOX\x03" + raw[4:])
    with pytest.raises(mer.MerError, match="incorrect header identifier"):
        mer.SplineDataSource(filename=str(bad))


def test_uint8_density_file(tmp_path):
    """EUInt8 .vol payloads (gridvolume.cpp:251-262, 369-376): streamed to the device as value / 255"""
    import struct
    rng = np.random.default_rng(8)
    u8 = rng.integers(0, 256, (24, 20, 28), dtype=np.uint8)
    hdr = b"VOL\x03" + struct.pack("<iiiii", 3, 28, 20, 24, 1) + struct.pack("<6f", -1, -1, -1, 1, 1, 1)
    (tmp_path / "u8.vol").write_bytes(hdr + u8.tobytes())
    g = mer.GridDataSource(filename=str(tmp_path / "u8.vol"))
    ref = mer.GridDataSource(data=u8.astype(np.float32) / np.float32(255.0), min=BOX_MIN, max=BOX_MAX)
    pts = rng.uniform(-0.95, 0.95, (5000, 3)).astype(np.float32)
    assert np.array_equal(g.lookupFloat(pts), ref.lookupFloat(pts))


def test_streamed_loading_of_a_multi_slab_file(tmp_path):
    """.vol files are streamed to the device in 64 MiB slabs through two pinned buffers (a 1024^3 RIF is 4 GiB
    on disk); a 288^3 file spans two slabs, a truncated one is refused, density grids go the same way"""
    res = (288, 288, 288)
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, res)
    data = mer.fields.radial_rif(res, lo, hi)
    path = tmp_path / "big.vol"
    mer.fields.write_vol(path, data, lo, hi)
    assert os.path.getsize(path) > 64 << 20
    a = mer.SplineDataSource(filename=str(path))
    b = mer.SplineDataSource(data=data, min=lo, max=hi)
    assert a.getResolution() == res and np.array_equal(a.coefficients(), b.coefficients())
    g = mer.GridDataSource(filename=str(path))
    pts = np.random.default_rng(3).uniform(-0.9, 0.9, (1000, 3)).astype(np.float32)
    assert np.array_equal(g.lookupFloat(pts), mer.GridDataSource(data=data, min=lo, max=hi).lookupFloat(pts))
    cut = tmp_path / "cut.vol"
    cut.write_bytes(open(path, "rb").read()[: 48 + (70 << 20)])
    with pytest.raises(mer.MerError, match="truncated"):
        mer.SplineDataSource(filename=str(cut))
