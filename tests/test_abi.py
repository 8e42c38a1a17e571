"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol that
include/mitsubaer_b200.h declares (and nothing is declared that is not bound), struct layouts match,
and every compute entry point FAILS LOUDLY when no GPU is usable (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import mitsubaer_b200 as mer
from mitsubaer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mitsubaer_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mer_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(_abi.lib, n), "libmitsubaer_b200.so does not export " + n
    assert sorted(_abi.SIGNATURES) == names
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mer_[a-z0-9_]+)", out))
    assert exported == set(names)


def test_struct_layouts_match_the_header(tmp_path):
    """compile a C probe against the header and compare sizeof/offsetof with the ctypes mirrors"""
    probe = tmp_path / "probe.c"
    fields = {"mer_volume_desc": ["res", "bbox_min", "bbox_max", "has_transform", "world_to_volume"],
              "mer_medium_desc": ["sigma_a", "sigma_s", "stepsize", "medium_sampling_weight", "strategy", "channel",
                                  "sampling_density", "shape_type", "shape", "hg_g", "density_scale", "albedo", "boundary", "radiance_scaling"],
              "mer_render_desc": ["width", "spp_total", "sample_stride", "seed", "cam_origin", "fov_deg", "filter",
                                  "max_depth", "env_radiance", "has_quad", "quad_radiance", "pool_paths", "steps_per_pass",
                                  "direct_connections", "connection", "frames", "min_bound", "bin_width", "calibrated_transient", "light_tracing", "emitter_type", "beam_power", "modulation", "phase_deg"],
              "mer_render_stats": ["samples", "ray_steps", "passes", "connections", "connection_steps", "kernel_launches", "device_ms"],
              "mer_medium_sampling_records": ["success", "t", "nsteps"],
              "mer_connection_params": ["tol2", "rrweight", "boundary_precision", "max_iterations", "start_mode"],
              "mer_connection_records": ["success", "dir_to_p2", "distance", "transmittance", "evaluations"]}
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "mitsubaer_b200.h"', "int main(void){"]
    for s, fs in fields.items():
        body.append('printf("%s %%zu\\n", sizeof(%s));' % (s, s))
        for f in fs:
            body.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    body.append("return 0;}")
    probe.write_text("\n".join(body))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    mirrors = {"mer_volume_desc": _abi.VolumeDesc, "mer_medium_desc": _abi.MediumDesc, "mer_render_desc": _abi.RenderDesc,
               "mer_render_stats": _abi.RenderStats, "mer_medium_sampling_records": _abi.SamplingRecords,
               "mer_connection_params": _abi.ConnectionParams, "mer_connection_records": _abi.ConnectionRecords}
    from oracle import oracle as orc
    oracle_mirrors = {"mer_volume_desc": orc.VolumeDesc, "mer_medium_desc": orc.MediumDesc,
                      "mer_render_desc": orc.RenderDesc, "mer_render_stats": orc.RenderStats}
    for s, fs in fields.items():
        for mirror in (mirrors[s], oracle_mirrors.get(s)):
            if mirror is None:
                continue
            assert C.sizeof(mirror) == int(got[s]), s
            for f in fs:
                assert getattr(mirror, f).offset == int(got[s + "." + f]), (s, f)


def test_version_and_error_plumbing():
    assert _abi.lib.mer_abi_version() == 10
    assert isinstance(mer.kernel_launch_count(), int)
    with pytest.raises(mer.MerError, match=r"interval \(-1, 1\)"):
        mer.HGPhaseFunction(g=-1.5)
    with pytest.raises(mer.MerError, match="unknown sampling strategy"):
        mer.HeterogeneousRefractiveMedium(strategy="nope")
    with pytest.raises(mer.MerError, match="maxDepth"):
        mer.EikonalVolPathIntegrator(maxDepth=0)


def test_vol_io_is_host_only(tmp_path):
    rng = np.random.default_rng(1)
    data = rng.random((23, 30, 20)).astype(np.float32)  # mfiles/Test.m: 20 x 30 x 23
    mer.fields.write_vol(tmp_path / "a.vol", data, (-1, -2, 0.5), (1, 1, 2))
    back, lo, hi = mer.fields.read_vol(tmp_path / "a.vol")
    assert np.array_equal(back, data) and tuple(lo) == (-1, -2, 0.5) and tuple(hi) == (1, 1, 2)
    with pytest.raises(mer.MerError, match="cannot open"):
        mer.fields.read_vol(tmp_path / "missing.vol")
    # EUInt8 payloads (gridvolume.cpp:251-262, 369-376: value / 255), one channel
    import struct
    u8 = rng.integers(0, 256, (5, 6, 7), dtype=np.uint8)
    hdr = b"VOL\x03" + struct.pack("<iiiii", 3, 7, 6, 5, 1) + struct.pack("<6f", 0, 0, 0, 1, 1, 1)
    (tmp_path / "u8.vol").write_bytes(hdr + u8.tobytes())
    back, lo, hi = mer.fields.read_vol(tmp_path / "u8.vol")
    assert back.shape == (5, 6, 7) and np.array_equal(back, u8.astype(np.float32) / np.float32(255.0))
    # untrusted headers: sizes are validated before anything is multiplied or allocated
    for res in ((-1, 6, 5), (7, 0, 5), (1 << 20, 6, 5)):
        (tmp_path / "bad.vol").write_bytes(b"VOL\x03" + struct.pack("<iiiii", 1, *res, 1) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + b"\0" * 64)
        with pytest.raises(mer.MerError, match="resolution out of range"):
            mer.fields.read_vol(tmp_path / "bad.vol")
    (tmp_path / "short.vol").write_bytes(b"VOL\x03" + struct.pack("<iiiii", 1, 7, 6, 5, 1) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + b"\0" * 100)
    with pytest.raises(mer.MerError, match="truncated"):
        mer.fields.read_vol(tmp_path / "short.vol")
    # three interleaved channels: an albedo grid (gridvolume.cpp:251-262, 401-460); float32 and uint8
    rgb = rng.random((4, 3, 5, 3)).astype(np.float32)
    mer.fields.write_vol(tmp_path / "rgb.vol", rgb, (0, 0, 0), (1, 1, 1))
    assert open(tmp_path / "rgb.vol", "rb").read()[:24] == b"VOL\x03" + struct.pack("<iiiii", 1, 5, 3, 4, 3)
    back, lo, hi = mer.fields.read_vol(tmp_path / "rgb.vol")
    assert back.shape == (4, 3, 5, 3) and np.array_equal(back, rgb)
    u8 = rng.integers(0, 256, (4, 3, 5, 3), dtype=np.uint8)
    (tmp_path / "rgb8.vol").write_bytes(b"VOL\x03" + struct.pack("<iiiii", 3, 5, 3, 4, 3) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + u8.tobytes())
    back, lo, hi = mer.fields.read_vol(tmp_path / "rgb8.vol")
    assert np.array_equal(back, u8.astype(np.float32) / np.float32(255.0))
    # float16 payloads and two-channel files are refused with the reference's own messages (gridvolume.cpp:243-258), not misread
    for enc, ch, msg in ((2, 1, "float16 volumes are not yet supported"), (1, 2, r"unsupported float32 volume data file \(2 channels, only 1 and 3"),
                         (3, 2, "unsupported uint8 volume data file"), (4, 3, "quantized-direction")):
        (tmp_path / "f16.vol").write_bytes(b"VOL\x03" + struct.pack("<iiiii", enc, 2, 2, 2, ch) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + b"\0" * 96)
        with pytest.raises(mer.MerError, match=msg):
            mer.fields.read_vol(tmp_path / "f16.vol")


@pytest.mark.skipif(mer.device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    """without a B200 every compute entry point must fail with MER_ERR_CUDA, never compute on the host"""
    data = np.ones((8, 8, 8), np.float32)
    with pytest.raises(mer.MerError) as e:
        mer.SplineDataSource(data=data, min=(0, 0, 0), max=(1, 1, 1))
    assert e.value.code == _abi.MER_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(mer.MerError):
        mer.GridDataSource(data=data, min=(0, 0, 0), max=(1, 1, 1))
    with pytest.raises(mer.MerError):
        mer.HGPhaseFunction(g=0.5).eval(np.array([[0, 0, 1.0]]), np.array([[0, 0, 1.0]]))
    with pytest.raises(mer.MerError):
        mer.develop(np.zeros((2, 2, 5), np.float32))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mitsubaer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "fields.py" and "oracle" not in txt, (dirpath, f)


def test_field_generators():
    """mfiles/createLinearRIFWithBox.m, createRadialRIFWithBox.m, createRIFFromSD.m"""
    f = mer.fields
    lin = f.linear_rif((226, 226, 51), (-225, -225, 25), (225, 225, 125))
    assert lin.shape == (51, 226, 226) and lin[0, 0, 0] == np.float32(1.3) and lin[3, 225, 7] == np.float32(1.6)
    assert np.all(lin[:, 10, :] == np.float32(1.3 + 0.3 / 225 * 10))
    rad = f.radial_rif((21, 21, 11), (-1, -1, 0), (1, 1, 1))
    assert rad[5, 10, 10] == 2.0 and abs(rad[0, 0, 0] - 1.0) < 1e-6
    sd = f.sphere_sdf((33, 33, 33), (-1, -1, -1), (1, 1, 1), radius=0.8)
    n = f.rif_from_sd(sd, 1.10, 1.50, 2.0)
    assert abs(n.max() - 1.5) < 1e-6 and n.min() == np.float32(1.1) and n[0, 0, 0] == np.float32(1.1)
    lo, hi = f.padded_bbox((-1, -1, -1), (1, 1, 1), (64, 64, 64))
    pitch = (hi - lo) / 63
    assert np.allclose(lo + 3 * pitch, -1, atol=1e-6) and np.allclose(hi - 3 * pitch, 1, atol=1e-6)


def test_plain_c_client_builds_and_fails_loudly_without_gpu(tmp_path):
    """integration/example_host.c: the boundary is usable from C99 with nothing but the header"""
    exe = tmp_path / "example_host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "integration", "example_host.c"), "-L", os.path.dirname(_abi.LIB_PATH),
                    "-lmitsubaer_b200", "-lm", "-Wl,-rpath," + os.path.dirname(_abi.LIB_PATH), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    if mer.device_count() == 0:
        assert "no CPU fallback" in out.stdout
    else:
        assert "eikonal steps" in out.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's keys,
    timed on the oracle port, no GPU needed"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "samples_per_sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
