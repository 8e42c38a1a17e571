"""Synthetic RIF / density fields and .vol I/O (fixtures for the path).

The three RIF generators restate the reference's MATLAB scripts with the resolution as a
parameter (formulas evaluated in float64, stored float32 like mfiles/writeGridToVol.m:28-34):
  linear_rif    mfiles/createLinearRIFWithBox.m:6-25   n = nmin + (nmax-nmin) * j/(Ny-1)   (varies along y)
  radial_rif    mfiles/createRadialRIFWithBox.m:7-34   n = 2 - (r/R)^2, R = half-diagonal of the bbox
  rif_from_sd   mfiles/createRIFFromSD.m:3-38          n = nmin + k * max(-SDF, 0)^r, k = (nmax-nmin)/h^r
Arrays are indexed [z][y][x] (x fastest), the .vol order.
`xp` may be numpy or torch (torch lets bench.py build the 1024^3 grids directly in HBM).
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import check, lib


def _axes(res, bbox_min, bbox_max, xp, **kw):
    ax = [xp.linspace(float(bbox_min[i]), float(bbox_max[i]), int(res[i]), dtype=xp.float64, **kw) for i in range(3)]
    return ax[0][None, None, :], ax[1][None, :, None], ax[2][:, None, None]


def linear_rif(res, bbox_min, bbox_max, nmin=1.3, nmax=1.6, xp=np, **kw):
    x, y, z = _axes(res, bbox_min, bbox_max, xp, **kw)
    j = xp.arange(int(res[1]), dtype=xp.float64, **kw)[None, :, None]
    n = nmin + (nmax - nmin) / (int(res[1]) - 1) * j + 0 * x + 0 * z
    return n.to(xp.float32) if xp is not np else n.astype(np.float32)


def radial_rif(res, bbox_min, bbox_max, xp=np, **kw):
    x, y, z = _axes(res, bbox_min, bbox_max, xp, **kw)
    c = [(float(bbox_max[i]) + float(bbox_min[i])) / 2 for i in range(3)]
    R = float(np.linalg.norm([float(bbox_max[i]) - c[i] for i in range(3)]))
    r2 = (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2
    n = 2 - r2 / (R * R)
    return n.to(xp.float32) if xp is not np else n.astype(np.float32)


def sphere_sdf(res, bbox_min, bbox_max, centre=(0, 0, 0), radius=0.8, xp=np, **kw):
    """signed distance to a sphere (negative inside)"""
    x, y, z = _axes(res, bbox_min, bbox_max, xp, **kw)
    return ((x - centre[0]) ** 2 + (y - centre[1]) ** 2 + (z - centre[2]) ** 2) ** 0.5 - radius


def rif_from_sd(sdf, nmin=1.10, nmax=1.50, r=2.0, xp=np):
    d = -sdf
    d = xp.clip(d, 0, None) if xp is np else d.clamp(min=0)
    h = float(d.max())
    k = (nmax - nmin) / h ** r
    n = nmin + k * d ** r
    return n.to(xp.float32) if xp is not np else n.astype(np.float32)


def sine_density(res, bbox_min, bbox_max, xp=np, **kw):
    """clamp(0.5 + 0.5 sin(6 pi x) sin(6 pi y) sin(6 pi z), 0, 1) on the unit-normalised box (§8d C2)"""
    x, y, z = _axes(res, (0, 0, 0), (1, 1, 1), xp, **kw)
    d = 0.5 + 0.5 * xp.sin(6 * np.pi * x) * xp.sin(6 * np.pi * y) * xp.sin(6 * np.pi * z)
    d = xp.clip(d, 0, 1) if xp is np else d.clamp(0, 1)
    return d.to(xp.float32) if xp is not np else d.astype(np.float32)


def padded_bbox(box_min, box_max, res, pad_voxels=3):
    """bbox for a `res` grid such that the box sits `pad_voxels` voxels inside it (so that
    insideVolumeLimits, 2 strides + Epsilon, contains the shape: SURVEY.md §8d)"""
    lo = np.asarray(box_min, np.float64)
    hi = np.asarray(box_max, np.float64)
    n = np.asarray(res, np.float64)
    pitch = (hi - lo) / (n - 1 - 2 * pad_voxels)
    return (lo - pad_voxels * pitch).astype(np.float32), (hi + pad_voxels * pitch).astype(np.float32)


# ---------------------------------------------------------------- .vol v3 I/O through the C ABI
def write_vol(path, data, bbox_min, bbox_max):
    """data[z][y][x] (or data[z][y][x][3], an albedo grid) float32 -> Mitsuba .vol v3 (mfiles/writeGridToVol.m)"""
    data = np.ascontiguousarray(data, dtype=np.float32)
    d = _abi.VolumeDesc()
    d.res[:] = [data.shape[2], data.shape[1], data.shape[0]]
    d.bbox_min[:] = [float(x) for x in bbox_min]
    d.bbox_max[:] = [float(x) for x in bbox_max]
    write = lib.mer_vol_write_spectrum if data.ndim == 4 and data.shape[3] == 3 else lib.mer_vol_write
    check(write(str(path).encode(), C.byref(d), data.ctypes.data_as(C.POINTER(C.c_float))))


def read_vol(path):
    """-> (data[z][y][x] or data[z][y][x][3], bbox_min, bbox_max)  (mfiles/readVolToGrid.m, splinevolume.cpp:204-273)"""
    d = _abi.VolumeDesc()
    enc, ch = C.c_int32(), C.c_int32()
    check(lib.mer_vol_read_header(str(path).encode(), C.byref(d), C.byref(enc), C.byref(ch)))
    data = np.zeros((d.res[2], d.res[1], d.res[0]) + ((ch.value,) if ch.value > 1 else ()), np.float32)
    check(lib.mer_vol_read_data(str(path).encode(), data.ctypes.data_as(C.POINTER(C.c_float)), data.size))
    return data, np.array(d.bbox_min[:], np.float32), np.array(d.bbox_max[:], np.float32)
