"""Host-side mirror of the Mitsuba plugin interface for the eikonal path.

Same class names, property names, child names and error behaviour as the reference's plugins
(SURVEY.md Appendix C), implemented as thin owners of C-ABI handles:

  SplineDataSource               <volume type="splinevolume">           src/volume/splinevolume.cpp
  GridDataSource                 <volume type="gridvolume">             src/volume/gridvolume.cpp
  HGPhaseFunction                <phase type="hg">                      src/phase/hg.cpp
  HeterogeneousRefractiveMedium  <medium type="heterogeneousrefractive"> src/medium/heterogeneousrefractive.cpp
  EikonalVolPathIntegrator       <integrator type="ervolpath">  (new plugin name, SURVEY.md R3:
                                 volpath control flow + libbidir's curved-walk semantics)

All arithmetic happens in libmitsubaer_b200.so on the GPU; nothing here computes.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import check, lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _spectrum(v):
    """Properties::getSpectrum for RGB: a scalar or 3 values"""
    a = np.asarray(v, dtype=np.float32).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, 3)
    if a.size != 3:
        raise _abi.MerError(_abi.MER_ERR_INVALID, "spectrum must have 1 or 3 components (SPECTRUM_SAMPLES=3)")
    return a


def make_volume_desc(res, bbox_min, bbox_max, to_world=None):
    d = _abi.VolumeDesc()
    d.res[:] = [int(r) for r in res]
    d.bbox_min[:] = [float(x) for x in bbox_min]
    d.bbox_max[:] = [float(x) for x in bbox_max]
    if to_world is None:
        d.has_transform = 0
        d.world_to_volume[:] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    else:
        m = np.asarray(to_world, dtype=np.float64)
        if m.shape == (3, 4):
            m = np.vstack([m, [0, 0, 0, 1]])
        inv = np.linalg.inv(m.reshape(4, 4))  # m_worldToVolume = m_volumeToWorld.inverse()
        d.has_transform = 1
        d.world_to_volume[:] = [float(x) for x in inv[:3, :].reshape(12)]
    return d


class _Volume:
    handle = None

    def getAABB(self):
        d = self.desc
        return np.array(d.bbox_min[:], np.float32), np.array(d.bbox_max[:], np.float32)

    def getResolution(self):
        return tuple(self.desc.res[:])

    def getStepSize(self):
        """0.5 * min voxel pitch (splinevolume.cpp:183-185, gridvolume.cpp:196-198)"""
        lo, hi = self.getAABB()
        return float(np.min(0.5 * (hi - lo) / (np.array(self.getResolution(), np.float32) - 1)))

    def getMaximumFloatValue(self):
        return 1.0


class SplineDataSource(_Volume):
    """props: `filename` | (`data`, `res`, `min`, `max`), `toWorld`, `min`/`max` override, `device`,
    `mode` ('tricubic' parity mode | 'trilinear_packed' fast mode)."""

    def __init__(self, props=None, **kw):
        props = dict(props or {}, **kw)
        self.device = int(props.get("device", 0))
        mode = props.get("mode", "tricubic")
        self.mode = {"tricubic": _abi.RIF_TRICUBIC, "trilinear_packed": _abi.RIF_TRILINEAR_PACKED}.get(mode, mode)
        h = C.c_void_p()
        if "filename" in props:
            ov = None
            if "min" in props and "max" in props or "toWorld" in props:
                lo = props.get("min", (0, 0, 0))
                hi = props.get("max", (0, 0, 0))  # max<=min => keep the file's bbox
                ov = make_volume_desc((2, 2, 2), lo, hi, props.get("toWorld"))
            check(lib.mer_rif_create_from_file(self.device, str(props["filename"]).encode(),
                                               C.byref(ov) if ov is not None else None, self.mode, C.byref(h)))
        elif "data_ptr" in props:  # data already resident in HBM (e.g. a torch tensor's data_ptr())
            desc = make_volume_desc(props["res"], props["min"], props["max"], props.get("toWorld"))
            check(lib.mer_rif_create_device(self.device, C.byref(desc), C.c_void_p(int(props["data_ptr"])),
                                            self.mode, C.byref(h)))
        else:
            data = _f32(props["data"]).reshape(-1)
            res = props.get("res") or tuple(reversed(np.shape(props["data"])))
            desc = make_volume_desc(res, props["min"], props["max"], props.get("toWorld"))
            if data.size != desc.res[0] * desc.res[1] * desc.res[2]:
                raise _abi.MerError(_abi.MER_ERR_INVALID, "data size does not match res")
            check(lib.mer_rif_create(self.device, C.byref(desc), _fp(data), self.mode, C.byref(h)))
        self.handle = h
        self.desc = _abi.VolumeDesc()
        m = C.c_int()
        check(lib.mer_rif_desc(self.handle, C.byref(self.desc), C.byref(m)))

    def __del__(self):
        if getattr(self, "handle", None):
            lib.mer_rif_destroy(self.handle)
            self.handle = None

    def _eval(self, what, p):
        p = _f32(p, (-1, 3))
        n = p.shape[0]
        f = np.zeros(n, np.float32)
        g = np.zeros((n, 3), np.float32)
        check(lib.mer_rif_eval_batch(self.handle, what, n, _fp(p), _fp(f), _fp(g)))
        return f, g

    def value(self, p):
        return self._eval(_abi.EVAL_VALUE, p)[0]

    def gradient(self, p):
        return self._eval(_abi.EVAL_GRADIENT, p)[1]

    def valueAndGradient(self, p):
        return self._eval(_abi.EVAL_VALUE_AND_GRADIENT, p)

    def valueGradientAndHessian(self, p):
        p = _f32(p, (-1, 3))
        n = p.shape[0]
        f, g, H = np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        check(lib.mer_rif_eval_hessian_batch(self.handle, n, _fp(p), _fp(f), _fp(g), _fp(H)))
        return f, g, H

    def insideVolumeLimits(self, p):
        p = _f32(p, (-1, 3))
        out = np.zeros(p.shape[0], np.uint8)
        check(lib.mer_rif_inside_limits_batch(self.handle, p.shape[0], _fp(p), out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out.astype(bool)

    def coefficients(self):
        r = self.getResolution()
        out = np.zeros(r[0] * r[1] * r[2], np.float32)
        check(lib.mer_rif_coefficients(self.handle, _fp(out)))
        return out.reshape(r[2], r[1], r[0])

    def isAcousticRIF(self):
        return False


class GridDataSource(_Volume):
    """props: `filename` | (`data`, `res`, `min`, `max`), `toWorld`, `device`.  `data` of shape [z, y, x] makes a
    density grid, [z, y, x, 3] (or `channels` = 3) an albedo grid (gridvolume.cpp:251-262, 578-579); a file says
    which it is in its header."""

    def __init__(self, props=None, **kw):
        props = dict(props or {}, **kw)
        self.device = int(props.get("device", 0))
        h = C.c_void_p()
        if "filename" in props:
            ov = None
            if "min" in props and "max" in props or "toWorld" in props:
                ov = make_volume_desc((2, 2, 2), props.get("min", (0, 0, 0)), props.get("max", (0, 0, 0)),
                                      props.get("toWorld"))
            check(lib.mer_grid_create_from_file(self.device, str(props["filename"]).encode(),
                                                C.byref(ov) if ov is not None else None, C.byref(h)))
            self.desc = _abi.VolumeDesc()
            enc, ch = C.c_int32(), C.c_int32()
            check(lib.mer_vol_read_header(str(props["filename"]).encode(), C.byref(self.desc), C.byref(enc), C.byref(ch)))
            if ov is not None and ov.bbox_max[0] > ov.bbox_min[0]:
                self.desc.bbox_min[:] = ov.bbox_min[:]
                self.desc.bbox_max[:] = ov.bbox_max[:]
        elif "data_ptr" in props:
            self.desc = make_volume_desc(props["res"], props["min"], props["max"], props.get("toWorld"))
            check(lib.mer_grid_create_device(self.device, C.byref(self.desc), C.c_void_p(int(props["data_ptr"])),
                                             C.byref(h)))
        else:
            shape = np.shape(props["data"])
            channels = int(props.get("channels", 3 if len(shape) == 4 and shape[3] == 3 else 1))
            data = _f32(props["data"]).reshape(-1)
            res = props.get("res") or tuple(reversed(shape[:3]))
            self.desc = make_volume_desc(res, props["min"], props["max"], props.get("toWorld"))
            if channels not in (1, 3) or data.size != self.desc.res[0] * self.desc.res[1] * self.desc.res[2] * channels:
                raise _abi.MerError(_abi.MER_ERR_INVALID, "data size does not match res")
            create = lib.mer_grid_create if channels == 1 else lib.mer_grid_create_spectrum
            check(create(self.device, C.byref(self.desc), _fp(data), C.byref(h)))
        self.handle = h
        self.channels = int(lib.mer_grid_channels(h))

    def __del__(self):
        if getattr(self, "handle", None):
            lib.mer_grid_destroy(self.handle)
            self.handle = None

    def lookupFloat(self, p):
        p = _f32(p, (-1, 3))
        out = np.zeros(p.shape[0], np.float32)
        check(lib.mer_grid_lookup_batch(self.handle, p.shape[0], _fp(p), _fp(out)))
        return out

    def lookupSpectrum(self, p):
        p = _f32(p, (-1, 3))
        out = np.zeros((p.shape[0], 3), np.float32)
        check(lib.mer_grid_lookup_spectrum_batch(self.handle, p.shape[0], _fp(p), _fp(out)))
        return out

    def supportsFloatLookups(self):
        return self.channels == 1

    def supportsSpectrumLookups(self):
        return self.channels == 3

    # ---- HeterogeneousMedium (straight rays, Woodcock tracking) on this density grid: heterogeneous.cpp:546-658
    def _rays(self, ray_o, ray_d, mint, maxt):
        ro, rd = _f32(ray_o, (-1, 3)), _f32(ray_d, (-1, 3))
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, np.float32), (n,)))
        maxt = np.ascontiguousarray(np.broadcast_to(np.asarray(maxt, np.float32), (n,)))
        return ro, rd, mint, maxt, n

    def sampleDistance(self, ray_o, ray_d, mint, maxt, scale, seed):
        """-> (success, t, densityAtT); the Sampler of ray i is the Philox stream (seed, i)"""
        ro, rd, mint, maxt, n = self._rays(ray_o, ray_d, mint, maxt)
        ok, t, dens = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32)
        check(lib.mer_grid_sample_distance_batch(self.handle, float(scale), n, _fp(ro), _fp(rd), _fp(mint), _fp(maxt), int(seed),
                                                 ok.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(t), _fp(dens)))
        return ok.astype(bool), t, dens

    def evalTransmittance(self, ray_o, ray_d, mint, maxt, scale, seed):
        ro, rd, mint, maxt, n = self._rays(ray_o, ray_d, mint, maxt)
        out = np.zeros(n, np.float32)
        check(lib.mer_grid_eval_transmittance_batch(self.handle, float(scale), n, _fp(ro), _fp(rd), _fp(mint), _fp(maxt), int(seed), _fp(out)))
        return out


class HGPhaseFunction:
    """props: `g` (default 0.8, must lie in (-1, 1): hg.cpp:46-53)"""

    def __init__(self, props=None, **kw):
        props = dict(props or {}, **kw)
        self.g = float(props.get("g", 0.8))
        self.device = int(props.get("device", 0))
        if self.g >= 1 or self.g <= -1:
            raise _abi.MerError(_abi.MER_ERR_INVALID, "The asymmetry parameter must lie in the interval (-1, 1)!")

    def sample(self, wi, xi):
        """-> (wo, pdf); the returned weight is 1 like hg.cpp:97"""
        wi = _f32(wi, (-1, 3))
        xi = _f32(xi, (-1, 2))
        wo = np.zeros_like(wi)
        pdf = np.zeros(wi.shape[0], np.float32)
        check(lib.mer_hg_sample_batch(self.device, self.g, wi.shape[0], _fp(wi), _fp(xi), _fp(wo), _fp(pdf)))
        return wo, pdf

    def eval(self, wi, wo):
        wi = _f32(wi, (-1, 3))
        wo = _f32(wo, (-1, 3))
        out = np.zeros(wi.shape[0], np.float32)
        check(lib.mer_hg_eval_batch(self.device, self.g, wi.shape[0], _fp(wi), _fp(wo), _fp(out)))
        return out

    pdf = eval

    def getMeanCosine(self):
        return self.g


_STRATEGIES = {"balance": _abi.STRATEGY_BALANCE, "single": _abi.STRATEGY_SINGLE, "manual": _abi.STRATEGY_MANUAL,
               "maximum": _abi.STRATEGY_MAXIMUM}


class HeterogeneousRefractiveMedium:
    """props (heterogeneousrefractive.cpp:205-297, medium.cpp:27-37, materials.h:90-140):
    `stepsize`, `strategy`, `channel`, `samplingDensity`, `mediumSamplingWeight`,
    (`sigmaS`,`sigmaA`) | (`sigmaT`,`albedo`), `scale`; children via addChild():
    "rif" (SplineDataSource), "density" (GridDataSource, optional: new composition R2), a phase function.
    `shape`: the containment predicate, ('box', min, max) | ('sphere', centre, radius)."""

    def __init__(self, props=None, **kw):
        props = dict(props or {}, **kw)
        self.props = props
        self.rif = None
        self.sdf = None
        self.density = None
        self.phase = None
        self.handle = None
        strategy = props.get("strategy", "balance")
        if strategy not in _STRATEGIES:
            raise _abi.MerError(_abi.MER_ERR_INVALID, "Specified an unknown sampling strategy")

    def addChild(self, name, child):
        if isinstance(child, HGPhaseFunction):
            if self.phase is not None:
                raise _abi.MerError(_abi.MER_ERR_INVALID, "Medium: phase function already set")
            self.phase = child
        elif isinstance(child, SplineDataSource) and name == "rif":
            self.rif = child
        elif isinstance(child, GridDataSource) and name == "density":
            self.density = child
        elif isinstance(child, GridDataSource) and name == "albedo":  # heterogeneous.cpp:266-268
            self.albedo_grid = child
        elif isinstance(child, SplineDataSource) and name == "sdf":
            self.sdf = child
        else:
            raise _abi.MerError(_abi.MER_ERR_INVALID, 'Medium: Invalid child node! ("%s")' % type(child).__name__)
        return self

    def configure(self):
        p = self.props
        if self.rif is None:
            raise _abi.MerError(_abi.MER_ERR_INVALID, "No RIF specified!")
        d = _abi.MediumDesc()
        scale = float(p.get("scale", 1.0))
        if "sigmaT" in p and "albedo" not in p:
            raise _abi.MerError(_abi.MER_ERR_INVALID, "Medium: sigmaT needs albedo (src/medium/materials.h:112-120)")
        for name, why in (("monochromatic", "give the same value for the three channels"),
                          ("makesensordirectconnections", "use the integrator's lightTracing")):
            if bool(p.get(name, False)):  # these change what the reference computes: refuse rather than render differently
                raise _abi.MerError(_abi.MER_ERR_UNSUPPORTED, "%s=true is not carried by this path (%s)" % (name, why))
        if "sigmaT" in p and "albedo" in p:
            st, al = _spectrum(p["sigmaT"]) * scale, _spectrum(p["albedo"])
            ss, sa = st * al, st * (1 - al)
        else:
            ss = _spectrum(p.get("sigmaS", 0.0)) * scale
            sa = _spectrum(p.get("sigmaA", 0.0)) * scale
        d.sigma_a[:] = [float(x) for x in sa]
        d.sigma_s[:] = [float(x) for x in ss]
        d.stepsize = float(p.get("stepsize", 1e-3))
        d.medium_sampling_weight = float(p.get("mediumSamplingWeight", -1))
        d.strategy = _STRATEGIES[p.get("strategy", "balance")]
        d.channel = int(p.get("channel", -1))
        d.sampling_density = float(p.get("samplingDensity", 0.0))
        shape = p.get("shape")
        if shape is None:  # default: the RIF's interpolatable box shrunk to the data bbox minus 3 voxels
            lo, hi = self.rif.getAABB()
            pitch = (hi - lo) / (np.array(self.rif.getResolution(), np.float32) - 1)
            shape = ("box", lo + 3 * pitch, hi - 3 * pitch)
        if shape[0] == "box":
            d.shape_type = _abi.SHAPE_BOX
            d.shape[:] = [float(x) for x in list(shape[1]) + list(shape[2])]
        elif shape[0] == "sphere":
            d.shape_type = _abi.SHAPE_SPHERE
            d.shape[:] = [float(x) for x in list(shape[1]) + [shape[2], 0.0, 0.0]]
        elif shape[0] == "sdf":  # any closed shape: the `sdf` child volume, bounded by the given box
            d.shape_type = _abi.SHAPE_SDF
            d.shape[:] = [float(x) for x in list(shape[1]) + list(shape[2])]
            if self.sdf is None:
                raise _abi.MerError(_abi.MER_ERR_INVALID, 'shape ("sdf", min, max) needs an "sdf" child volume')
        else:
            raise _abi.MerError(_abi.MER_ERR_INVALID, "unknown shape")
        d.hg_g = self.phase.g if self.phase is not None else 0.0  # Medium::configure: isotropic default
        d.density_scale = float(p.get("densityScale", p.get("scale", 1.0))) if self.density is not None else 0.0
        d.albedo[:] = [float(x) for x in _spectrum(p.get("albedo", 0.0))]
        bsdf = str(p.get("bsdf", "null"))  # the container shape's <bsdf>: "null" | "hdielectric"
        if bsdf not in ("null", "hdielectric"):
            raise _abi.MerError(_abi.MER_ERR_INVALID, 'the container\'s bsdf must be "null" or "hdielectric"')
        d.boundary = _abi.BOUNDARY_HDIELECTRIC if bsdf == "hdielectric" else _abi.BOUNDARY_INDEX_MATCHED
        scaling = str(p.get("radianceScaling", "reference"))  # "reference": refRatioSq as the fork; "physical": its reciprocal
        if scaling not in ("reference", "physical"):
            raise _abi.MerError(_abi.MER_ERR_INVALID, 'radianceScaling must be "reference" or "physical"')
        d.radiance_scaling = 1 if scaling == "physical" else 0
        h = C.c_void_p()
        check(lib.mer_medium_create(C.byref(d), self.rif.handle, self.density.handle if self.density else None,
                                    C.byref(h)))
        self.handle = h
        if getattr(self, "albedo_grid", None) is not None:
            check(lib.mer_medium_set_albedo_grid(self.handle, self.albedo_grid.handle))
        aggressive = bool(p.get("aggressivetracing", False))
        if self.sdf is not None or aggressive:
            check(lib.mer_medium_set_sdf(self.handle, self.sdf.handle if self.sdf is not None else None, 1 if aggressive else 0))
        self.desc = _abi.MediumDesc()
        sd = C.c_float()
        check(lib.mer_medium_resolved(self.handle, C.byref(self.desc), C.byref(sd)))
        self.samplingDensity = sd.value
        self.mediumSamplingWeight = self.desc.medium_sampling_weight
        return self

    def __del__(self):
        if getattr(self, "handle", None):
            lib.mer_medium_destroy(self.handle)
            self.handle = None

    def setAlbedoVolume(self, grid):
        """attach / replace / detach (None) the `albedo` child of a configured medium (heterogeneous.cpp:262-268)"""
        check(lib.mer_medium_set_albedo_grid(self.handle, grid.handle if grid is not None else None))
        self.albedo_grid = grid
        return self

    def isheterogeneousrefractive(self):
        return True

    # ---- hot calls
    def trace(self, p, v, dist):
        p = np.array(p, np.float32).reshape(-1, 3)
        v = np.array(v, np.float32).reshape(-1, 3)
        n = p.shape[0]
        dist = np.ascontiguousarray(np.broadcast_to(np.asarray(dist, np.float32), (n,)))
        ok = np.zeros(n, np.uint8)
        ds = np.zeros(n, np.float32)
        opl = np.zeros(n, np.float32)
        ns = np.zeros(n, np.int32)
        check(lib.mer_medium_trace_batch(self.handle, n, _fp(p), _fp(v), _fp(dist),
                                         ok.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(ds), _fp(opl),
                                         ns.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(p=p, v=v, success=ok.astype(bool), dist_surf=ds, opl=opl, nsteps=ns)

    def traceTillBoundary(self, p, v):
        p = np.array(p, np.float32).reshape(-1, 3)
        v = np.array(v, np.float32).reshape(-1, 3)
        n = p.shape[0]
        ds = np.zeros(n, np.float32)
        opl = np.zeros(n, np.float32)
        ns = np.zeros(n, np.int32)
        check(lib.mer_medium_trace_till_boundary_batch(self.handle, n, _fp(p), _fp(v), _fp(ds), _fp(opl),
                                                       ns.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(p=p, v=v, dist_surf=ds, opl=opl, nsteps=ns)

    def sampleDistance(self, ray_o, ray_d, ray_mint, xi):
        """Medium::sampleDistance over a batch; xi[n][2] replays sampler->next1D()"""
        ro = _f32(ray_o, (-1, 3))
        rd = _f32(ray_d, (-1, 3))
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(ray_mint, np.float32), (n,)))
        xi = _f32(xi, (-1, 2))
        r = dict(success=np.zeros(n, np.uint8), t=np.zeros(n, np.float32), p=np.zeros((n, 3), np.float32),
                 d=np.zeros((n, 3), np.float32), optical_length=np.zeros(n, np.float32),
                 ref_ratio_sq=np.zeros(n, np.float32), transmittance=np.zeros((n, 3), np.float32),
                 pdf_success=np.zeros(n, np.float32), pdf_failure=np.zeros(n, np.float32),
                 sigma_s=np.zeros((n, 3), np.float32), nsteps=np.zeros(n, np.int32))
        rec = _abi.SamplingRecords()
        rec.success = r["success"].ctypes.data_as(C.POINTER(C.c_uint8))
        rec.nsteps = r["nsteps"].ctypes.data_as(C.POINTER(C.c_int32))
        for k in ("t", "p", "d", "optical_length", "ref_ratio_sq", "transmittance", "pdf_success", "pdf_failure",
                  "sigma_s"):
            setattr(rec, k, _fp(r[k]))
        check(lib.mer_medium_sample_distance_batch(self.handle, n, _fp(ro), _fp(rd), _fp(mint), _fp(xi), C.byref(rec)))
        r["success"] = r["success"].astype(bool)
        return r

    # ---- curved direct connections, function level
    def derivativeTrace(self, p, v, nsteps):
        """nsteps x er_derivativestep -> p, v, dp/dv0, dv/dv0"""
        p = np.array(p, np.float32).reshape(-1, 3)
        v = np.array(v, np.float32).reshape(-1, 3)
        n = p.shape[0]
        ns = np.ascontiguousarray(np.broadcast_to(np.asarray(nsteps, np.int32), (n,)))
        A, B = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        check(lib.mer_medium_derivative_trace_batch(self.handle, n, _fp(p), _fp(v), ns.ctypes.data_as(C.POINTER(C.c_int32)), _fp(A), _fp(B)))
        return dict(p=p, v=v, dpdv0=A, dvdv0=B)

    def connectionResidual(self, p1, p2, v0, is_sensor_sample=False):
        """computefdfBDPT -> error, derror (transposed Jacobian), status, nsteps"""
        p1, p2, v0 = _f32(p1, (-1, 3)), _f32(p2, (-1, 3)), _f32(v0, (-1, 3))
        n = p1.shape[0]
        err, J = np.zeros((n, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        st, ns = np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(lib.mer_medium_connection_residual_batch(self.handle, int(self.props.get("boundaryprecision", 3)), n, _fp(p1), _fp(p2), _fp(v0),
                                                       1 if is_sensor_sample else 0, _fp(err), _fp(J),
                                                       st.ctypes.data_as(C.POINTER(C.c_int32)), ns.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(error=err, derror=J, status=st, nsteps=ns)

    def eval(self, vsp, vtp, seed_dir, is_sensor_sample=False, seed=1, start_mode=0):
        """HeterogeneousRefractiveMedium::eval over a batch of (p1 = vsp, p2 = vtp) pairs
        (start_mode: 0/1 random first guess as in the reference, 2 = first guess along seed_dir)"""
        p1, p2, sd = _f32(vsp, (-1, 3)), _f32(vtp, (-1, 3)), _f32(seed_dir, (-1, 3))
        n = p1.shape[0]
        cp = _abi.ConnectionParams(float(self.props.get("tol2", 1e-6)), float(self.props.get("rrweight", 1e-2)),
                                   int(self.props.get("boundaryprecision", 3)), int(self.props.get("ceresmaxiterations", 20)), int(start_mode))
        r = dict(success=np.zeros(n, np.uint8), dir_to_p2=np.zeros((n, 3), np.float32), rev_dir_to_p1=np.zeros((n, 3), np.float32),
                 optical_length=np.zeros(n, np.float32), distance=np.zeros(n, np.float32), weight=np.zeros(n, np.float32),
                 transmittance=np.zeros((n, 3), np.float32), pdf_success=np.zeros(n, np.float32), pdf_failure=np.zeros(n, np.float32),
                 evaluations=np.zeros(n, np.int32))
        rec = _abi.ConnectionRecords()
        rec.success = r["success"].ctypes.data_as(C.POINTER(C.c_uint8))
        rec.evaluations = r["evaluations"].ctypes.data_as(C.POINTER(C.c_int32))
        for k in ("dir_to_p2", "rev_dir_to_p1", "optical_length", "distance", "weight", "transmittance", "pdf_success", "pdf_failure"):
            setattr(rec, k, _fp(r[k]))
        check(lib.mer_medium_connect_batch(self.handle, C.byref(cp), n, _fp(p1), _fp(p2), _fp(sd), 1 if is_sensor_sample else 0, int(seed),
                                           C.byref(rec)))
        r["success"] = r["success"].astype(bool)
        return r

    def evalTransmittance(self, mint, maxt):
        mint = _f32(mint).reshape(-1)
        maxt = _f32(maxt).reshape(-1)
        out = np.zeros((mint.size, 3), np.float32)
        check(lib.mer_medium_eval_transmittance_batch(self.handle, mint.size, _fp(mint), _fp(maxt), _fp(out)))
        return out


_MODULATIONS = {"none": 0, "sine": 1, "square": 2, "hamiltonian": 3}


class EikonalVolPathIntegrator:
    """<integrator type="ervolpath">: props `maxDepth` (-1), `rrDepth` (5) (MonteCarloIntegrator,
    src/librender/integrator.cpp:190-225) + scheduling knobs `poolPaths`, `stepsPerPass`.
    `directConnections` (false): next-event estimation of the quad emitter from every scattering vertex along the curved
    connection (makeDirectConnections, heterogeneousrefractive.cpp:1087-1163); the solver reads the medium's `tol2`,
    `rrweight`, `boundaryprecision`, `ceresmaxiterations` and the integrator's `connectionStart` ("straight" | "random")."""

    def __init__(self, props=None, **kw):
        props = dict(props or {}, **kw)
        self.maxDepth = int(props.get("maxDepth", -1))
        self.rrDepth = int(props.get("rrDepth", 5))
        self.poolPaths = int(props.get("poolPaths", 0))
        self.stepsPerPass = int(props.get("stepsPerPass", 0))
        dc = props.get("directConnections", False)  # False | True | "mis" (power heuristic with phase sampling, volpath.cpp:120-147)
        self.directConnections = 2 if str(dc).lower() == "mis" else (1 if (dc is True or str(dc).lower() == "true" or dc == 1) else 0)
        self.lightTracing = bool(props.get("lightTracing", False))  # emitter-side walk + t = 1 sensor connections
        self.connectionStart = str(props.get("connectionStart", "straight"))
        if self.connectionStart not in ("straight", "random"):
            raise _abi.MerError(_abi.MER_ERR_INVALID, 'connectionStart must be "straight" or "random"')
        if self.maxDepth == 0 or self.maxDepth < -1:
            raise _abi.MerError(_abi.MER_ERR_INVALID,
                                "maxDepth must be set to -1 (infinite) or a value greater than zero!")

    def render_desc(self, scene, sample_begin=0, sample_stride=1, medium=None):
        """`scene`: dict with sensor/film/emitter parameters (what the C++ shim reads off the Scene)"""
        r = _abi.RenderDesc()
        r.width, r.height = int(scene["width"]), int(scene["height"])
        r.spp_total = int(scene["sampleCount"])
        r.sample_begin, r.sample_stride = int(sample_begin), int(sample_stride)
        r.seed = int(scene.get("seed", 20201201))
        r.cam_origin[:] = [float(x) for x in scene["origin"]]
        r.cam_target[:] = [float(x) for x in scene["target"]]
        r.cam_up[:] = [float(x) for x in scene.get("up", (0, 1, 0))]
        r.fov_deg = float(scene.get("fov", 40.0))
        r.filter = {"box": _abi.FILTER_BOX, "gaussian": _abi.FILTER_GAUSSIAN}[scene.get("rfilter", "gaussian")]
        r.max_depth, r.rr_depth = self.maxDepth, self.rrDepth
        r.env_radiance[:] = [float(x) for x in _spectrum(scene.get("envRadiance", 1.0))]
        quad = scene.get("quad")
        r.has_quad = 1 if quad else 0
        if quad:
            r.quad_origin[:] = [float(x) for x in quad["origin"]]
            r.quad_u[:] = [float(x) for x in quad["u"]]
            r.quad_v[:] = [float(x) for x in quad["v"]]
            r.quad_radiance[:] = [float(x) for x in _spectrum(quad["radiance"])]
        r.pool_paths, r.steps_per_pass = self.poolPaths, self.stepsPerPass
        r.direct_connections = int(self.directConnections)
        mp = getattr(medium, "props", None) or {}
        r.connection.tol2 = float(mp.get("tol2", 1e-6))
        r.connection.rrweight = float(mp.get("rrweight", 1e-2))
        r.connection.boundary_precision = int(mp.get("boundaryprecision", 3))
        r.connection.max_iterations = int(mp.get("ceresmaxiterations", 20))
        r.connection.start_mode = 1 if self.connectionStart == "random" else 2
        r.light_tracing = 1 if self.lightTracing else 0
        em = scene.get("emitter")  # light tracing: dict(type="collimated", origin, direction, power) or the quad (default)
        if em and em.get("type", "quad") == "collimated":
            r.emitter_type = _abi.EMITTER_COLLIMATED
            r.beam_origin[:] = [float(x) for x in em.get("origin", (0, 0, 0))]
            r.beam_direction[:] = [float(x) for x in em.get("direction", (0, 0, 1))]
            r.beam_power[:] = [float(x) for x in _spectrum(em.get("power", 1.0))]
        elif em and em.get("type", "quad") != "quad":
            raise _abi.MerError(_abi.MER_ERR_INVALID, 'emitter type must be "quad" or "collimated"')
        tr = scene.get("transient")  # <film>: decomposition="transient", minBound, maxBound, binWidth, calibratedTransient
        if tr:
            width = float(tr.get("binWidth", 1.0))
            if not width > 0:
                raise _abi.MerError(_abi.MER_ERR_INVALID, "transient film: binWidth must be positive")
            r.frames = int(np.ceil((float(tr["maxBound"]) - float(tr["minBound"])) / width))  # film.cpp:73
            r.min_bound, r.bin_width = float(tr["minBound"]), width
            r.calibrated_transient = 1 if tr.get("calibrated", tr.get("calibratedTransient", False)) else 0
            mod = str(tr.get("modulation", "none")).lower()  # continuous-wave ToF: pathlengthsampler.cpp
            if mod not in _MODULATIONS:
                raise _abi.MerError(_abi.MER_ERR_INVALID, 'The "modulation" parameter must be "none", "sine", "square" or "hamiltonian" on this path')
            r.modulation = _MODULATIONS[mod]
            r.lambda_ = float(tr.get("lambda", 1.0))
            r.phase_deg = float(tr.get("phase", 0.0))
        return r

    def render(self, scene, medium, sample_begin=0, sample_stride=1):
        """-> (film[H][W][5] = [R,G,B,alpha,weight], stats dict); with scene["transient"]: film[H][W][3*frames+2]"""
        r = self.render_desc(scene, sample_begin, sample_stride, medium)
        film = np.zeros((r.height, r.width, 3 * (max(int(r.frames), 1) if not r.modulation else 1) + 2), np.float32)
        stats = _abi.RenderStats()
        check(lib.mer_render(medium.handle, C.byref(r), _fp(film), C.byref(stats)))
        return film, stats.as_dict()

    def render_multi(self, scene, media, sample_begin=0, sample_stride=1):
        """the same frame on several GPUs in one call (mer_render_multi): media = the scene's medium created on each GPU
        (grids replicated); sample indices are interleaved over the GPUs and the films reduced onto the first one"""
        r = self.render_desc(scene, sample_begin, sample_stride, media[0])
        film = np.zeros((r.height, r.width, 3 * (max(int(r.frames), 1) if not r.modulation else 1) + 2), np.float32)
        stats = _abi.RenderStats()
        handles = (C.c_void_p * len(media))(*[m.handle for m in media])
        check(lib.mer_render_multi(handles, len(media), C.byref(r), _fp(film), C.byref(stats)))
        return film, stats.as_dict()

    def render_device(self, scene, medium, film_ptr, stream=None, sample_begin=0, sample_stride=1):
        """accumulate into a device film buffer (e.g. torch tensor .data_ptr()); -> stats dict"""
        r = self.render_desc(scene, sample_begin, sample_stride, medium)
        stats = _abi.RenderStats()
        check(lib.mer_render_device(medium.handle, C.byref(r), C.c_void_p(int(film_ptr)), C.byref(stats),
                                    C.c_void_p(int(stream)) if stream else None))
        return stats.as_dict()


def develop(film, device=0):
    """HDRFilm::develop: [H][W][5] -> RGB [H][W][3]; transient film [H][W][3*frames+2] -> [H][W][frames][3]"""
    film = _f32(film)
    H, W, ch = film.shape
    frames = (ch - 2) // 3
    rgb = np.zeros((H, W, frames, 3), np.float32)
    check(lib.mer_film_develop_frames(device, W, H, frames, _fp(film), _fp(rgb)))
    return rgb[:, :, 0, :] if frames == 1 else rgb
