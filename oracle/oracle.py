"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (mitsubaer_b200/) never imports it.

  * `Oracle(dtype)`  -> oracle/libmer_oracle.so   (restatement of SURVEY.md §8a rows a1-a24,
                        float or double FLOAT)
  * `RefSpline(dtype)` -> oracle/_ref/libmer_refspline_{f,d}.so  (the reference's own
                        include/mitsuba/core/basisspline.h compiled verbatim; rows a1-a4)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class VolumeDesc(C.Structure):
    """mirror of mer_volume_desc (include/mitsubaer_b200.h)"""
    _fields_ = [("res", C.c_int32 * 3), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("has_transform", C.c_int32), ("world_to_volume", C.c_float * 12)]


class MediumDesc(C.Structure):
    """mirror of mer_medium_desc"""
    _fields_ = [("sigma_a", C.c_float * 3), ("sigma_s", C.c_float * 3), ("stepsize", C.c_float),
                ("medium_sampling_weight", C.c_float), ("strategy", C.c_int32), ("channel", C.c_int32),
                ("sampling_density", C.c_float), ("shape_type", C.c_int32), ("shape", C.c_float * 6),
                ("hg_g", C.c_float), ("density_scale", C.c_float), ("albedo", C.c_float * 3), ("boundary", C.c_int32),
                ("radiance_scaling", C.c_int32)]


class ConnectionParams(C.Structure):
    """mirror of mer_connection_params"""
    _fields_ = [("tol2", C.c_float), ("rrweight", C.c_float), ("boundary_precision", C.c_int32), ("max_iterations", C.c_int32),
                ("start_mode", C.c_int32)]


class RenderDesc(C.Structure):
    """mirror of mer_render_desc"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp_total", C.c_int32),
                ("sample_begin", C.c_int32), ("sample_stride", C.c_int32), ("seed", C.c_uint64),
                ("cam_origin", C.c_float * 3), ("cam_target", C.c_float * 3), ("cam_up", C.c_float * 3),
                ("fov_deg", C.c_float), ("filter", C.c_int32), ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
                ("env_radiance", C.c_float * 3), ("has_quad", C.c_int32), ("quad_origin", C.c_float * 3),
                ("quad_u", C.c_float * 3), ("quad_v", C.c_float * 3), ("quad_radiance", C.c_float * 3),
                ("pool_paths", C.c_int32), ("steps_per_pass", C.c_int32), ("direct_connections", C.c_int32),
                ("connection", ConnectionParams), ("frames", C.c_int32), ("min_bound", C.c_float), ("bin_width", C.c_float),
                ("calibrated_transient", C.c_int32), ("light_tracing", C.c_int32), ("emitter_type", C.c_int32),
                ("beam_origin", C.c_float * 3), ("beam_direction", C.c_float * 3), ("beam_power", C.c_float * 3),
                ("modulation", C.c_int32), ("lambda_", C.c_float), ("phase_deg", C.c_float)]


class RenderStats(C.Structure):
    """mirror of mer_render_stats"""
    _fields_ = [("samples", C.c_uint64), ("ray_steps", C.c_uint64), ("scatter_events", C.c_uint64),
                ("null_collisions", C.c_uint64), ("boundary_exits", C.c_uint64),
                ("nonfinite_dropped", C.c_uint64), ("passes", C.c_uint64), ("connections", C.c_uint64),
                ("connections_failed", C.c_uint64), ("connection_steps", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_ms", C.c_float), ("step_kernel_ms", C.c_float), ("step_launches", C.c_uint64), ("tail_ms", C.c_float),
                ("block_fetches", C.c_uint64), ("step_lanes_per_sm", C.c_uint32), ("reserved0", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def volume_desc(res, bbox_min, bbox_max, world_to_volume=None):
    d = VolumeDesc()
    d.res[:] = [int(r) for r in res]
    d.bbox_min[:] = [float(x) for x in bbox_min]
    d.bbox_max[:] = [float(x) for x in bbox_max]
    if world_to_volume is None:
        d.has_transform = 0
        d.world_to_volume[:] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]
    else:
        d.has_transform = 1
        d.world_to_volume[:] = [float(x) for x in np.asarray(world_to_volume, dtype=np.float64).reshape(12)]
    return d


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def build(target="oracle"):
    """(re)build with oracle/Makefile; `ref` is a no-op when /root/reference is absent"""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class RefSpline:
    """The reference's Spline<3>, compiled verbatim (oracle/ref_spline.cpp)."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        suf = "f" if self.dtype == np.float32 else "d"
        path = os.path.join(REF_DIR, "libmer_refspline_%s.so" % suf)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.ct = C.c_float if suf == "f" else C.c_double
        self.lib.ref_spline_build.restype = C.c_void_p
        self.lib.ref_spline_stride.restype = self.ct
        assert self.lib.ref_spline_sizeof_float() == self.dtype.itemsize
        self.h = None

    @staticmethod
    def available():
        return os.path.exists(os.path.join(REF_DIR, "libmer_refspline_f.so"))

    def build(self, data, res, bbox_min, bbox_max):
        data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        self.res = tuple(int(r) for r in res)
        assert data.size == self.res[0] * self.res[1] * self.res[2]
        N = (C.c_int * 3)(*self.res)
        bmin = (C.c_float * 3)(*[float(x) for x in bbox_min])
        bmax = (C.c_float * 3)(*[float(x) for x in bbox_max])
        self.h = C.c_void_p(self.lib.ref_spline_build(_ptr(data, C.c_float), N, bmin, bmax))
        return self

    def coefficients(self):
        out = np.empty(self.res[0] * self.res[1] * self.res[2], dtype=self.dtype)
        self.lib.ref_spline_coeffs(self.h, _ptr(out, self.ct))
        return out

    def eval(self, p, what=2):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        f = np.zeros(n, dtype=self.dtype)
        g = np.zeros((n, 3), dtype=self.dtype)
        self.lib.ref_spline_eval(self.h, C.c_int(what), C.c_size_t(n), _ptr(p, self.ct), _ptr(f, self.ct),
                                 _ptr(g, self.ct))
        return f, g

    def eval_hessian(self, p):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        f = np.zeros(n, dtype=self.dtype)
        g = np.zeros((n, 3), dtype=self.dtype)
        H = np.zeros((n, 3, 3), dtype=self.dtype)
        self.lib.ref_spline_eval_hessian(self.h, C.c_size_t(n), _ptr(p, self.ct), _ptr(f, self.ct),
                                         _ptr(g, self.ct), _ptr(H, self.ct))
        return f, g, H

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_spline_free(self.h)
            self.h = None


def _maxexp(fn, sigma_t, what, x):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    st = (C.c_float * 3)(*[float(v) for v in sigma_t])
    a, b = np.zeros_like(x), np.zeros_like(x)
    if fn(st, C.c_int(what), C.c_size_t(x.size), _ptr(x, C.c_float), _ptr(a, C.c_float), _ptr(b, C.c_float)) != 0:
        raise ValueError("sigmaT must vary across channels")
    return (a, b) if what == 0 else a


class RefPhase:
    """The reference's HGPhaseFunction (src/phase/hg.cpp), Frame, coordinateSystem and fresnelDielectricExt
    (src/libcore/util.cpp), compiled verbatim (oracle/ref_phase.cpp -> oracle/_ref/libmer_refphase.so)."""

    def __init__(self):
        path = os.path.join(REF_DIR, "libmer_refphase.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    @staticmethod
    def available():
        return os.path.exists(os.path.join(REF_DIR, "libmer_refphase.so"))

    def hg_sample(self, g, wi, xi):
        wi = np.ascontiguousarray(wi, dtype=np.float32).reshape(-1, 3)
        xi = np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
        wo = np.zeros_like(wi)
        pdf = np.zeros(wi.shape[0], dtype=np.float32)
        self.lib.ref_hg_sample(C.c_float(g), C.c_size_t(wi.shape[0]), _ptr(wi, C.c_float), _ptr(xi, C.c_float),
                               _ptr(wo, C.c_float), _ptr(pdf, C.c_float))
        return wo, pdf

    def hg_eval(self, g, wi, wo):
        wi = np.ascontiguousarray(wi, dtype=np.float32).reshape(-1, 3)
        wo = np.ascontiguousarray(wo, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(wi.shape[0], dtype=np.float32)
        self.lib.ref_hg_eval(C.c_float(g), C.c_size_t(wi.shape[0]), _ptr(wi, C.c_float), _ptr(wo, C.c_float), _ptr(out, C.c_float))
        return out

    def coordinate_system(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
        b, c = np.zeros_like(a), np.zeros_like(a)
        self.lib.ref_coordinate_system(C.c_size_t(a.shape[0]), _ptr(a, C.c_float), _ptr(b, C.c_float), _ptr(c, C.c_float))
        return b, c

    def hdielectric_sample(self, d, N, eta, u, importance=False):
        """HSmoothDielectric::sample (hdielectric.cpp:244-300) for rays along d at a surface with normal N and interior index eta
        -> (direction out, weight, relative index of the event, transmitted?)"""
        d = np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 3)
        N = np.ascontiguousarray(N, dtype=np.float32).reshape(-1, 3)
        n = d.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, np.float32), (n,)))
        u = np.ascontiguousarray(u, dtype=np.float32).reshape(-1)
        out, w, es, tr = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.ref_hdielectric_sample(C.c_size_t(n), _ptr(d, C.c_float), _ptr(N, C.c_float), _ptr(eta, C.c_float), _ptr(u, C.c_float), C.c_int(1 if importance else 0),
                    _ptr(out, C.c_float), _ptr(w, C.c_float), _ptr(es, C.c_float), _ptr(tr, C.c_int))
        return out, w, es, tr.astype(bool)

    def maxexp(self, sigma_t, what, x):
        """MaxExpDist (src/medium/maxexp.h): what = 0 sample(u) -> (t, pdf); 1 pdf(t); 2 cdf(t)"""
        return _maxexp(self.lib.ref_maxexp, sigma_t, what, x)

    def fresnel_dielectric_ext(self, cos_i, eta):
        cos_i = np.ascontiguousarray(cos_i, dtype=np.float32).reshape(-1)
        eta = np.ascontiguousarray(eta, dtype=np.float32).reshape(-1)
        F, ct = np.zeros_like(cos_i), np.zeros_like(cos_i)
        self.lib.ref_fresnel_dielectric_ext(C.c_size_t(cos_i.size), _ptr(cos_i, C.c_float), _ptr(eta, C.c_float), _ptr(F, C.c_float),
                                            _ptr(ct, C.c_float))
        return F, ct


class RefTrace:
    """The reference's stepper — er_step, trace, aggressive_trace, traceTillBoundary, insideShape (= hackForSphere: a
    hard-coded sphere) of src/medium/heterogeneousrefractive.cpp and SplineDataSource's lookup wrappers of
    src/volume/splinevolume.cpp — compiled verbatim (oracle/ref_trace.cpp -> oracle/_ref/libmer_reftrace.so), FLOAT = float."""

    def __init__(self, data, bmin=None, bmax=None, stepsize=1e-3):
        """data: a float grid [z][y][x] with its bounding box, or the path of a .vol file, which is then read by the reference's
        own SplineDataSource::loadFromFile (splinevolume.cpp:204-317)"""
        path = os.path.join(REF_DIR, "libmer_reftrace.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_medium_create.restype = C.c_void_p
        self.lib.ref_medium_create_from_file.restype = C.c_void_p
        if isinstance(data, (str, os.PathLike)):
            self.h = C.c_void_p(self.lib.ref_medium_create_from_file(os.fspath(data).encode(), C.c_float(stepsize)))
            return
        data = np.ascontiguousarray(data, dtype=np.float32)
        N = (C.c_int * 3)(data.shape[2], data.shape[1], data.shape[0])  # arrays are [z][y][x]
        lo = (C.c_float * 3)(*[float(v) for v in bmin])
        hi = (C.c_float * 3)(*[float(v) for v in bmax])
        self.h = C.c_void_p(self.lib.ref_medium_create(_ptr(data, C.c_float), N, lo, hi, C.c_float(stepsize)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_medium_free(self.h)
            self.h = None

    @staticmethod
    def available():
        return os.path.exists(os.path.join(REF_DIR, "libmer_reftrace.so"))

    def container(self):
        """(centre, radius) of the sphere the compiled insideShape() tests against"""
        cr = (C.c_float * 4)()
        self.lib.ref_medium_container(cr)
        return np.array(cr[:3], np.float32), float(np.float32(cr[3]))

    @staticmethod
    def _pv(p, v):
        return (np.array(p, dtype=np.float32, order="C", copy=True).reshape(-1, 3),
                np.array(v, dtype=np.float32, order="C", copy=True).reshape(-1, 3))

    def value_gradient(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        f, g = np.zeros(p.shape[0], np.float32), np.zeros_like(p)
        self.lib.ref_rif_value_gradient(self.h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(f, C.c_float), _ptr(g, C.c_float))
        return f, g

    def inside_limits(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(p.shape[0], np.int32)
        self.lib.ref_rif_inside_limits(self.h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_int))
        return out.astype(bool)

    def inside_shape(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(p.shape[0], np.int32)
        self.lib.ref_inside_shape(self.h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_int))
        return out.astype(bool)

    def er_step(self, p, v, stepsize, opl=None):
        p, v = self._pv(p, v)
        n = p.shape[0]
        h = np.ascontiguousarray(np.broadcast_to(np.asarray(stepsize, np.float32), (n,)), dtype=np.float32)
        o = np.zeros(n, np.float32) if opl is None else np.array(opl, dtype=np.float32, copy=True)
        self.lib.ref_er_step(self.h, C.c_size_t(n), _ptr(p, C.c_float), _ptr(v, C.c_float), _ptr(h, C.c_float), _ptr(o, C.c_float))
        return p, v, o

    def trace(self, p, v, dist):
        """-> p, v, distSurf, opticalDistance, success  (trace(), heterogeneousrefractive.cpp:671-691)"""
        p, v = self._pv(p, v)
        n = p.shape[0]
        d = np.ascontiguousarray(dist, dtype=np.float32).reshape(-1)
        ds, o, ok = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.ref_trace(self.h, C.c_size_t(n), _ptr(p, C.c_float), _ptr(v, C.c_float), _ptr(d, C.c_float), _ptr(ds, C.c_float),
                           _ptr(o, C.c_float), _ptr(ok, C.c_int))
        return p, v, ds, o, ok.astype(bool)

    def trace_till_boundary(self, p, v):
        """-> p, v, distSurf, opticalDistance  (traceTillBoundary(), :742-776)"""
        p, v = self._pv(p, v)
        n = p.shape[0]
        ds, o = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.lib.ref_trace_till_boundary(self.h, C.c_size_t(n), _ptr(p, C.c_float), _ptr(v, C.c_float), _ptr(ds, C.c_float), _ptr(o, C.c_float))
        return p, v, ds, o

    def configure(self, sigma_a, sigma_s, strategy, sampling_density, medium_sampling_weight, sdf=None, sdf_min=None, sdf_max=None,
                  aggressive=False):
        """the members the constructor resolves from the properties (heterogeneousrefractive.cpp:201-300), handed over resolved;
        strategy: "balance" | "single" | "manual" | "maximum"; sdf: optional signed-distance grid [z][y][x] (the `sdf` child)"""
        sa = (C.c_float * 3)(*[float(x) for x in np.broadcast_to(np.asarray(sigma_a, np.float32), (3,))])
        ss = (C.c_float * 3)(*[float(x) for x in np.broadcast_to(np.asarray(sigma_s, np.float32), (3,))])
        st = {"balance": 0, "single": 1, "manual": 2, "maximum": 3}[strategy]
        if sdf is not None:
            sdf = np.ascontiguousarray(sdf, dtype=np.float32)
            N = (C.c_int * 3)(sdf.shape[2], sdf.shape[1], sdf.shape[0])
            lo = (C.c_float * 3)(*[float(v) for v in sdf_min])
            hi = (C.c_float * 3)(*[float(v) for v in sdf_max])
            self.lib.ref_medium_configure(self.h, sa, ss, C.c_int(st), C.c_float(sampling_density), C.c_float(medium_sampling_weight),
                                          _ptr(sdf, C.c_float), N, lo, hi, C.c_int(1 if aggressive else 0))
        else:
            self.lib.ref_medium_configure(self.h, sa, ss, C.c_int(st), C.c_float(sampling_density), C.c_float(medium_sampling_weight),
                                          None, None, None, None, C.c_int(0))
        return self

    def resolve(self, sigma_a, sigma_s, strategy, medium_sampling_weight=-1.0, sampling_density=0.0, channel=-1):
        """the constructor's resolution of the free-flight sampling (heterogeneousrefractive.cpp:238-293), compiled verbatim
        -> (mediumSamplingWeight, samplingDensity, strategy index)"""
        sa = (C.c_float * 3)(*[float(x) for x in np.broadcast_to(np.asarray(sigma_a, np.float32), (3,))])
        ss = (C.c_float * 3)(*[float(x) for x in np.broadcast_to(np.asarray(sigma_s, np.float32), (3,))])
        w, d, st = C.c_float(), C.c_float(), C.c_int()
        self.lib.ref_medium_resolve(self.h, sa, ss, strategy.encode(), C.c_float(medium_sampling_weight), C.c_float(sampling_density), C.c_int(channel),
                                    C.byref(w), C.byref(d), C.byref(st))
        return w.value, d.value, st.value

    def sample_distance(self, ro, rd, mint, xi):
        """Medium::sampleDistance (heterogeneousrefractive.cpp:402-568) over a batch; xi[n][2] replays sampler->next1D()"""
        ro = np.ascontiguousarray(ro, dtype=np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rd, dtype=np.float32).reshape(-1, 3)
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, dtype=np.float32), (n,)))
        xi = np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
        r = dict(success=np.zeros(n, np.int32), t=np.zeros(n, np.float32), p=np.zeros((n, 3), np.float32), d=np.zeros((n, 3), np.float32),
                 optical_length=np.zeros(n, np.float32), ref_ratio_sq=np.zeros(n, np.float32), transmittance=np.zeros((n, 3), np.float32),
                 pdf_success=np.zeros(n, np.float32), pdf_failure=np.zeros(n, np.float32))
        self.lib.ref_sample_distance(self.h, C.c_size_t(n), _ptr(ro, C.c_float), _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(xi, C.c_float),
                                     _ptr(r["success"], C.c_int), _ptr(r["t"], C.c_float), _ptr(r["p"], C.c_float), _ptr(r["d"], C.c_float),
                                     _ptr(r["optical_length"], C.c_float), _ptr(r["ref_ratio_sq"], C.c_float), _ptr(r["transmittance"], C.c_float),
                                     _ptr(r["pdf_success"], C.c_float), _ptr(r["pdf_failure"], C.c_float))
        r["success"] = r["success"].astype(bool)
        return r

    def eval_transmittance(self, mint, maxt):
        mint = np.ascontiguousarray(mint, dtype=np.float32).reshape(-1)
        maxt = np.ascontiguousarray(maxt, dtype=np.float32).reshape(-1)
        out = np.zeros((mint.size, 3), np.float32)
        self.lib.ref_eval_transmittance(self.h, C.c_size_t(mint.size), _ptr(mint, C.c_float), _ptr(maxt, C.c_float), _ptr(out, C.c_float))
        return out

    def set_connection(self, precision, tol):
        """`boundaryprecision` and `tol` of the medium (the shooting problem's bisection depth and acceptance)"""
        self.lib.ref_medium_set_connection(self.h, C.c_int(int(precision)), C.c_float(tol))
        return self

    def derivative_trace(self, p, v, nsteps):
        """er_derivativestep (:798-814) nsteps times from dpdv0 = 0, dvdv0 = I"""
        p, v = self._pv(p, v)
        n = p.shape[0]
        A = np.zeros((n, 3, 3), np.float32)
        B = np.ascontiguousarray(np.broadcast_to(np.eye(3, dtype=np.float32), (n, 3, 3)))
        self.lib.ref_derivative_trace(self.h, C.c_size_t(n), _ptr(p, C.c_float), _ptr(v, C.c_float), _ptr(A, C.c_float), _ptr(B, C.c_float),
                                      C.c_int(int(nsteps)))
        return dict(p=p, v=v, dpdv0=A, dvdv0=B)

    def connection_residual(self, p1, p2, v0, is_sensor=False):
        """computefdfBDPT (:816-939) as DirectConnectionCostFunction::Evaluate calls it -> error [n,3], derror [n,3,3] (derror.m)"""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 3)
        v0 = np.ascontiguousarray(v0, dtype=np.float32).reshape(-1, 3)
        n = p1.shape[0]
        err, J = np.zeros((n, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        self.lib.ref_connection_residual(self.h, C.c_size_t(n), _ptr(p1, C.c_float), _ptr(p2, C.c_float), _ptr(v0, C.c_float),
                                         C.c_int(1 if is_sensor else 0), _ptr(err, C.c_float), _ptr(J, C.c_float))
        return dict(error=err, derror=J)

    def path_lengths(self, p1, p2, dir_to_p2, is_sensor=False):
        """computePathLengthsTillClosestP2 (:941-1030) -> success, revDirToP1, opticalDistToP2, distToP2"""
        p1 = np.ascontiguousarray(p1, dtype=np.float32).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir_to_p2, dtype=np.float32).reshape(-1, 3)
        n = p1.shape[0]
        rev, od, dist, ok = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.ref_path_lengths(self.h, C.c_size_t(n), _ptr(p1, C.c_float), _ptr(p2, C.c_float), _ptr(d, C.c_float), C.c_int(1 if is_sensor else 0),
                                  _ptr(rev, C.c_float), _ptr(od, C.c_float), _ptr(dist, C.c_float), _ptr(ok, C.c_int))
        return dict(success=ok.astype(bool), rev_dir=rev, optical_dist=od, dist=dist)

    def aggressive_trace(self, p, v, dist):
        p, v = self._pv(p, v)
        n = p.shape[0]
        d = np.ascontiguousarray(dist, dtype=np.float32).reshape(-1)
        o = np.zeros(n, np.float32)
        self.lib.ref_aggressive_trace(self.h, C.c_size_t(n), _ptr(p, C.c_float), _ptr(v, C.c_float), _ptr(d, C.c_float), _ptr(o, C.c_float))
        return p, v, o


class RefGrid:
    """The reference's GridDataSource::lookupFloat (src/volume/gridvolume.cpp:337-388) with the Transform functions its
    configure() uses, compiled verbatim (oracle/ref_volume.cpp -> oracle/_ref/libmer_reftrace.so); float32 or uint8 payloads."""

    def __init__(self, data, bmin, bmax):
        path = os.path.join(REF_DIR, "libmer_reftrace.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_grid_create_channels.restype = C.c_void_p
        data = np.ascontiguousarray(data)
        assert data.dtype in (np.float32, np.uint8)
        self.channels = 3 if data.ndim == 4 else 1  # [z, y, x] density or [z, y, x, 3] albedo (lookupSpectrum, gridvolume.cpp:386-463)
        assert data.ndim == 3 or data.shape[3] == 3
        N = (C.c_int * 3)(data.shape[2], data.shape[1], data.shape[0])
        lo = (C.c_float * 3)(*[float(v) for v in bmin])
        hi = (C.c_float * 3)(*[float(v) for v in bmax])
        self.h = C.c_void_p(self.lib.ref_grid_create_channels(data.ctypes.data_as(C.c_void_p), N, lo, hi, C.c_int(1 if data.dtype == np.float32 else 3),
                                                              C.c_int(self.channels)))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_grid_free(self.h)
            self.h = None

    def lookup(self, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(p.shape[0], np.float32)
        self.lib.ref_grid_lookup(self.h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_float))
        return out

    def lookup_spectrum(self, p):
        assert self.channels == 3
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((p.shape[0], 3), np.float32)
        self.lib.ref_grid_lookup_spectrum(self.h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_float))
        return out


class RefHeterogeneousMedium:
    """The reference's straight-ray HeterogeneousMedium::sampleDistance / evalTransmittance (src/medium/heterogeneous.cpp:546-672)
    over a RefGrid, compiled verbatim; method "woodcock" | "simpson"; xi[n][k]: what sampler->next1D() returns, in order."""

    def __init__(self, grid, bmin, bmax, scale, max_density, method="woodcock", step_size=0.0, albedo=(0.9, 0.9, 0.9)):
        self.grid, self.lib = grid, grid.lib
        self.lib.ref_hetmedium_create.restype = C.c_void_p
        lo = (C.c_float * 3)(*[float(v) for v in bmin])
        hi = (C.c_float * 3)(*[float(v) for v in bmax])
        al = (C.c_float * 3)(*[float(v) for v in albedo])
        self.h = C.c_void_p(self.lib.ref_hetmedium_create(grid.h, lo, hi, C.c_float(scale), C.c_float(max_density),
                                                          C.c_int(1 if method == "woodcock" else 0), C.c_float(step_size), al))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_hetmedium_free(self.h)
            self.h = None

    def _rays(self, ro, rd, mint, maxt, xi):
        ro = np.ascontiguousarray(ro, dtype=np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rd, dtype=np.float32).reshape(-1, 3)
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, np.float32), (n,)))
        maxt = np.ascontiguousarray(np.broadcast_to(np.asarray(maxt, np.float32), (n,)))
        xi = np.ascontiguousarray(xi, dtype=np.float32).reshape(n, -1)
        return ro, rd, mint, maxt, xi, n

    def sample_distance(self, ro, rd, mint, maxt, xi):
        ro, rd, mint, maxt, xi, n = self._rays(ro, rd, mint, maxt, xi)
        ok, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
        ss, T = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self.lib.ref_hetmedium_sample_distance(self.h, C.c_size_t(n), _ptr(ro, C.c_float), _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(maxt, C.c_float),
                                               _ptr(xi, C.c_float), C.c_size_t(xi.shape[1]), _ptr(ok, C.c_int), _ptr(t, C.c_float), _ptr(ss, C.c_float), _ptr(T, C.c_float))
        return ok.astype(bool), t, ss, T

    def eval_transmittance(self, ro, rd, mint, maxt, xi):
        ro, rd, mint, maxt, xi, n = self._rays(ro, rd, mint, maxt, xi)
        out = np.zeros(n, np.float32)
        self.lib.ref_hetmedium_eval_transmittance(self.h, C.c_size_t(n), _ptr(ro, C.c_float), _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(maxt, C.c_float),
                                                  _ptr(xi, C.c_float), C.c_size_t(xi.shape[1]), _ptr(out, C.c_float))
        return out


class RefFilm:
    """The reference's reconstruction-filter table (rfilter.cpp / rfilter.h, gaussian.cpp, box.cpp) and ImageBlock::put
    (imageblock.h), compiled verbatim (oracle/ref_film.cpp -> oracle/_ref/libmer_reftrace.so)."""

    def __init__(self):
        path = os.path.join(REF_DIR, "libmer_reftrace.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def filter_table(self, ftype):
        vals = np.zeros(32, np.float32)
        r, s, b = C.c_float(), C.c_float(), C.c_int()
        self.lib.ref_filter_table(C.c_int(ftype), _ptr(vals, C.c_float), C.byref(r), C.byref(s), C.byref(b))
        return vals, r.value, s.value, b.value

    def mi_weight(self, pdf_a, pdf_b):
        """VolumetricPathTracer::miWeight (volpath.cpp:430-433) compiled verbatim"""
        a = np.ascontiguousarray(pdf_a, dtype=np.float32).reshape(-1)
        b = np.ascontiguousarray(pdf_b, dtype=np.float32).reshape(-1)
        out = np.zeros(a.shape[0], np.float32)
        self.lib.ref_mi_weight(C.c_size_t(a.shape[0]), _ptr(a, C.c_float), _ptr(b, C.c_float), _ptr(out, C.c_float))
        return out

    def camera_rays(self, origin, target, up, fov, W, H, sample_pos):
        """PerspectiveCameraImpl::sampleRay (perspective.cpp:247-269) behind Transform::lookAt / perspective and the reference's
        own 4x4 inversion -> ray origins and directions for pixel samples [n][2] (fractional pixel coordinates)"""
        sp = np.ascontiguousarray(sample_pos, dtype=np.float32).reshape(-1, 2)
        o, d = np.zeros((sp.shape[0], 3), np.float32), np.zeros((sp.shape[0], 3), np.float32)
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        self.lib.ref_camera_rays(f3(origin), f3(target), f3(up), C.c_float(fov), C.c_int(W), C.c_int(H), C.c_size_t(sp.shape[0]), _ptr(sp, C.c_float),
                                 _ptr(o, C.c_float), _ptr(d, C.c_float))
        return o, d

    def film_put(self, ftype, W, H, pos, values):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 2)
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(pos.shape[0], -1)
        film = np.zeros((H, W, values.shape[1]), np.float32)
        ok = np.zeros(pos.shape[0], np.int32)
        self.lib.ref_film_put(C.c_int(ftype), C.c_int(W), C.c_int(H), C.c_int(values.shape[1]), C.c_size_t(pos.shape[0]), _ptr(pos, C.c_float),
                              _ptr(values, C.c_float), _ptr(film, C.c_float), _ptr(ok, C.c_int))
        return film, ok.astype(bool)


class Oracle:
    """The restated path (oracle/mer_oracle.cpp) in float (`Float`) or double (-DFLOATDEBUG)."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.suf = "_f" if self.dtype == np.float32 else "_d"
        path = os.path.join(HERE, "libmer_oracle.so")
        if not os.path.exists(path):
            build("oracle")
        self.lib = C.CDLL(path)
        self.ct = C.c_float if self.dtype == np.float32 else C.c_double
        for name in ("orc_rif_create", "orc_medium_create"):
            getattr(self.lib, name + self.suf).restype = C.c_void_p
        self.lib.orc_grid_create.restype = C.c_void_p
        self.lib.orc_grid_create_spectrum.restype = C.c_void_p

    def _fn(self, name):
        return getattr(self.lib, name + self.suf)

    def num_threads(self):
        return int(self.lib.orc_num_threads())

    # ---- volumes
    def rif_create(self, desc, data):
        data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        assert data.size == desc.res[0] * desc.res[1] * desc.res[2]
        return C.c_void_p(self._fn("orc_rif_create")(C.byref(desc), _ptr(data, C.c_float)))

    def rif_destroy(self, h):
        self._fn("orc_rif_destroy")(h)

    def rif_coefficients(self, h, n):
        out = np.empty(n, dtype=self.dtype)
        self._fn("orc_rif_coefficients")(h, _ptr(out, self.ct))
        return out

    def rif_eval(self, h, p, what=2):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        f = np.zeros(n, dtype=self.dtype)
        g = np.zeros((n, 3), dtype=self.dtype)
        self._fn("orc_rif_eval")(h, C.c_int(what), C.c_size_t(n), _ptr(p, self.ct), _ptr(f, self.ct),
                                 _ptr(g, self.ct))
        return f, g

    def rif_eval_hessian(self, h, p):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        f = np.zeros(n, dtype=self.dtype)
        g = np.zeros((n, 3), dtype=self.dtype)
        H = np.zeros((n, 3, 3), dtype=self.dtype)
        self._fn("orc_rif_eval_hessian")(h, C.c_size_t(n), _ptr(p, self.ct), _ptr(f, self.ct), _ptr(g, self.ct),
                                         _ptr(H, self.ct))
        return f, g, H

    def rif_inside_limits(self, h, p):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        out = np.zeros(p.shape[0], dtype=np.uint8)
        self._fn("orc_rif_inside_limits")(h, C.c_size_t(p.shape[0]), _ptr(p, self.ct), _ptr(out, C.c_uint8))
        return out.astype(bool)

    def grid_create(self, desc, data):
        data = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        return C.c_void_p(self.lib.orc_grid_create(C.byref(desc), _ptr(data, C.c_float)))

    def mi_weight(self, pdf_a, pdf_b):
        """the restated power heuristic the MIS connections use (volpath.cpp:430-433)"""
        a = np.ascontiguousarray(pdf_a, dtype=np.float32).reshape(-1)
        b = np.ascontiguousarray(pdf_b, dtype=np.float32).reshape(-1)
        out = np.zeros(a.shape[0], np.float32)
        self.lib.orc_mi_weight(C.c_size_t(a.shape[0]), _ptr(a, C.c_float), _ptr(b, C.c_float), _ptr(out, C.c_float))
        return out

    def grid_create_spectrum(self, desc, rgb):
        """3-channel grid, rgb[z, y, x, 3] (gridvolume.cpp:293-329, 401-421)"""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1)
        return C.c_void_p(self.lib.orc_grid_create_spectrum(C.byref(desc), _ptr(rgb, C.c_float)))

    def grid_lookup_spectrum(self, h, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros((p.shape[0], 3), dtype=np.float32)
        self.lib.orc_grid_lookup_spectrum(h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_float))
        return out

    def grid_destroy(self, h):
        self.lib.orc_grid_destroy(h)

    def grid_lookup(self, h, p):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(p.shape[0], dtype=np.float32)
        self.lib.orc_grid_lookup(h, C.c_size_t(p.shape[0]), _ptr(p, C.c_float), _ptr(out, C.c_float))
        return out

    # ---- phase
    def hg_sample(self, g, wi, xi):
        wi = np.ascontiguousarray(wi, dtype=np.float32).reshape(-1, 3)
        xi = np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
        n = wi.shape[0]
        wo = np.zeros((n, 3), dtype=np.float32)
        pdf = np.zeros(n, dtype=np.float32)
        self.lib.orc_hg_sample(C.c_float(g), C.c_size_t(n), _ptr(wi, C.c_float), _ptr(xi, C.c_float),
                               _ptr(wo, C.c_float), _ptr(pdf, C.c_float))
        return wo, pdf

    def hg_eval(self, g, wi, wo):
        wi = np.ascontiguousarray(wi, dtype=np.float32).reshape(-1, 3)
        wo = np.ascontiguousarray(wo, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(wi.shape[0], dtype=np.float32)
        self.lib.orc_hg_eval(C.c_float(g), C.c_size_t(wi.shape[0]), _ptr(wi, C.c_float), _ptr(wo, C.c_float),
                             _ptr(out, C.c_float))
        return out

    def maxexp(self, sigma_t, what, x):
        return _maxexp(self.lib.orc_maxexp, sigma_t, what, x)

    def coordinate_system(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3)
        b, c = np.zeros_like(a), np.zeros_like(a)
        self.lib.orc_coordinate_system(C.c_size_t(a.shape[0]), _ptr(a, C.c_float), _ptr(b, C.c_float), _ptr(c, C.c_float))
        return b, c

    def fresnel_dielectric_ext(self, cos_i, eta):
        cos_i = np.ascontiguousarray(cos_i, dtype=np.float32).reshape(-1)
        eta = np.ascontiguousarray(eta, dtype=np.float32).reshape(-1)
        F, ct = np.zeros_like(cos_i), np.zeros_like(cos_i)
        self.lib.orc_fresnel_dielectric_ext(C.c_size_t(cos_i.size), _ptr(cos_i, C.c_float), _ptr(eta, C.c_float), _ptr(F, C.c_float),
                                            _ptr(ct, C.c_float))
        return F, ct

    # ---- medium
    def medium_create(self, desc, rif, density=None):
        return C.c_void_p(self._fn("orc_medium_create")(C.byref(desc), rif, density))

    def medium_destroy(self, h):
        self._fn("orc_medium_destroy")(h)

    def rif_eval_hessian_world(self, h, p):
        p = np.ascontiguousarray(p, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        f, g, H = np.zeros(n, self.dtype), np.zeros((n, 3), self.dtype), np.zeros((n, 3, 3), self.dtype)
        self._fn("orc_rif_eval_hessian_world")(h, C.c_size_t(n), _ptr(p, self.ct), _ptr(f, self.ct), _ptr(g, self.ct), _ptr(H, self.ct))
        return f, g, H

    def derivative_trace(self, h, p, v, nsteps):
        p = np.array(p, dtype=self.dtype).reshape(-1, 3)
        v = np.array(v, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        ns = np.ascontiguousarray(np.broadcast_to(np.asarray(nsteps, np.int32), (n,)))
        A, B = np.zeros((n, 3, 3), self.dtype), np.zeros((n, 3, 3), self.dtype)
        self._fn("orc_medium_derivative_trace")(h, C.c_size_t(n), _ptr(p, self.ct), _ptr(v, self.ct), _ptr(ns, C.c_int32),
                                                _ptr(A, self.ct), _ptr(B, self.ct))
        return dict(p=p, v=v, dpdv0=A, dvdv0=B)

    def connection_residual(self, h, p1, p2, v0, is_sensor=False):
        p1 = np.ascontiguousarray(p1, dtype=self.dtype).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=self.dtype).reshape(-1, 3)
        v0 = np.ascontiguousarray(v0, dtype=self.dtype).reshape(-1, 3)
        n = p1.shape[0]
        err, J = np.zeros((n, 3), self.dtype), np.zeros((n, 3, 3), self.dtype)
        st, ns = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._fn("orc_medium_connection_residual")(h, C.c_size_t(n), _ptr(p1, self.ct), _ptr(p2, self.ct), _ptr(v0, self.ct),
                                                   C.c_int(1 if is_sensor else 0), _ptr(err, self.ct), _ptr(J, self.ct),
                                                   _ptr(st, C.c_int32), _ptr(ns, C.c_int32))
        return dict(error=err, derror=J, status=st, nsteps=ns)

    def connect(self, h, p1, p2, dseed, is_sensor=False, tol2=1e-6, rrweight=1e-2, precision=3, max_iterations=20, seed=1, start_mode=0):
        p1 = np.ascontiguousarray(p1, dtype=self.dtype).reshape(-1, 3)
        p2 = np.ascontiguousarray(p2, dtype=self.dtype).reshape(-1, 3)
        ds = np.ascontiguousarray(dseed, dtype=self.dtype).reshape(-1, 3)
        n = p1.shape[0]
        r = dict(success=np.zeros(n, np.uint8), dir_to_p2=np.zeros((n, 3), self.dtype), rev_dir=np.zeros((n, 3), self.dtype),
                 optical_dist=np.zeros(n, self.dtype), dist=np.zeros(n, self.dtype), weight=np.zeros(n, self.dtype),
                 transmittance=np.zeros((n, 3), np.float32), pdf_success=np.zeros(n, np.float32), pdf_failure=np.zeros(n, np.float32),
                 evaluations=np.zeros(n, np.int32))
        self._fn("orc_medium_connect")(h, C.c_size_t(n), _ptr(p1, self.ct), _ptr(p2, self.ct), _ptr(ds, self.ct), C.c_int(1 if is_sensor else 0),
                                       C.c_float(tol2), C.c_float(rrweight), C.c_int(precision), C.c_int(max_iterations), C.c_uint64(seed),
                                       _ptr(r["success"], C.c_uint8), _ptr(r["dir_to_p2"], self.ct), _ptr(r["rev_dir"], self.ct),
                                       _ptr(r["optical_dist"], self.ct), _ptr(r["dist"], self.ct), _ptr(r["weight"], self.ct),
                                       _ptr(r["transmittance"], C.c_float), _ptr(r["pdf_success"], C.c_float), _ptr(r["pdf_failure"], C.c_float),
                                       _ptr(r["evaluations"], C.c_int32), C.c_int(start_mode))
        r["success"] = r["success"].astype(bool)
        return r

    def medium_set_albedo_grid(self, h, grid):
        self._fn("orc_medium_set_albedo_grid")(h, grid)

    def medium_set_sdf(self, h, sdf, aggressive=True):
        self._fn("orc_medium_set_sdf")(h, sdf, C.c_int(1 if aggressive else 0))

    def grid_sample_distance(self, grid, desc, scale, ro, rd, mint, maxt, seed):
        ro = np.ascontiguousarray(ro, dtype=np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rd, dtype=np.float32).reshape(-1, 3)
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, np.float32), (n,)))
        maxt = np.ascontiguousarray(np.broadcast_to(np.asarray(maxt, np.float32), (n,)))
        ok, t, dens = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.lib.orc_grid_sample_distance(grid, C.byref(desc), C.c_float(scale), C.c_size_t(n), _ptr(ro, C.c_float),
                                          _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(maxt, C.c_float), C.c_uint64(seed),
                                          _ptr(ok, C.c_uint8), _ptr(t, C.c_float), _ptr(dens, C.c_float))
        return ok.astype(bool), t, dens

    def grid_eval_transmittance(self, grid, desc, scale, ro, rd, mint, maxt, seed):
        ro = np.ascontiguousarray(ro, dtype=np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rd, dtype=np.float32).reshape(-1, 3)
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, np.float32), (n,)))
        maxt = np.ascontiguousarray(np.broadcast_to(np.asarray(maxt, np.float32), (n,)))
        out = np.zeros(n, np.float32)
        self.lib.orc_grid_eval_transmittance(grid, C.byref(desc), C.c_float(scale), C.c_size_t(n), _ptr(ro, C.c_float),
                                             _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(maxt, C.c_float),
                                             C.c_uint64(seed), _ptr(out, C.c_float))
        return out

    def medium_resolved(self, h):
        w, sd = C.c_float(), C.c_float()
        self._fn("orc_medium_resolved")(h, C.byref(w), C.byref(sd))
        return w.value, sd.value

    def trace(self, h, p, v, dist):
        p = np.array(p, dtype=self.dtype).reshape(-1, 3)
        v = np.array(v, dtype=self.dtype).reshape(-1, 3)
        dist = np.ascontiguousarray(dist, dtype=self.dtype).reshape(-1)
        n = p.shape[0]
        ok = np.zeros(n, dtype=np.uint8)
        ds = np.zeros(n, dtype=self.dtype)
        opl = np.zeros(n, dtype=self.dtype)
        ns = np.zeros(n, dtype=np.int32)
        self._fn("orc_medium_trace")(h, C.c_size_t(n), _ptr(p, self.ct), _ptr(v, self.ct), _ptr(dist, self.ct),
                                     _ptr(ok, C.c_uint8), _ptr(ds, self.ct), _ptr(opl, self.ct),
                                     _ptr(ns, C.c_int32))
        return dict(p=p, v=v, success=ok.astype(bool), dist_surf=ds, opl=opl, nsteps=ns)

    def trace_till_boundary(self, h, p, v):
        p = np.array(p, dtype=self.dtype).reshape(-1, 3)
        v = np.array(v, dtype=self.dtype).reshape(-1, 3)
        n = p.shape[0]
        ds = np.zeros(n, dtype=self.dtype)
        opl = np.zeros(n, dtype=self.dtype)
        ns = np.zeros(n, dtype=np.int32)
        self._fn("orc_medium_trace_till_boundary")(h, C.c_size_t(n), _ptr(p, self.ct), _ptr(v, self.ct),
                                                   _ptr(ds, self.ct), _ptr(opl, self.ct), _ptr(ns, C.c_int32))
        return dict(p=p, v=v, dist_surf=ds, opl=opl, nsteps=ns)

    def sample_distance(self, h, ro, rd, mint, xi):
        ro = np.ascontiguousarray(ro, dtype=np.float32).reshape(-1, 3)
        rd = np.ascontiguousarray(rd, dtype=np.float32).reshape(-1, 3)
        n = ro.shape[0]
        mint = np.ascontiguousarray(np.broadcast_to(np.asarray(mint, dtype=np.float32), (n,)))
        xi = np.ascontiguousarray(xi, dtype=np.float32).reshape(-1, 2)
        r = dict(success=np.zeros(n, np.uint8), t=np.zeros(n, self.dtype), p=np.zeros((n, 3), self.dtype),
                 d=np.zeros((n, 3), self.dtype), optical_length=np.zeros(n, self.dtype),
                 ref_ratio_sq=np.zeros(n, self.dtype), transmittance=np.zeros((n, 3), np.float32),
                 pdf_success=np.zeros(n, np.float32), pdf_failure=np.zeros(n, np.float32),
                 sigma_s=np.zeros((n, 3), np.float32), nsteps=np.zeros(n, np.int32))
        self._fn("orc_medium_sample_distance")(
            h, C.c_size_t(n), _ptr(ro, C.c_float), _ptr(rd, C.c_float), _ptr(mint, C.c_float), _ptr(xi, C.c_float),
            _ptr(r["success"], C.c_uint8), _ptr(r["t"], self.ct), _ptr(r["p"], self.ct), _ptr(r["d"], self.ct),
            _ptr(r["optical_length"], self.ct), _ptr(r["ref_ratio_sq"], self.ct),
            _ptr(r["transmittance"], C.c_float), _ptr(r["pdf_success"], C.c_float),
            _ptr(r["pdf_failure"], C.c_float), _ptr(r["sigma_s"], C.c_float), _ptr(r["nsteps"], C.c_int32))
        r["success"] = r["success"].astype(bool)
        return r

    def eval_transmittance(self, sigma_t, mint, maxt):
        mint = np.ascontiguousarray(mint, dtype=np.float32).reshape(-1)
        maxt = np.ascontiguousarray(maxt, dtype=np.float32).reshape(-1)
        out = np.zeros((mint.size, 3), dtype=np.float32)
        st = (C.c_float * 3)(*[float(x) for x in sigma_t])
        self.lib.orc_eval_transmittance(st, C.c_size_t(mint.size), _ptr(mint, C.c_float), _ptr(maxt, C.c_float),
                                        _ptr(out, C.c_float))
        return out

    # ---- integrator
    def render(self, medium, rdesc, nthreads=0):
        frames = max(int(rdesc.frames), 1) if not rdesc.modulation else 1
        film = np.zeros((rdesc.height, rdesc.width, 3 * frames + 2), dtype=np.float32)
        stats = RenderStats()
        self._fn("orc_render")(medium, C.byref(rdesc), _ptr(film, C.c_float), C.byref(stats), C.c_int(nthreads))
        return film, stats

    def philox(self, seed, sample_id, n):
        out = np.zeros(n, dtype=np.float32)
        self.lib.orc_philox(C.c_uint64(seed), C.c_uint64(sample_id), C.c_size_t(n), _ptr(out, C.c_float))
        return out

    def filter_table(self, ftype):
        vals = np.zeros(32, dtype=np.float32)
        r, s = C.c_float(), C.c_float()
        self.lib.orc_filter_table(C.c_int(ftype), _ptr(vals, C.c_float), C.byref(r), C.byref(s))
        return vals, r.value, s.value

    def film_put(self, ftype, W, H, pos, values):
        """ImageBlock::put of n samples (pos [n][2], values [n][channels]) into a zeroed film, in order -> film [H][W][channels], ok [n]"""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 2)
        values = np.ascontiguousarray(values, dtype=np.float32).reshape(pos.shape[0], -1)
        film = np.zeros((H, W, values.shape[1]), np.float32)
        ok = np.zeros(pos.shape[0], np.int32)
        self.lib.orc_film_put(C.c_int(ftype), C.c_int(W), C.c_int(H), C.c_int(values.shape[1]), C.c_size_t(pos.shape[0]), _ptr(pos, C.c_float),
                              _ptr(values, C.c_float), _ptr(film, C.c_float), _ptr(ok, C.c_int))
        return film, ok.astype(bool)

    def hdielectric_sample(self, d, N, eta, u, importance=False):
        """HSmoothDielectric::sample (hdielectric.cpp:244-300) for rays along d at a surface with normal N and interior index eta
        -> (direction out, weight, relative index of the event, transmitted?)"""
        d = np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 3)
        N = np.ascontiguousarray(N, dtype=np.float32).reshape(-1, 3)
        n = d.shape[0]
        eta = np.ascontiguousarray(np.broadcast_to(np.asarray(eta, np.float32), (n,)))
        u = np.ascontiguousarray(u, dtype=np.float32).reshape(-1)
        out, w, es, tr = np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.orc_hdielectric_sample(C.c_size_t(n), _ptr(d, C.c_float), _ptr(N, C.c_float), _ptr(eta, C.c_float), _ptr(u, C.c_float), C.c_int(1 if importance else 0),
                    _ptr(out, C.c_float), _ptr(w, C.c_float), _ptr(es, C.c_float), _ptr(tr, C.c_int))
        return out, w, es, tr.astype(bool)

    def camera_ray(self, rdesc, sample_pos):
        sp = np.ascontiguousarray(sample_pos, dtype=np.float32).reshape(-1, 2)
        d = np.zeros((sp.shape[0], 3), dtype=np.float32)
        self.lib.orc_camera_ray(C.byref(rdesc), C.c_size_t(sp.shape[0]), _ptr(sp, C.c_float), _ptr(d, C.c_float))
        return d

    def film_develop(self, film):
        """[H][W][3*frames+2] -> [H][W][3] (steady state) or [H][W][frames][3]"""
        H, W, ch = film.shape
        frames = (ch - 2) // 3
        film = np.ascontiguousarray(film, dtype=np.float32)
        rgb = np.zeros((H, W, frames, 3), dtype=np.float32)
        self.lib.orc_film_develop_frames(C.c_int(W), C.c_int(H), C.c_int(frames), _ptr(film, C.c_float), _ptr(rgb, C.c_float))
        return rgb[:, :, 0, :] if frames == 1 else rgb
