"""ctypes binding of libmitsubaer_b200.so (include/mitsubaer_b200.h).

This is the same stub a Mitsuba-side maintainer would write in C++;
here it is Python because the tests and bench.py are.  There is NO fallback: if the shared
library is missing this module raises at import, and every compute entry point raises
`MerError` when no sm_100-class GPU is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MER_LIB") or os.path.join(_HERE, "libmitsubaer_b200.so")  # MER_LIB: a development build to compare

MER_OK, MER_ERR_INVALID, MER_ERR_CUDA, MER_ERR_UNSUPPORTED, MER_ERR_OOM = 0, 1, 2, 3, 4
RIF_TRICUBIC, RIF_TRILINEAR_PACKED = 0, 1
EVAL_VALUE, EVAL_GRADIENT, EVAL_VALUE_AND_GRADIENT = 0, 1, 2
SHAPE_BOX, SHAPE_SPHERE, SHAPE_SDF = 0, 1, 2
STRATEGY_BALANCE, STRATEGY_SINGLE, STRATEGY_MANUAL, STRATEGY_MAXIMUM = 0, 1, 2, 3
FILTER_BOX, FILTER_GAUSSIAN = 0, 1
BOUNDARY_INDEX_MATCHED, BOUNDARY_HDIELECTRIC = 0, 1
EMITTER_QUAD, EMITTER_COLLIMATED = 0, 1


class MerError(RuntimeError):
    """Raised for any non-zero mer_status (the analogue of Log(EError) throwing)."""

    def __init__(self, code, message):
        super().__init__("mitsubaer_b200 error %d: %s" % (code, message))
        self.code = code


class VolumeDesc(C.Structure):
    _fields_ = [("res", C.c_int32 * 3), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("has_transform", C.c_int32), ("world_to_volume", C.c_float * 12)]


class MediumDesc(C.Structure):
    _fields_ = [("sigma_a", C.c_float * 3), ("sigma_s", C.c_float * 3), ("stepsize", C.c_float),
                ("medium_sampling_weight", C.c_float), ("strategy", C.c_int32), ("channel", C.c_int32),
                ("sampling_density", C.c_float), ("shape_type", C.c_int32), ("shape", C.c_float * 6),
                ("hg_g", C.c_float), ("density_scale", C.c_float), ("albedo", C.c_float * 3), ("boundary", C.c_int32),
                ("radiance_scaling", C.c_int32)]


class SamplingRecords(C.Structure):
    _fields_ = [("success", C.POINTER(C.c_uint8)), ("t", C.POINTER(C.c_float)), ("p", C.POINTER(C.c_float)),
                ("d", C.POINTER(C.c_float)), ("optical_length", C.POINTER(C.c_float)),
                ("ref_ratio_sq", C.POINTER(C.c_float)), ("transmittance", C.POINTER(C.c_float)),
                ("pdf_success", C.POINTER(C.c_float)), ("pdf_failure", C.POINTER(C.c_float)),
                ("sigma_s", C.POINTER(C.c_float)), ("nsteps", C.POINTER(C.c_int32))]


class ConnectionParams(C.Structure):
    _fields_ = [("tol2", C.c_float), ("rrweight", C.c_float), ("boundary_precision", C.c_int32), ("max_iterations", C.c_int32),
                ("start_mode", C.c_int32)]


class ConnectionRecords(C.Structure):
    _fields_ = [("success", C.POINTER(C.c_uint8)), ("dir_to_p2", C.POINTER(C.c_float)), ("rev_dir_to_p1", C.POINTER(C.c_float)),
                ("optical_length", C.POINTER(C.c_float)), ("distance", C.POINTER(C.c_float)), ("weight", C.POINTER(C.c_float)),
                ("transmittance", C.POINTER(C.c_float)), ("pdf_success", C.POINTER(C.c_float)), ("pdf_failure", C.POINTER(C.c_float)),
                ("evaluations", C.POINTER(C.c_int32))]


class RenderDesc(C.Structure):
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
    _fields_ = [("samples", C.c_uint64), ("ray_steps", C.c_uint64), ("scatter_events", C.c_uint64),
                ("null_collisions", C.c_uint64), ("boundary_exits", C.c_uint64),
                ("nonfinite_dropped", C.c_uint64), ("passes", C.c_uint64), ("connections", C.c_uint64),
                ("connections_failed", C.c_uint64), ("connection_steps", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_ms", C.c_float), ("step_kernel_ms", C.c_float), ("step_launches", C.c_uint64), ("tail_ms", C.c_float),
                ("block_fetches", C.c_uint64), ("step_lanes_per_sm", C.c_uint32), ("reserved0", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_fp = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "mer_rif_create": (C.c_int, [C.c_int, C.POINTER(VolumeDesc), _fp, C.c_int, C.POINTER(_vp)]),
    "mer_rif_create_device": (C.c_int, [C.c_int, C.POINTER(VolumeDesc), _vp, C.c_int, C.POINTER(_vp)]),
    "mer_rif_create_from_file": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(VolumeDesc), C.c_int, C.POINTER(_vp)]),
    "mer_rif_destroy": (None, [_vp]),
    "mer_rif_coefficients": (C.c_int, [_vp, _fp]),
    "mer_rif_desc": (C.c_int, [_vp, C.POINTER(VolumeDesc), C.POINTER(C.c_int)]),
    "mer_rif_eval_batch": (C.c_int, [_vp, C.c_int, C.c_size_t, _fp, _fp, _fp]),
    "mer_rif_eval_device": (C.c_int, [_vp, C.c_int, C.c_size_t, _vp, _vp, _vp, _vp]),
    "mer_rif_inside_limits_batch": (C.c_int, [_vp, C.c_size_t, _fp, _u8p]),
    "mer_grid_create": (C.c_int, [C.c_int, C.POINTER(VolumeDesc), _fp, C.POINTER(_vp)]),
    "mer_grid_create_device": (C.c_int, [C.c_int, C.POINTER(VolumeDesc), _vp, C.POINTER(_vp)]),
    "mer_grid_create_from_file": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(VolumeDesc), C.POINTER(_vp)]),
    "mer_grid_destroy": (None, [_vp]),
    "mer_grid_lookup_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp]),
    "mer_grid_create_spectrum": (C.c_int, [C.c_int, C.POINTER(VolumeDesc), _fp, C.POINTER(_vp)]),
    "mer_grid_channels": (C.c_int, [_vp]),
    "mer_grid_lookup_spectrum_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp]),
    "mer_grid_sample_distance_batch": (C.c_int, [_vp, C.c_float, C.c_size_t, _fp, _fp, _fp, _fp, C.c_uint64, _u8p, _fp, _fp]),
    "mer_grid_eval_transmittance_batch": (C.c_int, [_vp, C.c_float, C.c_size_t, _fp, _fp, _fp, _fp, C.c_uint64, _fp]),
    "mer_vol_read_header": (C.c_int, [C.c_char_p, C.POINTER(VolumeDesc), _i32p, _i32p]),
    "mer_vol_read_data": (C.c_int, [C.c_char_p, _fp, C.c_size_t]),
    "mer_vol_write": (C.c_int, [C.c_char_p, C.POINTER(VolumeDesc), _fp]),
    "mer_vol_write_spectrum": (C.c_int, [C.c_char_p, C.POINTER(VolumeDesc), _fp]),
    "mer_hg_sample_batch": (C.c_int, [C.c_int, C.c_float, C.c_size_t, _fp, _fp, _fp, _fp]),
    "mer_hg_eval_batch": (C.c_int, [C.c_int, C.c_float, C.c_size_t, _fp, _fp, _fp]),
    "mer_medium_create": (C.c_int, [C.POINTER(MediumDesc), _vp, _vp, C.POINTER(_vp)]),
    "mer_medium_destroy": (None, [_vp]),
    "mer_medium_set_sdf": (C.c_int, [_vp, _vp, C.c_int]),
    "mer_medium_set_albedo_grid": (C.c_int, [_vp, _vp]),
    "mer_medium_resolved": (C.c_int, [_vp, C.POINTER(MediumDesc), _fp]),
    "mer_medium_trace_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _fp, _u8p, _fp, _fp, _i32p]),
    "mer_medium_trace_device": (C.c_int, [_vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mer_medium_trace_counted_device": (C.c_int, [_vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mer_medium_trace_till_boundary_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _fp, _fp, _i32p]),
    "mer_medium_sample_distance_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _fp, _fp, C.POINTER(SamplingRecords)]),
    "mer_medium_eval_transmittance_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _fp]),
    "mer_rif_eval_hessian_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _fp, _fp]),
    "mer_medium_derivative_trace_batch": (C.c_int, [_vp, C.c_size_t, _fp, _fp, _i32p, _fp, _fp]),
    "mer_medium_connection_residual_batch": (C.c_int, [_vp, C.c_int, C.c_size_t, _fp, _fp, _fp, C.c_int, _fp, _fp, _i32p, _i32p]),
    "mer_medium_connect_batch": (C.c_int, [_vp, C.POINTER(ConnectionParams), C.c_size_t, _fp, _fp, _fp, C.c_int, C.c_uint64,
                                           C.POINTER(ConnectionRecords)]),
    "mer_render": (C.c_int, [_vp, C.POINTER(RenderDesc), _fp, C.POINTER(RenderStats)]),
    "mer_render_device": (C.c_int, [_vp, C.POINTER(RenderDesc), _vp, C.POINTER(RenderStats), _vp]),
    "mer_film_develop": (C.c_int, [C.c_int, C.c_int32, C.c_int32, _fp, _fp]),
    "mer_film_develop_frames": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, _fp, _fp]),
    "mer_last_error": (C.c_char_p, []),
    "mer_abi_version": (C.c_int, []),
    "mer_render_multi": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp]),
    "mer_device_count": (C.c_int, []),
    "mer_kernel_launch_count": (C.c_uint64, []),
    "mer_trim_memory": (C.c_int, [C.c_int]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(mitsubaer_b200 has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so does not match the header
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


class _Lib:
    """The shared library, loaded on first use.  `import mitsubaer_b200` touches it right away (so a missing or
    stale .so fails at import, loudly) unless MER_B200_DEFER_LOAD=1, which bench.py's CPU-reference arm sets so
    that the process timing the CPU baseline never maps the CUDA library (it only needs fields.py's generators)."""
    _handle = None

    def __getattr__(self, name):
        if _Lib._handle is None:
            _Lib._handle = _load()
        return getattr(_Lib._handle, name)


lib = _Lib()


def check(rc):
    if rc != MER_OK:
        raise MerError(rc, (lib.mer_last_error() or b"").decode("utf-8", "replace"))
