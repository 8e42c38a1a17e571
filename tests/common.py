"""Shared scene builders for the parity tests: the same synthetic fields are fed to the CUDA
path (through the C ABI) and to the CPU oracle."""
import numpy as np

from mitsubaer_b200 import fields

BOX_MIN = np.array([-1.0, -1.0, -1.0], np.float32)
BOX_MAX = np.array([1.0, 1.0, 1.0], np.float32)


def make_field(kind, res, seed=0):
    """-> (data[z][y][x], bbox_min, bbox_max) with the unit box 3 voxels inside the bbox"""
    res = (res, res, res) if np.isscalar(res) else tuple(res)
    lo, hi = fields.padded_bbox(BOX_MIN, BOX_MAX, res)
    if kind == "linear":
        data = fields.linear_rif(res, lo, hi)
    elif kind == "radial":
        data = fields.radial_rif(res, lo, hi)
    elif kind == "sd":
        data = fields.rif_from_sd(fields.sphere_sdf(res, lo, hi, radius=0.8))
    elif kind == "random":
        rng = np.random.default_rng(seed)
        data = (1.2 + 0.3 * rng.random((res[2], res[1], res[0]))).astype(np.float32)
    elif kind == "smooth":  # smooth non-symmetric field
        x = np.linspace(lo[0], hi[0], res[0])[None, None, :]
        y = np.linspace(lo[1], hi[1], res[1])[None, :, None]
        z = np.linspace(lo[2], hi[2], res[2])[:, None, None]
        data = (1.4 + 0.1 * np.sin(2.1 * x + 0.3) * np.cos(1.7 * y) + 0.05 * np.sin(3.0 * z + x * y)).astype(np.float32)
    else:
        raise ValueError(kind)
    return data, lo, hi


def random_points_in_box(n, seed, margin=0.0):
    rng = np.random.default_rng(seed)
    return (BOX_MIN + margin + rng.random((n, 3)) * (BOX_MAX - BOX_MIN - 2 * margin)).astype(np.float32)


def random_directions(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def medium_props(**over):
    p = dict(sigmaS=3.6, sigmaA=0.4, stepsize=2e-3, strategy="single", shape=("box", BOX_MIN, BOX_MAX))
    p.update(over)
    return p


def oracle_medium_desc(props, g=0.9, has_density=False):
    """the same resolution of properties as plugins.HeterogeneousRefractiveMedium.configure, for the oracle"""
    from oracle.oracle import MediumDesc
    d = MediumDesc()

    def spec(v):
        a = np.asarray(v, np.float32).reshape(-1)
        return np.repeat(a, 3) if a.size == 1 else a

    scale = float(props.get("scale", 1.0))
    if "sigmaT" in props and "albedo" in props:
        st, al = spec(props["sigmaT"]) * scale, spec(props["albedo"])
        ss, sa = st * al, st * (1 - al)
    else:
        ss, sa = spec(props.get("sigmaS", 0.0)) * scale, spec(props.get("sigmaA", 0.0)) * scale
    d.sigma_a[:] = [float(x) for x in sa]
    d.sigma_s[:] = [float(x) for x in ss]
    d.stepsize = float(props.get("stepsize", 1e-3))
    d.medium_sampling_weight = float(props.get("mediumSamplingWeight", -1))
    d.strategy = {"balance": 0, "single": 1, "manual": 2, "maximum": 3}[props.get("strategy", "balance")]
    d.channel = int(props.get("channel", -1))
    d.sampling_density = float(props.get("samplingDensity", 0.0))
    shape = props["shape"]
    if shape[0] == "box":
        d.shape_type = 0
        d.shape[:] = [float(x) for x in list(shape[1]) + list(shape[2])]
    elif shape[0] == "sdf":
        d.shape_type = 2
        d.shape[:] = [float(x) for x in list(shape[1]) + list(shape[2])]
    else:
        d.shape_type = 1
        d.shape[:] = [float(x) for x in list(shape[1]) + [shape[2], 0.0, 0.0]]
    d.hg_g = g
    d.density_scale = float(props.get("densityScale", props.get("scale", 1.0))) if has_density else 0.0
    d.albedo[:] = [float(x) for x in spec(props.get("albedo", 0.0))]
    d.boundary = 1 if props.get("bsdf", "null") == "hdielectric" else 0
    d.radiance_scaling = 1 if props.get("radianceScaling", "reference") == "physical" else 0
    return d


def scene_dict(width=64, height=64, spp=16, rfilter="gaussian", quad=True, seed=20201201):
    s = dict(width=width, height=height, sampleCount=spp, seed=seed, origin=(0.0, 0.0, -4.0), target=(0.0, 0.0, 0.0),
             up=(0.0, 1.0, 0.0), fov=40.0, rfilter=rfilter, envRadiance=1.0)
    if quad:
        s["quad"] = dict(origin=(-0.5, 1.5, -0.5), u=(1.0, 0.0, 0.0), v=(0.0, 0.0, 1.0), radiance=(8.0, 6.0, 4.0))
    return s


def oracle_render_desc(scene, max_depth=-1, rr_depth=5, sample_begin=0, sample_stride=1, direct_connections=False, props=None,
                       start="straight", light_tracing=False):
    from oracle.oracle import RenderDesc
    r = RenderDesc()
    r.width, r.height, r.spp_total = scene["width"], scene["height"], scene["sampleCount"]
    r.sample_begin, r.sample_stride = sample_begin, sample_stride
    r.seed = scene["seed"]
    r.cam_origin[:] = scene["origin"]
    r.cam_target[:] = scene["target"]
    r.cam_up[:] = scene["up"]
    r.fov_deg = scene["fov"]
    r.filter = {"box": 0, "gaussian": 1}[scene["rfilter"]]
    r.max_depth, r.rr_depth = max_depth, rr_depth
    env = np.asarray(scene["envRadiance"], np.float32).reshape(-1)
    r.env_radiance[:] = [float(x) for x in (np.repeat(env, 3) if env.size == 1 else env)]
    q = scene.get("quad")
    r.has_quad = 1 if q else 0
    if q:
        r.quad_origin[:] = q["origin"]
        r.quad_u[:] = q["u"]
        r.quad_v[:] = q["v"]
        r.quad_radiance[:] = q["radiance"]
    r.direct_connections = 2 if direct_connections == "mis" else (1 if direct_connections else 0)
    r.light_tracing = 1 if light_tracing else 0
    props = props or {}
    r.connection.tol2 = float(props.get("tol2", 1e-6))
    r.connection.rrweight = float(props.get("rrweight", 1e-2))
    r.connection.boundary_precision = int(props.get("boundaryprecision", 3))
    r.connection.max_iterations = int(props.get("ceresmaxiterations", 20))
    r.connection.start_mode = 1 if start == "random" else 2
    em = scene.get("emitter")
    if em and em.get("type", "quad") == "collimated":
        r.emitter_type = 1
        r.beam_origin[:] = em.get("origin", (0, 0, 0))
        r.beam_direction[:] = em.get("direction", (0, 0, 1))
        pw = np.asarray(em.get("power", 1.0), np.float32).reshape(-1)
        r.beam_power[:] = [float(x) for x in (np.repeat(pw, 3) if pw.size == 1 else pw)]
    tr = scene.get("transient")  # dict(minBound, maxBound, binWidth[, calibrated]) as on <film>
    if tr:
        r.frames = int(np.ceil((tr["maxBound"] - tr["minBound"]) / tr["binWidth"]))
        r.min_bound, r.bin_width = float(tr["minBound"]), float(tr["binWidth"])
        r.calibrated_transient = 1 if tr.get("calibrated", False) else 0
        r.modulation = {"none": 0, "sine": 1, "square": 2, "hamiltonian": 3}[tr.get("modulation", "none")]
        r.lambda_ = float(tr.get("lambda", 1.0))
        r.phase_deg = float(tr.get("phase", 0.0))
    return r


# ---- the reference's own container: insideShape() is hackForSphere(), a hard-coded sphere (heterogeneousrefractive.cpp:707-718)
REF_SPHERE_CENTRE = np.array([-0.22827, 1.2, 0.152505], np.float32)
REF_SPHERE_RADIUS = 0.3


def ref_sphere_scene(kind, res=40, n_rays=4000, seed=7):
    """A field on a grid around the reference's hard-coded sphere and rays that start inside it:
    -> (data[z][y][x], bbox_min, bbox_max, p0, unit directions, distances).  The caller scales the directions by n(p0)."""
    c = REF_SPHERE_CENTRE.astype(np.float64)
    lo, hi = fields.padded_bbox(c - 0.32, c + 0.32, (res,) * 3)
    if kind == "linear":
        data = fields.linear_rif((res,) * 3, lo, hi)
    elif kind == "radial":
        data = fields.radial_rif((res,) * 3, lo, hi)
    elif kind == "sd":
        data = fields.rif_from_sd(fields.sphere_sdf((res,) * 3, lo, hi, centre=tuple(c), radius=0.28))
    else:
        x = np.linspace(lo[0], hi[0], res)[None, None, :]
        y = np.linspace(lo[1], hi[1], res)[None, :, None]
        z = np.linspace(lo[2], hi[2], res)[:, None, None]
        data = (1.4 + 0.1 * np.sin(9.0 * x + 0.3) * np.cos(7.0 * y) + 0.05 * np.sin(11.0 * z + 5.0 * x * y)).astype(np.float32)
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_rays, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    q = rng.normal(size=(n_rays, 3))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    p0 = (c + q * (0.95 * REF_SPHERE_RADIUS * rng.random((n_rays, 1)) ** (1 / 3))).astype(np.float32)
    dist = (0.7 * rng.random(n_rays)).astype(np.float32)  # up to ~1.2 diameters: about half of the rays leave the sphere
    return data, lo.astype(np.float32), hi.astype(np.float32), p0, d.astype(np.float32), dist
