/*
 * mitsubaer_b200.h — C ABI of libmitsubaer_b200.so
 *
 * The drop-in boundary for MitsubaER's refractive-radiative-transfer hot path
 * (SURVEY.md §8b).  Plain pointers and sizes only; no torch / CUDA types in the
 * signatures (a CUDA stream is passed as `void *`).  Every entry point names the
 * reference interface it replaces (paths relative to the MitsubaER tree).
 *
 * Conventions
 *   - all functions return MER_OK (0) or a MER_ERR_* code; the message is
 *     available from mer_last_error() (thread local).  The Mitsuba-side shim
 *     turns a non-zero status into Log(EError, ...) (src/libcore/logger.cpp:100-147).
 *   - points / vectors are packed float[n][3]; spectra are float[3] (SPECTRUM_SAMPLES=3).
 *   - `*_batch` entry points take HOST buffers and do the H2D/D2H copies themselves;
 *     `*_device` entry points take DEVICE pointers on the handle's GPU and enqueue
 *     on `stream` without synchronising.
 *   - handles own all device allocations; they are immutable after creation and may
 *     be used concurrently from several host threads (like Mitsuba's const plugins).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     MER_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef MITSUBAER_B200_H
#define MITSUBAER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MER_ABI_VERSION 10

enum mer_status {
    MER_OK = 0,
    MER_ERR_INVALID = 1,     /* bad argument / descriptor */
    MER_ERR_CUDA = 2,        /* CUDA runtime failure or no usable device */
    MER_ERR_UNSUPPORTED = 3, /* feature of the reference that this path does not carry */
    MER_ERR_OOM = 4
};

/* ------------------------------------------------------------------ volumes */

/* .vol grid description: src/volume/splinevolume.cpp:204-273 (header fields),
 * `toWorld` / `min` / `max` properties :87-98.  Data is x-fastest float32,
 * data[(z*yres + y)*xres + x]; grid points sit ON the bbox. */
typedef struct mer_volume_desc {
    int32_t res[3];
    float bbox_min[3];
    float bbox_max[3];
    int32_t has_transform;     /* 0: world == volume space */
    float world_to_volume[12]; /* row-major 3x4 affine (inverse of `toWorld`) */
} mer_volume_desc;

/* RIF fetch modes */
enum mer_rif_mode {
    MER_RIF_TRICUBIC = 0,        /* the reference's prefiltered cubic B-spline (parity mode) */
    MER_RIF_TRILINEAR_PACKED = 1 /* float4 {n, dn/dx, dn/dy, dn/dz} per voxel, trilinear (fast mode) */
};

enum mer_eval_what { MER_EVAL_VALUE = 0, MER_EVAL_GRADIENT = 1, MER_EVAL_VALUE_AND_GRADIENT = 2 };

typedef struct mer_rif mer_rif;   /* <volume type="splinevolume"> : SplineDataSource */
typedef struct mer_grid mer_grid; /* <volume type="gridvolume">   : GridDataSource (density) */

/* SplineDataSource ctor + loadFromFile + Spline<3>::build (splinevolume.cpp:87-111,
 * 204-317; basisspline.h:124-138, 865-890).  `data` is a HOST pointer to
 * res[0]*res[1]*res[2] float32; upload + IIR prefilter run on `device`. */
int mer_rif_create(int device, const mer_volume_desc *desc, const float *data, int mode, mer_rif **out);
/* same, `data` already resident on `device` (not modified, may be freed after return) */
int mer_rif_create_device(int device, const mer_volume_desc *desc, const float *data_dev, int mode,
                          mer_rif **out);
/* splinevolume.cpp:204-317 incl. the 48-byte header parse */
int mer_rif_create_from_file(int device, const char *vol_path, const mer_volume_desc *override_or_null,
                             int mode, mer_rif **out);
void mer_rif_destroy(mer_rif *rif);
/* prefiltered B-spline coefficients, same layout as the input (basisspline.h:865-890 `coeff`) */
int mer_rif_coefficients(const mer_rif *rif, float *coeff_out_host);
int mer_rif_desc(const mer_rif *rif, mer_volume_desc *out, int *mode_out);
/* SplineDataSource::value / gradient / valueAndGradient (splinevolume.cpp:330-360) in the handle's mode */
int mer_rif_eval_batch(const mer_rif *rif, int what, size_t n, const float *p, float *value_out,
                       float *grad_out);
int mer_rif_eval_device(const mer_rif *rif, int what, size_t n, const float *p_dev, float *value_dev,
                        float *grad_dev, void *stream);
/* SplineDataSource::insideVolumeLimits (splinevolume.cpp:319-328, limits :280-281) */
int mer_rif_inside_limits_batch(const mer_rif *rif, size_t n, const float *p, uint8_t *inside_out);

/* GridDataSource ctor + configure (src/volume/gridvolume.cpp:188-199) */
int mer_grid_create(int device, const mer_volume_desc *desc, const float *data, mer_grid **out);
int mer_grid_create_device(int device, const mer_volume_desc *desc, const float *data_dev, mer_grid **out);
int mer_grid_create_from_file(int device, const char *vol_path, const mer_volume_desc *override_or_null,
                              mer_grid **out);
void mer_grid_destroy(mer_grid *grid);
/* GridDataSource::lookupFloat (gridvolume.cpp:337-363); single-channel grids only (supportsFloatLookups, :578) */
int mer_grid_lookup_batch(const mer_grid *grid, size_t n, const float *p, float *value_out);
/* three-channel grids, the `albedo` child of a medium (heterogeneous.cpp:262-268): rgb[(z*yres + y)*xres + x][3],
 * the float3 payload of gridvolume.cpp:293-329.  mer_grid_create_from_file makes one from a 3-channel float32 or
 * uint8 .vol file (gridvolume.cpp:251-262); mer_grid_channels tells which kind a handle is (:578-579). */
int mer_grid_create_spectrum(int device, const mer_volume_desc *desc, const float *rgb, mer_grid **out);
int mer_grid_channels(const mer_grid *grid);
/* GridDataSource::lookupSpectrum (gridvolume.cpp:386-463), rgb_out[3*n] */
int mer_grid_lookup_spectrum_batch(const mer_grid *grid, size_t n, const float *p, float *rgb_out);

/* HeterogeneousMedium on STRAIGHT rays, Woodcock tracking (src/medium/heterogeneous.cpp:239-242 majorant =
 * scale * 1, :613-658 sampleDistance, :546-587 evalTransmittance with 2 samples; AABB clipping
 * include/mitsuba/core/aabb.h:308-338).  The Sampler of ray i is the Philox4x32-10 stream (seed, i).
 * This is the arithmetic the curved-ray composition inside mer_render is built from. */
int mer_grid_sample_distance_batch(const mer_grid *grid, float scale, size_t n, const float *ray_o, const float *ray_d,
                                   const float *ray_mint, const float *ray_maxt, uint64_t seed, uint8_t *success_out,
                                   float *t_out, float *density_at_t_out);
int mer_grid_eval_transmittance_batch(const mer_grid *grid, float scale, size_t n, const float *ray_o,
                                      const float *ray_d, const float *ray_mint, const float *ray_maxt, uint64_t seed,
                                      float *transmittance_out);

/* .vol v3 I/O (mfiles/writeGridToVol.m:1-36, splinevolume.cpp:204-273). */
int mer_vol_read_header(const char *path, mer_volume_desc *out, int32_t *encoding, int32_t *channels);
int mer_vol_read_data(const char *path, float *data_out, size_t n_floats);
int mer_vol_write(const char *path, const mer_volume_desc *desc, const float *data);
/* the same with three interleaved channels per voxel (an albedo grid; header field `channels` = 3, gridvolume.cpp:240-262) */
int mer_vol_write_spectrum(const char *path, const mer_volume_desc *desc, const float *rgb);

/* ------------------------------------------------------------ phase function */

/* HGPhaseFunction::sample (src/phase/hg.cpp:76-98): wi points back along the incoming
 * ray; xi = sampler->next2D(); returns wo and pdf = eval (hg.cpp:100-105). */
int mer_hg_sample_batch(int device, float g, size_t n, const float *wi, const float *xi, float *wo_out,
                        float *pdf_out);
/* HGPhaseFunction::eval (hg.cpp:107-110) */
int mer_hg_eval_batch(int device, float g, size_t n, const float *wi, const float *wo, float *value_out);

/* -------------------------------------------------------------------- medium */

enum mer_shape_type {
    MER_SHAPE_BOX = 0,   /* hackForBox form  (heterogeneousrefractive.cpp:722-726): min<=p<=max */
    MER_SHAPE_SPHERE = 1, /* hackForSphere    (:714-720): |p-c|^2 < r^2 */
    /* any closed shape given by the signed-distance grid of mer_medium_set_sdf (the `sdf` child volume, :376-380, made by
     * mfiles/createRIFFromSD.m): inside <=> sdf(p) < 0.  shape[0..5] is an axis-aligned box that bounds the shape (rays are
     * sphere-traced from where they enter it).  Stands in for the reference's mesh containment test, which needs libigl's
     * fast winding number (not vendored). */
    MER_SHAPE_SDF = 2
};

enum mer_strategy { /* heterogeneousrefractive.cpp:194-199 */
    MER_STRATEGY_BALANCE = 0,
    MER_STRATEGY_SINGLE = 1,
    MER_STRATEGY_MANUAL = 2,
    MER_STRATEGY_MAXIMUM = 3 /* MaxExpDist: MER_ERR_UNSUPPORTED */
};

/* Properties of <medium type="heterogeneousrefractive"> (heterogeneousrefractive.cpp:201-297)
 * plus the Medium base class (src/librender/medium.cpp:27-37) and the nested <phase type="hg">. */
typedef struct mer_medium_desc {
    float sigma_a[3];
    float sigma_s[3];
    float stepsize;               /* `stepsize`, absolute scene units, default 1e-3 */
    float medium_sampling_weight; /* `mediumSamplingWeight`, -1 => max(max albedo, 0.5) (:239-255) */
    int32_t strategy;             /* mer_strategy, default balance */
    int32_t channel;              /* `channel` for strategy single; -1 => smallest sigma_t */
    float sampling_density;       /* `samplingDensity` for strategy manual */
    int32_t shape_type;           /* containment predicate (insideShape, :707-739) */
    float shape[6];               /* box: min xyz, max xyz; sphere: centre xyz, radius */
    float hg_g;                   /* <phase type="hg"> g */
    /* optional density grid (new composition): when a grid is attached,
     * sigma_t(p) = density(p) * density_scale (heterogeneous.cpp:239-242, 613-658),
     * sigma_s = albedo * sigma_t, sampled by Woodcock tracking ALONG the curved ray. */
    float density_scale;
    float albedo[3];
    /* BSDF of the container surface (the medium's shape):
     *   MER_BOUNDARY_INDEX_MATCHED  null surface, rays pass straight through (volpath.cpp:287-296)
     *   MER_BOUNDARY_HDIELECTRIC    <bsdf type="hdielectric">: smooth dielectric whose eta is the RIF at the hit
     *                               point, exterior index 1 (src/bsdfs/hdielectric.cpp:115-125, 244-300;
     *                               fresnelDielectricExt src/libcore/util.cpp:665-695) */
    int32_t boundary;
    /* what a camera-side path edge multiplies into the throughput for the change of index along it:
     *   MER_SCALING_REFERENCE  refRatioSq = (n_end / n_start)^2, as heterogeneousrefractive.cpp:469,501 + edge.cpp:96-98
     *   MER_SCALING_PHYSICAL   (n_start / n_end)^2, what the invariance of L / n^2 along a ray asks for: with it a lossless
     *                          medium in an hdielectric container under a uniform environment renders as exactly that
     *                          environment */
    int32_t radiance_scaling;
} mer_medium_desc;

enum mer_radiance_scaling { MER_SCALING_REFERENCE = 0, MER_SCALING_PHYSICAL = 1 };

enum mer_boundary { MER_BOUNDARY_INDEX_MATCHED = 0, MER_BOUNDARY_HDIELECTRIC = 1 };

typedef struct mer_medium mer_medium; /* <medium type="heterogeneousrefractive"> */

/* ctor + addChild("rif") + configure (heterogeneousrefractive.cpp:201-297, 366-382, 1177-1193).
 * `density_or_null` attaches a <volume name="density">.  The medium keeps references to rif / density;
 * destroy them after the medium. */
int mer_medium_create(const mer_medium_desc *desc, const mer_rif *rif, const mer_grid *density_or_null,
                      mer_medium **out);
void mer_medium_destroy(mer_medium *medium);
/* addChild("sdf") + `aggressivetracing` (heterogeneousrefractive.cpp:230, 1177-1193): with aggressive != 0,
 * mer_medium_sample_distance_batch sphere-traces the signed-distance spline before the tested trace
 * (:476-493, aggressive_trace :697-704).  The sdf volume must have the RIF's AABB (:372-377).  mer_render
 * keeps using tested tracing only (containment there is the analytic box / sphere predicate, R5) and
 * returns MER_ERR_UNSUPPORTED for an aggressive medium. */
int mer_medium_set_sdf(mer_medium *medium, const mer_rif *sdf, int aggressive);
/* addChild("albedo") (heterogeneous.cpp:262-268): a spatially varying single-scattering albedo for a medium with a
 * density grid; scattering events weigh the path by lookupSpectrum(p) (:646-649) instead of the constant
 * mer_medium_desc.albedo.  NULL detaches.  The medium keeps a reference; destroy the grid after the medium. */
int mer_medium_set_albedo_grid(mer_medium *medium, const mer_grid *albedo_or_null);
/* resolved parameters (after the -1 defaults are applied) */
int mer_medium_resolved(const mer_medium *medium, mer_medium_desc *out, float *sampling_density_out);

/* trace() (heterogeneousrefractive.cpp:671-691): p, v in/out (v = n * dir), dist in;
 * success (1 = still inside after `dist`), dist_surf, opl (accumulated optical length, starts at 0),
 * nsteps (er_step calls incl. remainder and step-back) out.  Any output may be NULL. */
int mer_medium_trace_batch(const mer_medium *medium, size_t n, float *p, float *v, const float *dist,
                           uint8_t *success_out, float *dist_surf_out, float *opl_out, int32_t *nsteps_out);
int mer_medium_trace_device(const mer_medium *medium, size_t n, float *p_dev, float *v_dev,
                            const float *dist_dev, uint8_t *success_dev, float *dist_surf_dev,
                            float *opl_dev, int32_t *nsteps_dev, void *stream);
/* the same, and *block_fetches_dev (device memory, may be NULL) is INCREMENTED by the number of 4x4x4 coefficient blocks
 * (8 nodes in the packed mode) the stepper gathered from the grid: the algorithmic memory traffic of the C4 sweep */
int mer_medium_trace_counted_device(const mer_medium *medium, size_t n, float *p_dev, float *v_dev,
                                    const float *dist_dev, uint8_t *success_dev, float *dist_surf_dev,
                                    float *opl_dev, int32_t *nsteps_dev, uint64_t *block_fetches_dev, void *stream);
/* traceTillBoundary() (:742-776) */
int mer_medium_trace_till_boundary_batch(const mer_medium *medium, size_t n, float *p, float *v,
                                         float *dist_surf_out, float *opl_out, int32_t *nsteps_out);

/* MediumSamplingRecord (include/mitsuba/render/medium.h:33-109) as SoA over a batch */
typedef struct mer_medium_sampling_records {
    uint8_t *success;     /* return value of sampleDistance */
    float *t;             /* [n] */
    float *p;             /* [n][3] */
    float *d;             /* [n][3] n-scaled direction at the end of the curved edge */
    float *optical_length;/* [n] */
    float *ref_ratio_sq;  /* [n] */
    float *transmittance; /* [n][3] */
    float *pdf_success;   /* [n] */
    float *pdf_failure;   /* [n] */
    float *sigma_s;       /* [n][3] (valid on success) */
    int32_t *nsteps;      /* [n] er_step calls */
} mer_medium_sampling_records;

/* Medium::sampleDistance (include/mitsuba/render/medium.h:130-131;
 * heterogeneousrefractive.cpp:402-568).  The Sampler is replayed: xi[n][2] holds the two
 * next1D() draws sampleDistance may consume (distance, balance channel).  Homogeneous
 * sigma only (the reference's behaviour, R2); any record pointer may be NULL. */
int mer_medium_sample_distance_batch(const mer_medium *medium, size_t n, const float *ray_o,
                                     const float *ray_d, const float *ray_mint, const float *xi,
                                     mer_medium_sampling_records *rec);
/* Medium::evalTransmittance (heterogeneousrefractive.cpp:393-400) */
int mer_medium_eval_transmittance_batch(const mer_medium *medium, size_t n, const float *mint,
                                        const float *maxt, float *transmittance_out);

/* ------------------------------------------------ curved direct connections
 * Function-level building blocks of HeterogeneousRefractiveMedium::eval / makeDirectConnections
 * (heterogeneousrefractive.cpp:571-640, 1087-1163).  The solver itself is Ceres BFGS in the reference
 * (not reproducible bit-wise: parity unpinned there); what is deterministic is exposed and parity-tested: */

/* SplineDataSource::valueGradientAndHessian (splinevolume.cpp:371-377, basisspline.h:539-606);
 * hess_out is row-major [n][9] */
int mer_rif_eval_hessian_batch(const mer_rif *rif, size_t n, const float *p, float *value_out, float *grad_out,
                               float *hess_out);
/* nsteps[i] calls of er_derivativestep (:798-814) from dp/dv0 = 0, dv/dv0 = I; p, v in/out; Jacobians [n][9] out */
int mer_medium_derivative_trace_batch(const mer_medium *medium, size_t n, float *p, float *v, const int32_t *nsteps,
                                      float *dpdv0_out, float *dvdv0_out);
/* computefdfBDPT (:816-939): residual p(t*) - p2 at the closest approach of the ray launched from p1 with
 * velocity v0, and its Jacobian (stored transposed like :936-938).  status: 0 inside, 1 left the object
 * (boundary normal from the sdf child, else the analytic box / sphere normal; Snell to exterior index 1 when the medium's
 * boundary is MER_BOUNDARY_HDIELECTRIC as in the reference, unrefracted when index-matched; straight extension), 2 degenerate, 3 left the object by total internal reflection. */
int mer_medium_connection_residual_batch(const mer_medium *medium, int boundary_precision, size_t n, const float *p1,
                                         const float *p2, const float *v0, int is_sensor_sample, float *error_out,
                                         float *derror_out, int32_t *status_out, int32_t *nsteps_out);

/* `tol2`, `rrweight`, `boundaryprecision`, `ceresmaxiterations` (heterogeneousrefractive.cpp:208-219) */
typedef struct mer_connection_params {
    float tol2;                 /* 1e-6: accept when 0.5 |r|^2 < tol2; reject when |p(t*) - p2|^2 > tol2 after the re-trace */
    float rrweight;             /* 1e-2: Russian roulette on solver failures */
    int32_t boundary_precision; /* 3: ceil(precision / log10 2) step halvings */
    int32_t max_iterations;     /* 20 */
    /* first guess of the launch direction: MER_START_RANDOM = every attempt starts uniformly in the hemisphere about the
     * seed direction (the reference, :1078-1105); MER_START_STRAIGHT = the first attempt starts from the seed direction
     * itself and only retries are random.  MER_START_DEFAULT: random for mer_medium_connect_batch (= eval()), straight for
     * the integrator's direct connections. */
    int32_t start_mode;
} mer_connection_params;

enum mer_start_mode { MER_START_DEFAULT = 0, MER_START_RANDOM = 1, MER_START_STRAIGHT = 2 };

/* what HeterogeneousRefractiveMedium::eval fills into the MediumSamplingRecord for a connection (:571-640) */
typedef struct mer_connection_records {
    uint8_t *success;      /* [n] */
    float *dir_to_p2;      /* [n][3] launch velocity at p1 (n-scaled) */
    float *rev_dir_to_p1;  /* [n][3] mRec.drev: unit arrival direction at p2, reversed */
    float *optical_length; /* [n] midpoint-rule optical length (:941-1030) */
    float *distance;       /* [n] curved geometric length */
    float *weight;         /* [n] 1 / rrweight^k */
    float *transmittance;  /* [n][3] exp(-sigma_t * distance) * weight */
    float *pdf_success;    /* [n] */
    float *pdf_failure;    /* [n] */
    int32_t *evaluations;  /* [n] residual evaluations (each one a Jacobian-carrying trace) */
} mer_connection_records;

/* HeterogeneousRefractiveMedium::eval / makeDirectConnections (:571-640, 1087-1163): for every pair find the launch
 * direction at p1 whose eikonal ray passes through p2 (seed directions: uniform on the hemisphere about `seed_dir`,
 * Sampler of pair i = Philox stream (seed, i)), then measure the path (computePathLengthsTillClosestP2 :941-1030) and
 * fill pdfs / transmittance.  The minimiser is Levenberg-Marquardt on computefdfBDPT's residual and Jacobian; the
 * reference uses Ceres 1.14 BFGS (not available, not bit-reproducible: parity unpinned at the solver). */
int mer_medium_connect_batch(const mer_medium *medium, const mer_connection_params *params, size_t n, const float *p1,
                             const float *p2, const float *seed_dir, int is_sensor_sample, uint64_t seed,
                             mer_connection_records *rec);

/* -------------------------------------------------------------- integrator */

enum mer_filter { MER_FILTER_BOX = 0, MER_FILTER_GAUSSIAN = 1 };

/* What the integrator shim extracts from the Scene (sensor, film, rfilter, emitters,
 * MonteCarloIntegrator properties: src/librender/integrator.cpp:190-225). */
typedef struct mer_render_desc {
    int32_t width, height;     /* <film> width/height */
    int32_t spp_total;         /* sampleCount of the whole render (RNG indexing) */
    int32_t sample_begin;      /* this call renders samples s = begin, begin+stride, ... < spp_total */
    int32_t sample_stride;     /* (sample-index sharding across GPUs) */
    uint64_t seed;
    float cam_origin[3], cam_target[3], cam_up[3]; /* <lookat> of the perspective sensor */
    float fov_deg;             /* `fov`, fovAxis = x (src/sensors/perspective.cpp:126-157) */
    int32_t filter;            /* mer_filter; box radius .5 (rfilters/box.cpp), gaussian stddev .5 */
    int32_t max_depth;         /* `maxDepth` (-1 = unbounded) */
    int32_t rr_depth;          /* `rrDepth` (5) */
    float env_radiance[3];     /* constant environment emitter */
    int32_t has_quad;          /* optional two-sided quad area emitter */
    float quad_origin[3], quad_u[3], quad_v[3], quad_radiance[3];
    int32_t pool_paths;        /* resident path slots (0 => default) */
    int32_t steps_per_pass;    /* er_steps per path per wavefront pass (0 => default) */
    /* next-event estimation along the curved path: 0 = the quad emitter is found only by hitting it;
     * 1 = every scattering vertex in the medium is connected to a uniformly sampled point of the quad by solving the
     * shooting problem of makeDirectConnections (heterogeneousrefractive.cpp:1087-1163), and paths that reach the quad
     * after a scattering event no longer count it.  Requires a quad and the tricubic RIF mode; through a density grid the
     * connection's transmittance is exp(-optical depth), midpoint rule on the re-trace's steps.
     * 2 = the same with multiple importance sampling as volpath does it (src/integrators/path/volpath.cpp:120-147,
     * 164-173, miWeight :430-433): the connection is weighted by p_nee^2 / (p_nee^2 + p_phase^2), and a phase-sampled
     * path that reaches the quad counts it with p_phase^2 / (p_phase^2 + p_nee^2), p_nee being the solid-angle density
     * of the curved connection (from the solver's Jacobian, solved for the point that was hit). */
    int32_t direct_connections;
    mer_connection_params connection; /* solver parameters for direct_connections = 1 (zeros => tol2 1e-6, rrweight 1e-2, 3, 20) */
    /* transient film: <film> properties decomposition="transient", minBound, maxBound, binWidth (src/librender/film.cpp
     * :56-78).  frames = ceil((maxBound - minBound) / binWidth) > 1 resolves every contribution by the optical path
     * length of its path (sum of the edges' opticalLength, exterior segments at index 1; bdpt_proc.cpp:147-176) into bin
     * floor((length - min_bound) / bin_width) and drops it when the bin is outside [0, frames) (:446-449).  The film then
     * has 3 * frames + 2 channels per pixel: RGB of every frame, then alpha and weight (bdpt_wr.cpp:52-56).
     * frames <= 1: steady state, 5 channels.  calibrated_transient skips the camera -> first surface segment (:163-171). */
    int32_t frames;
    float min_bound, bin_width;
    int32_t calibrated_transient;
    /* light tracing: 0 = camera path tracer.  1 = width*height*spp_total paths start at the emitter, walk
     * the medium with importance-mode weights (no refRatioSq, BSDF factor 1: edge.cpp:88-98, hdielectric.cpp:262-268) and
     * every scattering vertex is connected to the pinhole through the curved connection of makeDirectConnections with
     * isSensorSample = true (the t = 1 strategy, bdpt_proc.cpp:340-363); the arrival direction picks the pixel
     * (vertex.cpp:1339-1343), the perspective sensor's importance 1 / (A cos^3) weighs it.  Light that reaches the camera
     * without scattering in the medium is not sampled by this strategy.  The environment emitter is ignored. */
    int32_t light_tracing;
    int32_t emitter_type;   /* MER_EMITTER_QUAD: the quad above, cosine-weighted, both sides; MER_EMITTER_COLLIMATED */
    float beam_origin[3], beam_direction[3], beam_power[3]; /* <emitter type="collimated">: src/emitters/collimated.cpp:59-110 */
    /* continuous-wave time-of-flight camera: <film> properties modulation, lambda, phase (degrees) of the fork's
     * PathLengthSampler (src/librender/pathlengthsampler.cpp:6-35).  With a modulation the film has ONE frame
     * (film.cpp:76-78) and every contribution is multiplied by correlationFunction(path length) (:66-96, applied at
     * bdpt_proc.cpp:440-441); frames / min_bound / bin_width are then ignored.  mseq / depthselective codes are not carried. */
    int32_t modulation;
    float lambda, phase_deg;
} mer_render_desc;

enum mer_modulation { MER_MODULATION_NONE = 0, MER_MODULATION_SINE = 1, MER_MODULATION_SQUARE = 2, MER_MODULATION_HAMILTONIAN = 3 };

enum mer_emitter_type { MER_EMITTER_QUAD = 0, MER_EMITTER_COLLIMATED = 1 };

typedef struct mer_render_stats {
    uint64_t samples;        /* camera samples started */
    uint64_t ray_steps;      /* er_step calls (incl. remainder + step-back) */
    uint64_t scatter_events; /* real collisions */
    uint64_t null_collisions;/* Woodcock null collisions */
    uint64_t boundary_exits;
    uint64_t nonfinite_dropped; /* samples rejected like ImageBlock::put (imageblock.h:147-152) */
    uint64_t passes;         /* wavefront passes */
    uint64_t connections;          /* direct connections attempted (direct_connections = 1) */
    uint64_t connections_failed;   /* of those: no solution found / no transmitted path */
    uint64_t connection_steps;     /* Hessian-carrying leapfrog steps spent in the solver */
    uint64_t kernel_launches;
    float device_ms;         /* CUDA-event time of the render kernels */
    float step_kernel_ms;    /* of which: the step kernel (the dominant kernel), summed over its launches */
    uint64_t step_launches;  /* launches of the step kernel */
    float tail_ms;           /* of device_ms: the drain tail, from the first look at the pool that found fewer than half of
                                the path slots alive (the frame has run out of new samples) to the end */
    uint64_t block_fetches;  /* 4x4x4 coefficient blocks (256 bytes; 8 packed-trilinear nodes = 128 bytes) gathered from the grid:
                                the algorithmic memory traffic of the stepper (one per cell change, not per step) */
    uint32_t step_lanes_per_sm; /* resident threads per SM the step kernel ran with for most of the frame: 512, or 192 when the
                                   coefficient table exceeds the L2 and the in-run comparison of the two found 192 faster */
    uint32_t reserved0;
} mer_render_stats;

/* Integrator::render (include/mitsuba/render/integrator.h:61-96) for the eikonal volumetric
 * path tracer: SamplingIntegrator::renderBlock (src/librender/integrator.cpp:140-190) +
 * the volpath bounce loop (src/integrators/path/volpath.cpp:84-343) with the curved-ray walk
 * semantics of libbidir (vertex.cpp:247-279, edge.cpp:26-103) + ImageBlock::put
 * (include/mitsuba/render/imageblock.h:124-190).
 * film: width*height*5 float32 [R,G,B,alpha,weight] (Bitmap::ESpectrumAlphaWeight). */
int mer_render(const mer_medium *medium, const mer_render_desc *desc, float *film_host,
               mer_render_stats *stats_out);
/* film_dev is ACCUMULATED into (zero it first); asynchronous except for the pass-count readbacks */
int mer_render_device(const mer_medium *medium, const mer_render_desc *desc, float *film_dev,
                      mer_render_stats *stats_out, void *stream);
/* The same render on ALL the GPUs of the box in one call: the replacement of the scheduler behind Integrator::render
 * (src/librender/integrator.cpp:95-127: BlockedRenderProcess + sched->schedule/wait; src/librender/renderproc.cpp:142-148:
 * worker results added into the film).  media[g] is the scene's medium created on GPU g (grids replicated per GPU);
 * GPU g renders the sample indices s = begin + (g + k * ngpus) * stride of every pixel, one host thread per GPU, and the
 * per-GPU films are added on media[0]'s GPU with one ncclReduce over NVLink (libnccl.so.2, loaded at run time; peer copies
 * + an add kernel when it is missing or when two handles share a device), then copied to film_host.
 * stats: counters summed over the GPUs, times = the slowest GPU's. */
int mer_render_multi(const mer_medium *const *media, int32_t ngpus, const mer_render_desc *desc, float *film_host,
                     mer_render_stats *stats_out);
/* HDRFilm::develop (src/films/hdrfilm.cpp:527-540): rgb = sum(w*RGB)/sum(w); host buffers */
int mer_film_develop(int device, int32_t width, int32_t height, const float *film, float *rgb_out);
/* the same for a transient film of `frames` frames: film [H][W][3*frames+2] -> rgb_out [H][W][frames][3] */
int mer_film_develop_frames(int device, int32_t width, int32_t height, int32_t frames, const float *film, float *rgb_out);

/* ------------------------------------------------------------------- misc */

const char *mer_last_error(void);
int mer_abi_version(void);
/* number of usable sm_100-class devices (0 => every compute call fails) */
int mer_device_count(void);
/* total kernels launched by this library in this process (bench.py `gpu_launches`) */
uint64_t mer_kernel_launch_count(void);

/* Hands the memory this library caches on `device` back to the driver: the private stream-ordered pool that backs the
 * volume handles, the cached coefficient atlases and the render scratch (path pool).  The reference has no equivalent
 * (its volumes are mmap()ed files, src/volume/splinevolume.cpp:204-317); a host that shares the GPU calls this after
 * tearing a scene down.  Live handles are not affected. */
int mer_trim_memory(int device);

#ifdef __cplusplus
}
#endif
#endif /* MITSUBAER_B200_H */
