/*
 * oracle/ref_spline.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference's header-only spline
 *     /root/reference/include/mitsuba/core/basisspline.h
 * VERBATIM (from where it lies; nothing is copied into this repo) behind a C ABI, using
 * the shim headers in oracle/shim/.  Built twice by oracle/Makefile into
 *     oracle/_ref/libmer_refspline_f.so   (FLOAT = float,  Mitsuba's default Float)
 *     oracle/_ref/libmer_refspline_d.so   (FLOAT = double, the -DFLOATDEBUG build the authors used, R9)
 * It pins rows a1-a4 of SURVEY.md §8(a): the restated spline in oracle/mer_oracle.cpp and
 * the golden vectors in tests/golden/ are checked against this library.
 */
#include <iostream>
#include <math.h>
#include <mitsuba/core/platform.h>
/* `coeff` is a private member; expose it for the prefilter parity check without touching
 * the header (system headers are already included above, so only basisspline.h sees this). */
#define private public
#include <mitsuba/core/basisspline.h>
#undef private

using mitsuba::FLOAT;
typedef mitsuba::basisspline::Spline<3> Spline3;

extern "C" {

void *ref_spline_build(const float *data, const int *N, const float *bmin, const float *bmax) {
    FLOAT xmin[3], xmax[3];
    int n[3];
    for (int i = 0; i < 3; i++) {
        xmin[i] = (FLOAT) bmin[i];
        xmax[i] = (FLOAT) bmax[i];
        n[i] = N[i];
    }
    Spline3 *s = new Spline3();
    s->initialize(xmin, xmax, n);
    size_t total = (size_t) N[0] * N[1] * N[2];
    FLOAT *tmp = new FLOAT[total];
    for (size_t i = 0; i < total; i++)
        tmp[i] = (FLOAT) data[i]; /* float -> FLOAT copy as splinevolume.cpp:284-287 */
    s->build(tmp);
    delete[] tmp;
    return s;
}

void ref_spline_free(void *h) { delete (Spline3 *) h; }

int ref_spline_sizeof_float(void) { return (int) sizeof(FLOAT); }

/* what: 0 value, 1 gradient, 2 valueAndGradient.  p: [n][3] FLOAT. */
void ref_spline_eval(void *h, int what, size_t n, const FLOAT *p, FLOAT *f, FLOAT *g) {
    const Spline3 *s = (const Spline3 *) h;
    for (size_t i = 0; i < n; i++) {
        FLOAT x[3] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
        if (what == 0) {
            f[i] = s->value(x);
        } else if (what == 1) {
            mitsuba::VectorF v = s->gradient(x);
            g[3 * i] = v.x; g[3 * i + 1] = v.y; g[3 * i + 2] = v.z;
        } else {
            FLOAT fv;
            mitsuba::VectorF v;
            s->valueAndGradient(x, fv, v);
            f[i] = fv;
            g[3 * i] = v.x; g[3 * i + 1] = v.y; g[3 * i + 2] = v.z;
        }
    }
}

/* value, gradient and Hessian (row-major 3x3) — basisspline.h:539-606 */
void ref_spline_eval_hessian(void *h, size_t n, const FLOAT *p, FLOAT *f, FLOAT *g, FLOAT *H) {
    const Spline3 *s = (const Spline3 *) h;
    for (size_t i = 0; i < n; i++) {
        FLOAT x[3] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
        FLOAT fv;
        mitsuba::VectorF v;
        mitsuba::Matrix3x3F M;
        s->valueGradientAndHessian(x, fv, v, M);
        f[i] = fv;
        g[3 * i] = v.x; g[3 * i + 1] = v.y; g[3 * i + 2] = v.z;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                H[9 * i + 3 * a + b] = M.m[a][b];
    }
}

/* prefiltered coefficients (basisspline.h:865-890), same layout as the input */
void ref_spline_coeffs(void *h, FLOAT *out) {
    const Spline3 *s = (const Spline3 *) h;
    size_t total = (size_t) s->N[0] * s->N[1] * s->N[2];
    for (size_t i = 0; i < total; i++)
        out[i] = s->coeff[i];
}

FLOAT ref_spline_stride(void *h, int dim) { return ((const Spline3 *) h)->getStride(dim); }

} /* extern "C" */
