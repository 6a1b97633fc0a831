/*
 * uivr_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See uivr_oracle.h.
 *
 * One scalar sample at a time, written to read like python/integrators/volpathsimple.py.
 * Only + - * / sqrt fma and integer ops are used (compile with -ffp-contract=off), so the
 * CUDA path can reproduce every branch decision bit for bit (DESIGN.md "Arithmetic contract").
 */
#include "uivr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FMA(a, b, c) __builtin_fmaf((a), (b), (c))

/* ------------------------------------------------------------------------------------ */
/* RNG: TEA + PCG32 `independent` sampler  [UPSTREAM, SURVEY App. B.1-B.2]               */
/* ------------------------------------------------------------------------------------ */

void uivr_oracle_tea(uint32_t v0, uint32_t v1, uint32_t out[2]) {
    uint32_t sum = 0;
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    out[0] = v0;
    out[1] = v1;
}

typedef struct {
    uint64_t state, inc;
    uint64_t draws;
} rng_t;

static inline uint32_t pcg_next(rng_t* r) {
    uint64_t old = r->state;
    r->state = old * 0x5851f42d4c957f2dull + r->inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    r->draws++;
    return (xs >> rot) | (xs << ((0u - rot) & 31u));
}

static inline void pcg_seed(rng_t* r, uint64_t initstate, uint64_t initseq) {
    r->state = 0;
    r->inc = (initseq << 1u) | 1u;
    pcg_next(r);
    r->state += initstate;
    pcg_next(r);
    r->draws = 0;
}

/* sampler.seed(seed, wavefront): per-lane stream = PCG32(TEA(seed, idx)) */
static inline void sampler_seed(rng_t* r, uint32_t seed, uint32_t idx) {
    uint32_t v[2];
    uivr_oracle_tea(seed, idx, v);
    pcg_seed(r, v[0], v[1]);
}

static inline float u32_to_float(uint32_t x) {
    union { uint32_t u; float f; } c;
    c.u = (x >> 9) | 0x3f800000u;
    return c.f - 1.0f;
}

static inline float rng_f(rng_t* r) { return u32_to_float(pcg_next(r)); }

void uivr_oracle_pcg32_stream(uint64_t initstate, uint64_t initseq, int n, uint32_t* out) {
    rng_t r;
    pcg_seed(&r, initstate, initseq);
    for (int i = 0; i < n; ++i) out[i] = pcg_next(&r);
}

void uivr_oracle_sampler_floats(uint32_t seed, uint32_t idx, int n, float* out) {
    rng_t r;
    sampler_seed(&r, seed, idx);
    for (int i = 0; i < n; ++i) out[i] = rng_f(&r);
}

/* volpathsimple.py:99-107: the alt sampler's seed is derived from the bit pattern of lane
 * 0's `alt_seed_rnd`: the float of stream (seed_grad, idx 0) that follows the draws made
 * before :99 -- for mi.render 2 jitter draws (sample_rays) + 1 burned draw (:71), so the 4th
 * float; for render_batch the path sampler draws no jitter (batched.py:390, :437), so the 2nd.
 * A masked-out lane still yields the value it would have drawn, so this is a pure function of
 * seed_grad. */
static uint32_t alt_seed_after(uint32_t seed_grad, int skipped) {
    rng_t r;
    sampler_seed(&r, seed_grad, 0);
    for (int i = 0; i < skipped; ++i) rng_f(&r);
    union { float f; uint32_t u; } c;
    c.f = rng_f(&r);
    uint32_t v[2];
    uivr_oracle_tea(c.u, 1, v);
    return v[0];
}

uint32_t uivr_oracle_alt_seed(uint32_t seed_grad) { return alt_seed_after(seed_grad, 3); }
uint32_t uivr_oracle_alt_seed_batch(uint32_t seed_grad) { return alt_seed_after(seed_grad, 1); }

/* ------------------------------------------------------------------------------------ */
/* Exact-op transcendental replacements (DESIGN.md "Arithmetic contract")                */
/* ------------------------------------------------------------------------------------ */

/* -ln(1-u), u in [0,1).  x = 1-u is exact; x = m*2^e, m in [sqrt(.5), sqrt(2));
 * ln(m) = f - f^2/2 + f^3 q(f), f = m-1, q = degree-7 fit (rel. err 7e-9). */
static inline float neg_log1m(float u) {
    float x = 1.0f - u;
    union { float f; uint32_t u; } c;
    c.f = x;
    uint32_t ix = c.u + (0x3f800000u - 0x3f3504f3u);
    int e = (int) (ix >> 23) - 127;
    c.u = (ix & 0x007fffffu) + 0x3f3504f3u;
    float f = c.f - 1.0f;
    float q = -0x1.2d9544p-4f;
    q = FMA(q, f, 0x1.0276dcp-3f);
    q = FMA(q, f, -0x1.0e610cp-3f);
    q = FMA(q, f, 0x1.235f0ep-3f);
    q = FMA(q, f, -0x1.5467dap-3f);
    q = FMA(q, f, 0x1.9998b0p-3f);
    q = FMA(q, f, -0x1.00023cp-2f);
    q = FMA(q, f, 0x1.555564p-2f);
    float f2 = f * f;
    float lm = FMA(f2 * f, q, FMA(-0.5f, f2, f));
    float l = FMA((float) e, 0x1.62e430p-1f, lm);
    return -l;
}

/* sin(2 pi x), cos(2 pi x), x in [0,1): quadrant k = floor(4x+.5), r = 4x-k in [-.5,.5],
 * sin(pi/2 r) = r S(r^2), cos(pi/2 r) = C(r^2). */
static inline void sincos2pi(float x, float* s, float* c) {
    float y = 4.0f * x;
    int k = (int) (y + 0.5f);
    float r = y - (float) k;
    float z = r * r;
    float ps = -0x1.2d9b78p-8f;
    ps = FMA(ps, z, 0x1.465ec4p-4f);
    ps = FMA(ps, z, -0x1.4abbbap-1f);
    ps = FMA(ps, z, 0x1.921fb6p+0f);
    ps = ps * r;
    float pc = 0x1.d9c322p-11f;
    pc = FMA(pc, z, -0x1.55c57ap-6f);
    pc = FMA(pc, z, 0x1.03c1dcp-2f);
    pc = FMA(pc, z, -0x1.3bd3ccp+0f);
    pc = FMA(pc, z, 1.0f);
    switch (k & 3) {
        case 0: *s = ps;  *c = pc;  break;
        case 1: *s = pc;  *c = -ps; break;
        case 2: *s = -ps; *c = -pc; break;
        default: *s = -pc; *c = ps; break;
    }
}

void uivr_oracle_neg_log1m(const float* u, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = neg_log1m(u[i]);
}

void uivr_oracle_sincos2pi(const float* x, int n, float* s, float* c) {
    for (int i = 0; i < n; ++i) sincos2pi(x[i], &s[i], &c[i]);
}

/* exp(x): n = rint(x log2 e), r = x - n ln2 (two-part), degree-5 polynomial on r (Cephes expf
 * coefficients), scaled by 2^n through the exponent bits.  x clamped to [-87, 88]. */
static inline float exp_exact(float x) {
    x = x < -87.0f ? -87.0f : (x > 88.0f ? 88.0f : x);
    float n = rintf(x * 0x1.715476p+0f);
    float r = FMA(n, -0x1.62e400p-1f, x);
    r = FMA(n, -0x1.7f7d1cp-20f, r);
    float q = 1.9875691500e-4f;
    q = FMA(q, r, 1.3981999507e-3f);
    q = FMA(q, r, 8.3334519073e-3f);
    q = FMA(q, r, 4.1665795894e-2f);
    q = FMA(q, r, 1.6666665459e-1f);
    q = FMA(q, r, 5.0000001201e-1f);
    float e = FMA(q, r * r, r) + 1.0f;
    union { uint32_t u; float f; } sc;
    sc.u = (uint32_t) ((int) n + 127) << 23;
    return e * sc.f;
}

void uivr_oracle_exp(const float* x, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = exp_exact(x[i]);
}

/* warp::square_to_uniform_sphere [UPSTREAM, SURVEY App. B.7] */
static inline void uniform_sphere(float xi1, float xi2, float w[3]) {
    float z = FMA(-2.0f, xi2, 1.0f);
    float r2 = FMA(-z, z, 1.0f);
    float r = sqrtf(r2 > 0.0f ? r2 : 0.0f);
    float s, c;
    sincos2pi(xi1, &s, &c);
    w[0] = r * c;
    w[1] = r * s;
    w[2] = z;
}

/* ------------------------------------------------------------------------------------ */
/* Context                                                                               */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    const uivr_oracle_scene* sc;
    const uivr_oracle_batch* batch; /* ray-batch mode (batched.py), else NULL */
    const uivr_oracle_nerf* nerf;   /* emission-absorption ray marching (nerf.py), else NULL; `albedo` = emission grid */
    const float* sigma_t;  /* (Z,Y,X)   */
    const float* albedo;   /* (Z,Y,X,3) */
    int32_t mres[3];
    float mcs[3];          /* supergrid cell size 1/M */
    float* majorant;
    uint8_t* exit_mask;    /* per supergrid cell: bit o = nothing but empty cells ahead in octant o (see below) */
    float half_le[3];      /* 0.5 * radiance: NEE weight phase*mis*Le/pdf folded */
    /* adjoint accumulation (shared between threads, CAS-atomic doubles) */
    double* dsigma;
    double* dalbedo;
} ctx_t;

typedef struct {
    uint64_t c[UIVR_ORC_NUM_COUNTERS];
    uint64_t r[UIVR_ORC_NUM_COUNTERS]; /* the share of c spent re-walking NEE segments for their adjoint (:393-401) */
} counters_t;

/* ------------------------------------------------------------------------------------ */
/* GridVolume trilinear lookup + its adjoint  [UPSTREAM; a16, a17]      */
/* ------------------------------------------------------------------------------------ */

static inline int grid_cell(const float p[3], const int32_t res[3], int i0[3], float w[3]) {
    if (!(p[0] >= 0.0f && p[0] <= 1.0f && p[1] >= 0.0f && p[1] <= 1.0f && p[2] >= 0.0f &&
          p[2] <= 1.0f))
        return 0;
    for (int a = 0; a < 3; ++a) {
        float q = FMA(p[a], (float) res[a], -0.5f);
        float fl = floorf(q);
        i0[a] = (int) fl;
        w[a] = q - fl;
    }
    return 1;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float lerpf(float a, float b, float w) { return FMA(w, b - a, a); }

static void trilinear(const float* grid, const int32_t res[3], int nch, const float p[3],
                      float* out) {
    int i0[3];
    float w[3];
    if (!grid_cell(p, res, i0, w)) {
        for (int c = 0; c < nch; ++c) out[c] = 0.0f;
        return;
    }
    int x0 = clampi(i0[0], 0, res[0] - 1), x1 = clampi(i0[0] + 1, 0, res[0] - 1);
    int y0 = clampi(i0[1], 0, res[1] - 1), y1 = clampi(i0[1] + 1, 0, res[1] - 1);
    int z0 = clampi(i0[2], 0, res[2] - 1), z1 = clampi(i0[2] + 1, 0, res[2] - 1);
#define VOX(z, y, x, c) grid[(((size_t) (z) * res[1] + (y)) * res[0] + (x)) * nch + (c)]
    for (int c = 0; c < nch; ++c) {
        float c00 = lerpf(VOX(z0, y0, x0, c), VOX(z0, y0, x1, c), w[0]);
        float c10 = lerpf(VOX(z0, y1, x0, c), VOX(z0, y1, x1, c), w[0]);
        float c01 = lerpf(VOX(z1, y0, x0, c), VOX(z1, y0, x1, c), w[0]);
        float c11 = lerpf(VOX(z1, y1, x0, c), VOX(z1, y1, x1, c), w[0]);
        float c0 = lerpf(c00, c10, w[1]);
        float c1 = lerpf(c01, c11, w[1]);
        out[c] = lerpf(c0, c1, w[2]);
    }
#undef VOX
}

void uivr_oracle_trilinear(const float* grid, const int32_t res[3], int channels,
                           const float* p, int n, float* out) {
    for (int i = 0; i < n; ++i) trilinear(grid, res, channels, p + 3 * i, out + (size_t) channels * i);
}

static inline void atomic_add_double(double* addr, double v) {
    union { double d; uint64_t u; } old, neu;
    uint64_t* a = (uint64_t*) addr;
    old.u = __atomic_load_n(a, __ATOMIC_RELAXED);
    do {
        neu.d = old.d + v;
    } while (!__atomic_compare_exchange_n(a, &old.u, neu.u, 1, __ATOMIC_RELAXED,
                                          __ATOMIC_RELAXED));
}

/* Adjoint of the lookup: scatter-add g*w_k into the 8 voxels (nch channels). */
static void scatter(double* dgrid, const int32_t res[3], int nch, const float p[3],
                    const float* g) {
    int i0[3];
    float w[3];
    if (!grid_cell(p, res, i0, w)) return;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                int x = clampi(i0[0] + dx, 0, res[0] - 1);
                int y = clampi(i0[1] + dy, 0, res[1] - 1);
                int z = clampi(i0[2] + dz, 0, res[2] - 1);
                float wx = dx ? w[0] : 1.0f - w[0];
                float wy = dy ? w[1] : 1.0f - w[1];
                float wz = dz ? w[2] : 1.0f - w[2];
                float wk = (wx * wy) * wz;
                size_t base = (((size_t) z * res[1] + y) * res[0] + x) * nch;
                for (int c = 0; c < nch; ++c) atomic_add_double(&dgrid[base + c], (double) (g[c] * wk));
            }
}

/* sigma_t(p) = scale * grid(p)   (get_scattering_coefficients, a16) */
static inline float eval_sigma_t(const ctx_t* C, counters_t* K, const float p[3]) {
    float v;
    trilinear(C->sigma_t, C->sc->res, 1, p, &v);
    K->c[UIVR_ORC_SIGMA_TAPS]++;
    return C->sc->scale * v;
}

static inline void eval_albedo(const ctx_t* C, counters_t* K, const float p[3], float a[3]) {
    trilinear(C->albedo, C->sc->res, 3, p, a);
    K->c[UIVR_ORC_ALBEDO_TAPS]++;
}

/* d sigma_t grid += (scale*g) * w_k   (chain rule through sigma_t = scale*grid) */
static inline void scatter_sigma(const ctx_t* C, counters_t* K, const float p[3], float g) {
    float gs = C->sc->scale * g;
    scatter(C->dsigma, C->sc->res, 1, p, &gs);
    K->c[UIVR_ORC_SIGMA_SCATTERS]++;
}

static inline void scatter_albedo(const ctx_t* C, counters_t* K, const float p[3], const float g[3]) {
    scatter(C->dalbedo, C->sc->res, 3, p, g);
    K->c[UIVR_ORC_ALBEDO_SCATTERS]++;
}

/* ------------------------------------------------------------------------------------ */
/* Majorant supergrid  [UPSTREAM, SURVEY App. B.5, a18]                                  */
/* ------------------------------------------------------------------------------------ */

static inline int floordiv(int a, int b) { /* b > 0 */
    int q = a / b;
    if ((a % b) != 0 && a < 0) --q;
    return q;
}

void uivr_oracle_build_majorant(const float* sigma_t, const int32_t res[3], float scale,
                                int32_t factor, int32_t mres[3], float* out) {
    for (int a = 0; a < 3; ++a) {
        mres[a] = (factor > 1) ? res[a] / factor : 1;
        if (mres[a] < 1) mres[a] = 1;
    }
    for (int cz = 0; cz < mres[2]; ++cz)
        for (int cy = 0; cy < mres[1]; ++cy)
            for (int cx = 0; cx < mres[0]; ++cx) {
                int c[3] = {cx, cy, cz}, lo[3], hi[3];
                /* voxels whose trilinear footprint touches local [c/M, (c+1)/M] */
                for (int a = 0; a < 3; ++a) {
                    lo[a] = clampi(floordiv(2 * c[a] * res[a] - mres[a], 2 * mres[a]), 0, res[a] - 1);
                    hi[a] = clampi(floordiv(2 * (c[a] + 1) * res[a] - mres[a], 2 * mres[a]) + 1, 0, res[a] - 1);
                }
                float m = 0.0f;
                for (int z = lo[2]; z <= hi[2]; ++z)
                    for (int y = lo[1]; y <= hi[1]; ++y)
                        for (int x = lo[0]; x <= hi[0]; ++x) {
                            float v = sigma_t[((size_t) z * res[1] + y) * res[0] + x];
                            if (v > m) m = v;
                        }
                out[((size_t) cz * mres[1] + cy) * mres[0] + cx] = scale * m;
            }
}

/* Exit mask of the supergrid [part of OUR definition of Medium::sample_interaction, App. B.5]: bit o of
 * cell c is set iff the majorant is 0 in every cell of the box spanned by c and the grid corner that
 * octant o points to (octant bit a = the ray direction is negative along axis a).  A walk that enters
 * such a cell can only meet empty cells from there on, so it ends without a collision right away;
 * results are those of walking on to the boundary, only the number of supergrid reads changes. */
void uivr_oracle_build_exit_mask(const float* majorant, const int32_t mres[3], uint8_t* out) {
    const int mx = mres[0], my = mres[1], mz = mres[2];
    for (int o = 0; o < 8; ++o) {
        const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
        /* visit the cells so that the three neighbours ahead (c + s_a e_a) are done first */
        for (int kz = 0; kz < mz; ++kz)
            for (int ky = 0; ky < my; ++ky)
                for (int kx = 0; kx < mx; ++kx) {
                    const int x = sx > 0 ? mx - 1 - kx : kx, y = sy > 0 ? my - 1 - ky : ky, z = sz > 0 ? mz - 1 - kz : kz;
                    const size_t c = ((size_t) z * my + y) * mx + x;
                    int e = !(majorant[c] > 0.0f);
                    if (o == 0) out[c] = 0;
                    const int nx = x + sx, ny = y + sy, nz = z + sz;
                    if (e && nx >= 0 && nx < mx) e = (out[((size_t) z * my + y) * mx + nx] >> o) & 1;
                    if (e && ny >= 0 && ny < my) e = (out[((size_t) z * my + ny) * mx + x] >> o) & 1;
                    if (e && nz >= 0 && nz < mz) e = (out[((size_t) nz * my + y) * mx + x] >> o) & 1;
                    if (e) out[c] |= (uint8_t) (1u << o);
                }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Segments, box intersection, free-flight walk                                          */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    float o[3], d[3], inv_d[3]; /* local-space origin / direction, t in world units */
    float tmax;
} seg_t;

#define UIVR_INF (__builtin_inff())
#define ENTRY_EPS 1e-4f

static inline void dir_to_local(const float M[12], const float w[3], float d[3]) {
    for (int a = 0; a < 3; ++a)
        d[a] = FMA(M[4 * a + 0], w[0], FMA(M[4 * a + 1], w[1], M[4 * a + 2] * w[2]));
}

/* Distance from a point inside local [0,1]^3 to the boundary along d (scene.ray_intersect
 * from inside the medium, volpathsimple.py:234, :428, :637).  <= 0 means "no exit found". */
static inline float exit_distance(seg_t* s) {
    float t = UIVR_INF;
    for (int a = 0; a < 3; ++a) {
        if (s->d[a] != 0.0f) {
            s->inv_d[a] = 1.0f / s->d[a];
            float bound = s->d[a] > 0.0f ? 1.0f : 0.0f;
            float ta = (bound - s->o[a]) * s->inv_d[a];
            if (ta < t) t = ta;
        } else {
            s->inv_d[a] = UIVR_INF;
        }
    }
    return t;
}

/* new segment leaving local point p along world direction w; returns 0 on accidental escape */
static inline int make_segment(const ctx_t* C, const float p[3], const float w[3], seg_t* s) {
    for (int a = 0; a < 3; ++a) s->o[a] = p[a];
    dir_to_local(C->sc->to_local, w, s->d);
    s->tmax = exit_distance(s);
    return s->tmax > 0.0f && s->tmax < UIVR_INF;
}

static inline void seg_point(const seg_t* s, float t, float p[3]) {
    for (int a = 0; a < 3; ++a) p[a] = FMA(t, s->d[a], s->o[a]);
}

/* ------------------------------------------------------------------------------------ */
/* envmap emitter [UPSTREAM envmap.cpp as recalled; conventions defined in scene.py EnvMap] */
/* ------------------------------------------------------------------------------------ */

#define INV_4PI 0.07957747154594767f
#define INV_2PI2 0.05066059182116889f      /* 1 / (2 pi^2) */
#define ENV_EPS2 3.5527137e-15f            /* (2^-24)^2 */
#define ONE_MINUS_EPS 0.99999994f

/* atan2(y, x) / (2 pi) in [-0.5, 0.5]: octant reduction + Cephes atanf polynomial, exact ops only */
static inline float atan2_turns(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float hi = ax > ay ? ax : ay, lo = ax > ay ? ay : ax;
    float a = hi > 0.0f ? lo / hi : 0.0f;
    float off = 0.0f, t = a;
    if (a > 0.41421356f) {
        t = (a - 1.0f) / (a + 1.0f);
        off = 0.78539816f;
    }
    float z = t * t;
    float p = 8.05374449538e-2f;
    p = FMA(p, z, -1.38776856032e-1f);
    p = FMA(p, z, 1.99777106478e-1f);
    p = FMA(p, z, -3.33329491539e-1f);
    float r = (FMA(p * z, t, t) + off) * 0.15915494f;
    if (ay > ax) r = 0.25f - r;
    if (x < 0.0f) r = 0.5f - r;
    if (y < 0.0f) r = -r;
    return r;
}

void uivr_oracle_atan2_turns(const float* y, const float* x, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = atan2_turns(y[i], x[i]);
}

static inline void mat3_apply(const float M[9], const float v[3], float o[3]) {
    for (int a = 0; a < 3; ++a) o[a] = FMA(M[3 * a + 0], v[0], FMA(M[3 * a + 1], v[1], M[3 * a + 2] * v[2]));
}

/* eval_spectrum(uv) * scale and the sampling density of the patch containing uv */
static void env_lookup(const uivr_oracle_scene* sc, float tu, float tv, float le[3], float* pdf_uv) {
    const int W = sc->env_w, H = sc->env_h;
    float fx = tu * (float) W, fy = tv * (float) (H - 1);
    int px = (int) fx, py = (int) fy;
    px = px > W - 1 ? W - 1 : (px < 0 ? 0 : px);
    py = py > H - 2 ? H - 2 : (py < 0 ? 0 : py);
    float w1x = fx - (float) px, w1y = fy - (float) py, w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    const float* v00 = sc->env_data + 4 * ((size_t) py * (W + 1) + px);
    const float* v10 = v00 + 4;
    const float* v01 = v00 + 4 * (size_t) (W + 1);
    const float* v11 = v01 + 4;
    for (int c = 0; c < 3; ++c) {
        float s0 = FMA(w0x, v00[c], w1x * v10[c]);
        float s1 = FMA(w0x, v01[c], w1x * v11[c]);
        le[c] = FMA(w0y, s0, w1y * s1) * sc->env_scale;
    }
    *pdf_uv = v00[3];
}

static inline float env_inv_sin_theta(const float d[3]) {
    float s2 = FMA(d[0], d[0], d[2] * d[2]);
    return 1.0f / sqrtf(s2 > ENV_EPS2 ? s2 : ENV_EPS2);
}

/* Emitter::eval(si) and Emitter::pdf_direction for a ray leaving along `d_local` (segment direction) */
static void env_eval(const uivr_oracle_scene* sc, const float d_local[3], float le[3], float* pdf_dir) {
    float w[3], d[3];
    mat3_apply(sc->local_to_world, d_local, w);
    mat3_apply(sc->world_to_env, w, d);
    float u = atan2_turns(d[0], -d[2]);
    float cy = d[1];
    float sy = sqrtf(FMA(-cy, cy, 1.0f) > 0.0f ? FMA(-cy, cy, 1.0f) : 0.0f);
    float tv = 2.0f * atan2_turns(sy, cy);
    float tu = u - 0.5f / (float) sc->env_w;
    tu -= floorf(tu);
    tv = tv < 0.0f ? 0.0f : (tv > 1.0f ? 1.0f : tv);
    float pdf_uv;
    env_lookup(sc, tu, tv, le, &pdf_uv);
    *pdf_dir = (pdf_uv * env_inv_sin_theta(d)) * INV_2PI2;
}

/* upper bound in a CDF table: smallest i in [0, n) with cdf[i] > x, clamped to n - 1; returns the
 * position of x inside that bin in *frac */
static inline int cdf_find(const float* cdf, int n, float x, float* frac) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] > x) hi = mid; else lo = mid + 1;
    }
    float a = lo > 0 ? cdf[lo - 1] : 0.0f, den = cdf[lo] - a;
    float f = den > 0.0f ? (x - a) / den : 0.5f;
    *frac = f < 0.0f ? 0.0f : (f > ONE_MINUS_EPS ? ONE_MINUS_EPS : f);
    return lo;
}

/* Scene::sample_emitter_direction: world direction, solid-angle pdf, radiance */
static void env_sample(const uivr_oracle_scene* sc, float xi1, float xi2, float w[3], float* pdf_dir, float le[3]) {
    const int W = sc->env_w, H = sc->env_h;
    float f1, f2;
    int r = cdf_find(sc->env_marg, H - 1, xi2, &f2);
    int c = cdf_find(sc->env_cond + (size_t) r * W, W, xi1, &f1);
    float tu = ((float) c + f1) / (float) W, tv = ((float) r + f2) / (float) (H - 1);
    float pdf_uv;
    env_lookup(sc, tu, tv, le, &pdf_uv);
    pdf_uv = sc->env_data[4 * ((size_t) r * (W + 1) + c) + 3];
    float du = tu + 0.5f / (float) W;
    du -= floorf(du);
    float sp, cp, st, ct, d[3];
    sincos2pi(du, &sp, &cp);
    sincos2pi(0.5f * tv, &st, &ct);
    d[0] = sp * st;
    d[1] = ct;
    d[2] = -(cp * st);
    *pdf_dir = (pdf_uv * env_inv_sin_theta(d)) * INV_2PI2;
    mat3_apply(sc->env_to_world, d, w);
}

/* mi.ad.common.mis_weight: power heuristic */
static inline float mis_power(float a, float b) {
    float a2 = a * a;
    return a > 0.0f ? a2 / FMA(b, b, a2) : 0.0f;
}

/* analysis hooks (scripts/dda_stats.c counts supergrid steps by kind); no-ops in the oracle build */
#ifndef UIVR_ORACLE_WALK_HOOK
#define UIVR_ORACLE_WALK_HOOK(C, s, w)
#define UIVR_ORACLE_STEP_HOOK(C, w)
#endif

/* Medium::sample_interaction with supergrid DDA [UPSTREAM App. B.5, a14]; the DDA state is
 * carried along the segment instead of being restarted per call (same distribution). */
typedef struct {
    float t, tmax;
    int cell[3], step[3];
    float tn[3], dt[3];
    float sig_bar;
    int octant;            /* bit a: d[a] < 0 */
} walk_t;

static inline float majorant_at(const ctx_t* C, counters_t* K, const int cell[3]) {
    K->c[UIVR_ORC_MAJORANT_READS]++;
    return C->majorant[((size_t) cell[2] * C->mres[1] + cell[1]) * C->mres[0] + cell[0]];
}

/* test hook: with the exit mask switched off every walk runs on to the boundary; radiance and gradients
 * must not change by a single bit (tests/test_oracle.py), only the number of supergrid reads does */
static int g_use_exit_mask = 1;
void uivr_oracle_set_exit_mask(int enable) { g_use_exit_mask = enable; }

/* the walk has entered a cell behind which (along the ray's octant) every cell is empty */
static inline int walk_sees_exit(const ctx_t* C, const walk_t* w) {
    if (!g_use_exit_mask) return 0;
    const size_t c = ((size_t) w->cell[2] * C->mres[1] + w->cell[1]) * C->mres[0] + w->cell[0];
    return !(w->sig_bar > 0.0f) && ((C->exit_mask[c] >> w->octant) & 1);
}

static void walk_init(const ctx_t* C, counters_t* K, const seg_t* s, walk_t* w) {
    w->t = 0.0f;
    w->tmax = s->tmax;
    w->octant = (s->d[0] < 0.0f ? 1 : 0) | (s->d[1] < 0.0f ? 2 : 0) | (s->d[2] < 0.0f ? 4 : 0);
    for (int a = 0; a < 3; ++a) {
        int c = (int) floorf(s->o[a] * (float) C->mres[a]);
        c = clampi(c, 0, C->mres[a] - 1);
        w->cell[a] = c;
        if (s->d[a] > 0.0f) {
            w->tn[a] = ((float) (c + 1) * C->mcs[a] - s->o[a]) * s->inv_d[a];
            w->dt[a] = C->mcs[a] * s->inv_d[a];
            w->step[a] = 1;
        } else if (s->d[a] < 0.0f) {
            w->tn[a] = ((float) c * C->mcs[a] - s->o[a]) * s->inv_d[a];
            w->dt[a] = -(C->mcs[a] * s->inv_d[a]);
            w->step[a] = -1;
        } else {
            w->tn[a] = UIVR_INF;
            w->dt[a] = 0.0f;
            w->step[a] = 0;
        }
    }
    w->sig_bar = majorant_at(C, K, w->cell);
    UIVR_ORACLE_WALK_HOOK(C, s, w);
}

/* next tentative collision for uniform u; returns 0 when the segment end is reached */
static int walk_next(const ctx_t* C, counters_t* K, walk_t* w, float u, float* t_out,
                     float* sig_bar_out) {
    float tau = neg_log1m(u);
    for (;;) {
        UIVR_ORACLE_STEP_HOOK(C, w);
        if (walk_sees_exit(C, w)) return 0;
        int ax = 0;
        if (w->tn[1] < w->tn[ax]) ax = 1;
        if (w->tn[2] < w->tn[ax]) ax = 2;
        float t_end = w->tn[ax] < w->tmax ? w->tn[ax] : w->tmax;
        float len = t_end - w->t;
        if (len < 0.0f) len = 0.0f;
        if (w->sig_bar > 0.0f) {
            float dtau = w->sig_bar * len;
            if (tau < dtau) {
                float t = w->t + tau / w->sig_bar;
                if (t > t_end) t = t_end;
                w->t = t;
                *t_out = t;
                *sig_bar_out = w->sig_bar;
                return 1;
            }
            tau -= dtau;
        }
        if (t_end > w->t) w->t = t_end;
        if (!(w->tn[ax] < w->tmax)) return 0;
        w->cell[ax] += w->step[ax];
        if (w->cell[ax] < 0 || w->cell[ax] >= C->mres[ax]) return 0;
        w->tn[ax] += w->dt[ax];
        w->sig_bar = majorant_at(C, K, w->cell);
    }
}

/* ------------------------------------------------------------------------------------ */
/* estimate_transmittance: ratio tracking  (volpathsimple.py:436-504, a7)                */
/* adjoint != NULL: per tentative collision with tr > 0, d sigma_t(p) += -sum(adj)/sigma_n */
/* ------------------------------------------------------------------------------------ */

static float ratio_track(const ctx_t* C, counters_t* K, const seg_t* s, rng_t* rng,
                         const float* adjoint, int* n_null) {
    walk_t w;
    walk_init(C, K, s, &w);
    float T = 1.0f;
    int nn = 0; /* tentative collisions with sigma_n > 0: the ones the adjoint scatters at */
    for (;;) {
        float t, sb, p[3];
        if (!walk_next(C, K, &w, rng_f(rng), &t, &sb)) break;
        seg_point(s, t, p);
        float st = eval_sigma_t(C, K, p);
        float sn = sb - st;
        float tr = sn / sb; /* sb > 0 whenever a collision is returned */
        if (tr > 0.0f) nn += 1;
        if (adjoint && tr > 0.0f) {
            float asum = (adjoint[0] + adjoint[1]) + adjoint[2];
            scatter_sigma(C, K, p, -asum / sn);
        }
        T *= tr;
        if (T == 0.0f) break;
    }
    if (n_null) *n_null = nn;
    return T;
}

/* Test hook for the event counters only (results are unaffected): the CUDA adjoint kernel logs up to `capacity`
 * tentative collisions of a shadow walk and scatters the NEE adjoint from the log; a walk with more collisions is
 * walked a second time like here.  With a capacity set, only the second walks the CUDA path really skips are
 * booked as "replay" events; 0 (default) books all of them. */
static int g_nee_log_capacity = 0;
static uint64_t g_nee_log_overflows = 0;
void uivr_oracle_set_nee_log_capacity(int capacity) { g_nee_log_capacity = capacity; g_nee_log_overflows = 0; }
uint64_t uivr_oracle_nee_log_overflows(void) { return g_nee_log_overflows; }

/* sample_emitter_for_nee + sample_emitter (volpathsimple.py:380-433, a5/a6) for a constant
 * emitter and isotropic phase: contribution = beta * phase(1/4pi) * mis(1/2) * (Le*4pi) * T,
 * folded to beta * (0.5 Le) * T.  In the adjoint the walk is replayed from a cloned sampler
 * with adjoint = dL * contribution (:393-401). */
static void nee(const ctx_t* C, counters_t* K, const float p[3], const float beta[3],
                rng_t* rng, const float* dL, float contrib[3]) {
    float xi1 = rng_f(rng), xi2 = rng_f(rng);
    float w[3], wgt[3];
    int worked = 1;
    if (C->sc->env_data) {
        /* envmap: throughput * phase_val * mis_weight(ds.pdf, phase_pdf) * (Le / ds.pdf) * T  (:385-391, :419-423) */
        float pdf, le[3];
        env_sample(C->sc, xi1, xi2, w, &pdf, le);
        worked = pdf != 0.0f; /* sampling_worked (:421-423): no shadow ray, no draws */
        float mis = mis_power(pdf, INV_4PI);
        for (int c = 0; c < 3; ++c) wgt[c] = worked ? ((beta[c] * INV_4PI) * mis) * (le[c] / pdf) : 0.0f;
    } else {
        uniform_sphere(xi1, xi2, w);
        for (int c = 0; c < 3; ++c) wgt[c] = beta[c] * C->half_le[c];
    }
    seg_t s;
    int valid = worked && make_segment(C, p, w, &s);
    rng_t clone = *rng;
    float T = valid ? ratio_track(C, K, &s, rng, NULL, NULL) : 0.0f;
    for (int c = 0; c < 3; ++c) contrib[c] = wgt[c] * T;
    if (dL && valid) {
        float adj[3];
        for (int c = 0; c < 3; ++c) adj[c] = dL[c] * contrib[c];
        uint64_t d0 = clone.draws;
        counters_t before = *K;
        int n_null = 0;
        ratio_track(C, K, &s, &clone, adj, &n_null);
        rng->draws += clone.draws - d0; /* the replay consumes (cloned) draws too */
        if (g_nee_log_capacity > 0 && n_null > g_nee_log_capacity) {
            __atomic_fetch_add(&g_nee_log_overflows, 1, __ATOMIC_RELAXED);
        } else {
            for (int k = 0; k < UIVR_ORC_NUM_COUNTERS; ++k) K->r[k] += K->c[k] - before.c[k];
            K->r[UIVR_ORC_RNG_DRAWS] += clone.draws - d0;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Medium::sample_interaction_drt  [UPSTREAM App. B.6, a15]                              */
/* ------------------------------------------------------------------------------------ */

static int drt_sample(const ctx_t* C, counters_t* K, const seg_t* s, rng_t* alt, float* t_sel,
                      float* st_sel, float* D_out) {
    walk_t w;
    walk_init(C, K, s, &w);
    float T = 1.0f, D = 0.0f;
    int found = 0;
    for (;;) {
        float t, sb, p[3];
        if (!walk_next(C, K, &w, rng_f(alt), &t, &sb)) break;
        float u2 = rng_f(alt);
        seg_point(s, t, p);
        float st = eval_sigma_t(C, K, p);
        float wi = T / sb;
        D += wi;
        if (u2 <= wi / D) {
            *t_sel = t;
            *st_sel = st;
            found = 1;
        }
        T *= (sb - st) / sb;
        if (!(T > 0.0f)) break;
    }
    *D_out = D;
    return found;
}

/* ------------------------------------------------------------------------------------ */
/* DRTReservoir (volpathsimple.py:730-765, a13)                                          */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    float wsum[3], wcur[3];
    seg_t seg;
    int depth, valid;
} reservoir_t;

static inline float mean3(const float v[3]) { return ((v[0] + v[1]) + v[2]) * (1.0f / 3.0f); }

static void reservoir_update(reservoir_t* r, const seg_t* s, int depth, const float weight[3],
                             float u) {
    float ratio[3];
    for (int c = 0; c < 3; ++c) {
        r->wsum[c] += weight[c];
        ratio[c] = weight[c] / r->wsum[c];
    }
    if (u <= mean3(ratio)) {
        for (int c = 0; c < 3; ++c) r->wcur[c] = weight[c];
        r->seg = *s;
        r->depth = depth;
        r->valid = 1;
    }
}

/* ------------------------------------------------------------------------------------ */
/* VolpathSimpleIntegrator.sample main loop  (volpathsimple.py:110-285)                  */
/* ------------------------------------------------------------------------------------ */

static void drt_backprop(const ctx_t* C, counters_t* K, const seg_t* s, int depth,
                         const float adjoint[3], rng_t* alt);

typedef struct {
    seg_t seg;
    int depth, active, escaped, has_scattered;
} path_state_t;

/* adjoint == 0: primal (result accumulates, envmap added at the end).
 * adjoint == 1: `R` enters as the primal radiance (state_in) and is consumed by path replay. */
/* test hook: the adjoint takes "radiance still to come" as L - (radiance gathered so far) instead of subtracting
 * every NEE contribution from a running L (:214).  Same quantity, different rounding; this is the form the CUDA
 * slot-pool pipeline uses (its adjoint replay gathers L itself).  tests/test_oracle.py bounds the difference. */
static int g_remaining_by_difference = 0;
void uivr_oracle_set_remaining_by_difference(int enable) { g_remaining_by_difference = enable; }

static void path_loop(const ctx_t* C, counters_t* K, int adjoint, rng_t* rng, rng_t* alt,
                      path_state_t ps, const float dL[3], float R[3]) {
    const uivr_oracle_scene* sc = C->sc;
    float beta[3] = {1.0f, 1.0f, 1.0f};
    const float L0[3] = {R[0], R[1], R[2]};
    float gathered[3] = {0.0f, 0.0f, 0.0f};
    seg_t seg = ps.seg;
    int depth = ps.depth, active = ps.active, escaped = ps.escaped;
    int has_scattered = ps.has_scattered;
    reservoir_t rsv;
    memset(&rsv, 0, sizeof(rsv));
    const int use_rsv = adjoint && sc->use_drt && sc->use_drt_subsampling;

    while (active) {
        /* :117-121 Russian roulette never fires (rr_depth = max_depth+1000, opt_config.py:105)
         * but its draw is consumed; zero throughput kills the path */
        rng_f(rng);
        if (beta[0] == 0.0f && beta[1] == 0.0f && beta[2] == 0.0f) break;

        /* :126 sample_real_interaction (:323-377): analog delta tracking */
        walk_t w;
        walk_init(C, K, &seg, &w);
        int did_scatter = 0;
        float t_real = 0.0f, sigma_t = 0.0f, p[3] = {0, 0, 0};
        for (;;) {
            float t, sb;
            if (!walk_next(C, K, &w, rng_f(rng), &t, &sb)) break; /* escaped: no 2nd draw */
            seg_point(&seg, t, p);
            float st = eval_sigma_t(C, K, p);
            float r = st / sb;
            if (rng_f(rng) >= r) continue; /* null collision (:359) */
            did_scatter = 1;
            t_real = t;
            sigma_t = st;
            break;
        }
        const int did_escape = !did_scatter;
        if (did_scatter) {
            has_scattered = 1;
            K->c[UIVR_ORC_REAL_COLLISIONS]++;
        }

        /* :141 albedo */
        float albedo[3] = {1.0f, 1.0f, 1.0f};
        if (did_scatter) eval_albedo(C, K, p, albedo);

        if (adjoint) {
            if (sc->use_drt) {
                if (use_rsv) {
                    /* :521-539 reservoir over path segments, weight = throughput */
                    reservoir_update(&rsv, &seg, depth, beta, rng_f(alt));
                } else {
                    /* quadratic mode: DRT at every vertex, adjoint = dL * throughput (:146) */
                    float adj[3];
                    for (int c = 0; c < 3; ++c) adj[c] = dL[c] * beta[c];
                    drt_backprop(C, K, &seg, depth, adj, alt);
                }
            }
            /* :152-172 free-flight scattering gradient */
            if ((!sc->use_drt || sc->use_drt_mis) && did_scatter) {
                float m = 1.0f;
                if (sc->use_drt && sc->use_drt_mis) {
                    float s2 = sigma_t * sigma_t;
                    m = s2 / (1.0f + s2);
                }
                float inv_pdf = 1.0f / sigma_t;
                float gs = 0.0f, ga[3];
                for (int c = 0; c < 3; ++c) {
                    float Li = R[c] / (albedo[c] > 1e-8f ? albedo[c] : 1e-8f);
                    float term = ((m * dL[c]) * Li) * inv_pdf;
                    gs = FMA(term, albedo[c], gs);
                    ga[c] = term * sigma_t;
                }
                scatter_sigma(C, K, p, gs);
                scatter_albedo(C, K, p, ga);
            }
            /* :181-189 + :584-607 transmittance gradient, 4 uniform taps on the segment */
            {
                float interval = did_escape ? seg.tmax : t_real;
                float aw = FMA(dL[2], R[2], FMA(dL[1], R[1], dL[0] * R[0]));
                float g = -(aw * (interval * 0.25f));
                for (int k = 0; k < 4; ++k) {
                    float tk = rng_f(alt) * interval, pk[3];
                    seg_point(&seg, tk, pk);
                    scatter_sigma(C, K, pk, g);
                }
            }
        }

        /* :193-200 */
        for (int c = 0; c < 3; ++c) beta[c] *= albedo[c];
        if (did_scatter) depth += 1;
        active = did_scatter && (depth < sc->max_depth);

        /* :206-215 emitter sampling */
        if (sc->use_nee && did_scatter && active) {
            float contrib[3];
            nee(C, K, p, beta, rng, adjoint ? dL : NULL, contrib);
            for (int c = 0; c < 3; ++c) R[c] = adjoint ? R[c] - contrib[c] : R[c] + contrib[c];
            if (adjoint && g_remaining_by_difference)
                for (int c = 0; c < 3; ++c) {
                    gathered[c] = gathered[c] + contrib[c];
                    R[c] = L0[c] - gathered[c];
                }
        }

        /* :221-235 phase sampling (draws masked by did_scatter, not by active) */
        if (did_scatter) {
            rng_f(rng);
            float xi1 = rng_f(rng), xi2 = rng_f(rng), wo[3];
            uniform_sphere(xi1, xi2, wo);
            if (!make_segment(C, p, wo, &seg)) active = 0; /* :240-241 accidental escape */
        }
        /* :244-245 */
        if (did_escape) escaped = 1;
    }

    /* :249-259 DRT with the reservoir's segment */
    if (use_rsv && rsv.valid) {
        float d = mean3(rsv.wcur), ws = mean3(rsv.wsum), adj[3];
        for (int c = 0; c < 3; ++c) {
            float W = (d != 0.0f) ? (ws * rsv.wcur[c]) / d : 0.0f;
            adj[c] = W * dL[c];
        }
        drt_backprop(C, K, &rsv.seg, rsv.depth, adj, alt);
    }

    /* :263-285 envmap (primal only) */
    if (!adjoint && escaped && !(depth <= 0 && sc->hide_emitters)) {
        if (sc->env_data) {
            /* :270-285 emitter.eval(si) with hit_mis_weight = mis_weight(last_scatter_direction_pdf,
             * has_scattered ? emitter.pdf_direction : 0) */
            float le[3], pdf;
            env_eval(sc, seg.d, le, &pdf);
            float wmis = 1.0f;
            if (sc->use_nee) wmis = mis_power(has_scattered ? INV_4PI : 1.0f, has_scattered ? pdf : 0.0f);
            for (int c = 0; c < 3; ++c) R[c] = R[c] + (beta[c] * wmis) * le[c];
        } else {
            float wmis = (sc->use_nee && has_scattered) ? 0.5f : 1.0f;
            for (int c = 0; c < 3; ++c) R[c] = FMA(beta[c] * wmis, sc->radiance[c], R[c]);
        }
    }
}

/* backpropagate_scattering_drt without reservoir (:543-581) + sample_recursive (:610-655) */
static void drt_backprop(const ctx_t* C, counters_t* K, const seg_t* s, int depth,
                         const float adjoint[3], rng_t* alt) {
    const uivr_oracle_scene* sc = C->sc;
    float t_sel = 0, st = 0, D = 0;
    if (!drt_sample(C, K, s, alt, &t_sel, &st, &D)) return;
    float p[3], albedo[3];
    seg_point(s, t_sel, p);
    eval_albedo(C, K, p, albedo);

    /* sample_recursive: NEE with unit throughput ... */
    float Li[3] = {0, 0, 0};
    const float one[3] = {1.0f, 1.0f, 1.0f};
    if (sc->use_nee) nee(C, K, p, one, alt, NULL, Li);
    /* ... plus a phase-sampled detached continuation */
    rng_f(alt);
    float xi1 = rng_f(alt), xi2 = rng_f(alt), wo[3];
    uniform_sphere(xi1, xi2, wo);
    path_state_t ps;
    int ok = make_segment(C, p, wo, &ps.seg);
    ps.depth = depth + 1;
    ps.active = ok && (ps.depth < sc->max_depth);
    ps.escaped = 0;
    ps.has_scattered = ps.active;
    if (ps.active) {
        rng_f(alt); /* :99 alt_seed_rnd of the recursive sample() */
        float Lr[3] = {0, 0, 0};
        path_loop(C, K, 0, alt, NULL, ps, NULL, Lr);
        for (int c = 0; c < 3; ++c) Li[c] += Lr[c];
    }

    float m = sc->use_drt_mis ? 1.0f / (1.0f + st * st) : 1.0f;
    float gs = 0.0f, ga[3];
    for (int c = 0; c < 3; ++c) {
        float term = ((m * D) * adjoint[c]) * Li[c];
        gs = FMA(term, albedo[c], gs);
        ga[c] = term * st;
    }
    scatter_sigma(C, K, p, gs);
    scatter_albedo(C, K, p, ga);
}

/* ------------------------------------------------------------------------------------ */
/* Driver: ray generation, reach_medium, film  (batched.py:134-197, 212-326; App. B.3-B.4)*/
/* ------------------------------------------------------------------------------------ */

/* returns: 0 = missed the box (escaped), 1 = entered the medium, 2 = dead (corner case) */
/* perspective sensor frame: origin[3] left[3] up[3] dir[3] tan_x tan_y near_clip */
static int camera_segment_frame(const ctx_t* C, const float* F, float u, float v, seg_t* s);

static int camera_segment(const ctx_t* C, uint32_t pix, float jx, float jy, seg_t* s) {
    const uivr_oracle_scene* sc = C->sc;
    uint32_t px = pix % (uint32_t) sc->width, py = pix / (uint32_t) sc->width;
    float u = ((float) px + jx) * (1.0f / (float) sc->width);
    float v = ((float) py + jy) * (1.0f / (float) sc->height);
    float F[15];
    for (int a = 0; a < 3; ++a) {
        F[a] = sc->cam_origin[a]; F[3 + a] = sc->cam_left[a]; F[6 + a] = sc->cam_up[a]; F[9 + a] = sc->cam_dir[a];
    }
    F[12] = sc->tan_x; F[13] = sc->tan_y; F[14] = sc->near_clip;
    return camera_segment_frame(C, F, u, v, s);
}

/* sample_batch_pixels (batched.py:397-423): sensor + pixel of batch element b */
void uivr_oracle_batch_element(const uivr_oracle_batch* B, uint32_t b, uint32_t out[3]) {
    rng_t q;
    sampler_seed(&q, B->seed_pixels, b);
    float u0 = rng_f(&q), u1 = rng_f(&q), u2 = rng_f(&q);
    uint32_t si = (uint32_t) ((float) B->n_sensors * u0);
    uint32_t px = (uint32_t) ((float) B->film_w * u1), py = (uint32_t) ((float) B->film_h * u2);
    out[0] = si < (uint32_t) B->n_sensors ? si : (uint32_t) B->n_sensors - 1u;
    out[1] = px < (uint32_t) B->film_w ? px : (uint32_t) B->film_w - 1u;
    out[2] = py < (uint32_t) B->film_h ? py : (uint32_t) B->film_h - 1u;
}

/* sample_batch_rays (batched.py:426-467): ray of wavefront entry idx = b * spp + j */
static int batch_segment(const ctx_t* C, uint32_t idx, uint32_t spp, seg_t* s) {
    const uivr_oracle_batch* B = C->batch;
    uint32_t e[3];
    uivr_oracle_batch_element(B, idx / spp, e);
    rng_t o;
    sampler_seed(&o, B->seed_offsets, idx);
    float jx = rng_f(&o), jy = rng_f(&o);
    float u = ((float) e[1] + jx) * (1.0f / (float) B->film_w);
    float v = ((float) e[2] + jy) * (1.0f / (float) B->film_h);
    return camera_segment_frame(C, B->sensors + 16 * (size_t) e[0], u, v, s);
}

/* perspective sensor: primary ray through film position (u, v), in the medium's local space */
static void camera_ray_local(const ctx_t* C, const float* F, float u, float v, float ol[3], float dl[3]) {
    const uivr_oracle_scene* sc = C->sc;
    float cx = F[12] * FMA(-2.0f, u, 1.0f);
    float cy = F[13] * FMA(-2.0f, v, 1.0f);
    float d[3], o[3];
    for (int a = 0; a < 3; ++a) d[a] = FMA(cx, F[3 + a], FMA(cy, F[6 + a], F[9 + a]));
    float len = sqrtf(FMA(d[0], d[0], FMA(d[1], d[1], d[2] * d[2])));
    float inv_len = 1.0f / len;
    float near_t = F[14] * len;
    for (int a = 0; a < 3; ++a) {
        d[a] *= inv_len;
        o[a] = FMA(near_t, d[a], F[a]);
    }
    const float* M = sc->to_local;
    for (int a = 0; a < 3; ++a)
        ol[a] = FMA(M[4 * a + 0], o[0], FMA(M[4 * a + 1], o[1], FMA(M[4 * a + 2], o[2], M[4 * a + 3])));
    dir_to_local(M, d, dl);
}

/* scene.ray_intersect against the medium box from a ray origin not known to be inside
 * (volpathsimple.py:298): slab test against local [0,1]^3.
 * 0 = miss, 1 = enters at *t, 2 = origin inside, first hit is the far wall at *t */
static int box_entry(const float ol[3], const float dl[3], float* t) {
    float tn = -UIVR_INF, tf = UIVR_INF;
    for (int a = 0; a < 3; ++a) {
        if (dl[a] != 0.0f) {
            float inv = 1.0f / dl[a];
            float t0 = (0.0f - ol[a]) * inv, t1 = (1.0f - ol[a]) * inv;
            float lo = t0 < t1 ? t0 : t1, hi = t0 < t1 ? t1 : t0;
            if (lo > tn) tn = lo;
            if (hi < tf) tf = hi;
        } else if (ol[a] < 0.0f || ol[a] > 1.0f) {
            return 0;
        }
    }
    if (!(tn <= tf) || !(tf > 0.0f)) return 0;
    if (!(tn > 0.0f)) { *t = tf; return 2; }
    *t = tn;
    return 1;
}

/* si.spawn_ray across the null boundary INTO the medium (:306): entry point clamped into
 * [eps, 1-eps] */
static inline void entry_spawn(const float ol[3], const float dl[3], float tn, float o[3]) {
    for (int a = 0; a < 3; ++a) {
        float e = FMA(tn, dl[a], ol[a]);
        o[a] = e < ENTRY_EPS ? ENTRY_EPS : (e > 1.0f - ENTRY_EPS ? 1.0f - ENTRY_EPS : e);
    }
}

static int camera_segment_frame(const ctx_t* C, const float* F, float u, float v, seg_t* s) {
    float ol[3], dl[3], tn = 0.0f;
    camera_ray_local(C, F, u, v, ol, dl);
    for (int a = 0; a < 3; ++a) s->d[a] = dl[a]; /* also for rays that miss: the envmap lookup needs it */
    /* reach_medium (volpathsimple.py:292-319) */
    int hit = box_entry(ol, dl, &tn);
    if (hit != 1) return hit; /* 2: origin inside, the re-spawned ray misses */
    entry_spawn(ol, dl, tn, s->o);
    s->tmax = exit_distance(s);
    return (s->tmax > 0.0f && s->tmax < UIVR_INF) ? 1 : 2;
}

/* One full `sample()` call from the camera.  adjoint: R enters as state_in. */
static void sample_from_camera(const ctx_t* C, counters_t* K, int adjoint, uint32_t seed,
                               uint32_t alt_seed, uint32_t idx, uint32_t spp,
                               const float dL[3], float R[3]) {
    rng_t rng, alt;
    sampler_seed(&rng, seed, idx);
    if (adjoint) sampler_seed(&alt, alt_seed, idx);
    path_state_t ps;
    int status;
    if (C->batch) {
        status = batch_segment(C, idx, spp, &ps.seg);  /* the path sampler draws no jitter (batched.py:390) */
    } else {
        float jx = rng_f(&rng), jy = rng_f(&rng);
        status = camera_segment(C, idx / spp, jx, jy, &ps.seg);
    }
    rng_f(&rng); /* :71 colour-channel placeholder draw */
    ps.depth = 0;
    ps.active = (status == 1);
    ps.escaped = (status == 0);
    ps.has_scattered = 0;
    if (ps.active) {
        rng_f(&rng); /* :99 alt_seed_rnd */
        if (!adjoint) K->c[UIVR_ORC_CAMERA_HITS]++;
    }
    path_loop(C, K, adjoint, &rng, adjoint ? &alt : NULL, ps, dL, R);
    K->c[UIVR_ORC_RNG_DRAWS] += rng.draws + (adjoint ? alt.draws : 0);
}

/* ------------------------------------------------------------------------------------ */
/* NeRFIntegrator.sample (python/integrators/nerf.py:47-147): emission-absorption ray      */
/* marching, `queries_per_ray` forward-looking steps, one jitter draw per ray; the         */
/* adjoint replays the steps with path replay and backpropagates per step (:117-124).      */
/* ------------------------------------------------------------------------------------ */

static void nerf_sample(const ctx_t* C, counters_t* K, int adjoint, uint32_t seed, uint32_t idx,
                        uint32_t spp, const float dL[3], float R[3]) {
    const uivr_oracle_scene* sc = C->sc;
    const uivr_oracle_nerf* N = C->nerf;
    rng_t rng;
    sampler_seed(&rng, seed, idx);
    seg_t seg;
    int status;
    if (C->batch) {
        status = batch_segment(C, idx, spp, &seg);
    } else {
        float jx = rng_f(&rng), jy = rng_f(&rng);
        status = camera_segment(C, idx / spp, jx, jy, &seg);
    }
    const int active = (status == 1), escaped = (status == 0); /* :69-77 */
    float wsum = 0.0f, T = 1.0f;
    if (active) {
        if (!adjoint) K->c[UIVR_ORC_CAMERA_HITS]++;
        const int Q = N->queries_per_ray;
        /* :6-18 step_size_for_marching / query_t_for_step with mint = 0 */
        const float step = N->jittering_enabled ? seg.tmax / (float) Q : seg.tmax / (float) (Q - 1);
        const float jit = rng_f(&rng); /* :87, drawn even when jittering is off */
        float t_a = 0.0f;
        for (int j = 0; j < Q; ++j) {
            const float sj = (float) (j + 1);
            const float t_b = N->jittering_enabled ? step * (sj + jit) : step * sj;
            const float dt = t_b - t_a;
            float p[3], em[3];
            seg_point(&seg, t_b, p);
            const float raw = eval_sigma_t(C, K, p);                   /* :152-155 */
            const float sigma = (N->activation == 1 && !(raw > 0.0f)) ? 0.0f : raw;
            eval_albedo(C, K, p, em);                                  /* get_emission :160 */
            const int last = !(j + 1 < Q);
            const float a = last ? 1.0f : exp_exact(-(sigma * dt));    /* :103-105 */
            const float weight = (1.0f - a) * T;
            const float safe = a + 1e-10f;
            for (int c = 0; c < 3; ++c) {
                const float we = weight * em[c];
                R[c] = adjoint ? R[c] - we : R[c] + we;                /* :109-112 */
            }
            if (adjoint && !last) {
                /* :117-124  d/d emission_c = dL_c weight;
                 * d/d sigma = sum_c dL_c (em_c dt a T - (R_c / safe) dt a) */
                const float da = dt * a;
                float gs = 0.0f, ge[3];
                for (int c = 0; c < 3; ++c) {
                    ge[c] = dL[c] * weight;
                    const float inner = FMA(em[c], da * T, -((R[c] / safe) * da));
                    gs = FMA(dL[c], inner, gs);
                }
                if (N->activation == 1 && !(raw > 0.0f)) gs = 0.0f;
                scatter_sigma(C, K, p, gs);
                scatter_albedo(C, K, p, ge);
            }
            t_a = t_b;
            if (!last) { /* :114-120: masked by the updated still_walking */
                T *= safe;
                wsum += weight;
            }
        }
    }
    /* :134-143 composite with the background emitter (in both modes) */
    int active_e = escaped || active;
    if (N->hide_emitters) active_e = active_e && (wsum > 0.0f);
    if (active_e) {
        float le[3] = {sc->radiance[0], sc->radiance[1], sc->radiance[2]}, pdf;
        if (sc->env_data) env_eval(sc, seg.d, le, &pdf);
        for (int c = 0; c < 3; ++c) R[c] += (1.0f - wsum) * le[c];
    }
    K->c[UIVR_ORC_RNG_DRAWS] += rng.draws;
}

/* ------------------------------------------------------------------------------------ */
/* Threaded entry points                                                                 */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    ctx_t* C;
    int backward;
    uint32_t seed, alt_seed, spp;
    const uivr_oracle_shard* shard;
    const float* grad_image;
    double* image_acc; /* H*W*3 (forward) */
    float* sample_L;
    uint32_t* next_pixel;
    counters_t K;
    counters_t Kp;     /* backward: the share of K spent in the primal pass (batched.py:255-264) */
} job_t;

/* Events of the primal pass inside the most recent backward call (included in its `counters`).  The CUDA path's
 * adjoint replay gathers the primal radiance itself instead of running that pass, so its event counts are those
 * of the backward minus these -- which is how the parity tests compare them. */
static uint64_t g_last_backward_primal[UIVR_ORC_NUM_COUNTERS];
void uivr_oracle_last_backward_primal_counters(uint64_t* out) {
    memcpy(out, g_last_backward_primal, sizeof(g_last_backward_primal));
}
/* ... and of the NEE adjoint's second walk over every shadow segment (:393-401): the CUDA path logs the
 * tentative collisions of the first walk instead of walking again */
static uint64_t g_last_backward_replay[UIVR_ORC_NUM_COUNTERS];
void uivr_oracle_last_backward_replay_counters(uint64_t* out) {
    memcpy(out, g_last_backward_replay, sizeof(g_last_backward_replay));
}

static void* worker(void* arg) {
    job_t* J = (job_t*) arg;
    const uivr_oracle_scene* sc = J->C->sc;
    const uint32_t npix = (uint32_t) sc->width * (uint32_t) sc->height;
    const float inv_spp = 1.0f / (float) J->spp;
    for (;;) {
        uint32_t pix = __atomic_fetch_add(J->next_pixel, 1u, __ATOMIC_RELAXED);
        if (pix >= npix) break;
        if (J->shard && J->shard->shard_count > 1 &&
            (int) ((pix / (uint32_t) J->shard->shard_block) % (uint32_t) J->shard->shard_count) != J->shard->shard_rank)
            continue;
        double acc[3] = {0, 0, 0};
        for (uint32_t s = 0; s < J->spp; ++s) {
            uint32_t idx = pix * J->spp + s;
            float L[3] = {0, 0, 0};
            counters_t* Kprimal = J->backward ? &J->Kp : &J->K;
            if (J->C->nerf) nerf_sample(J->C, Kprimal, 0, J->seed, idx, J->spp, NULL, L);
            else sample_from_camera(J->C, Kprimal, 0, J->seed, 0, idx, J->spp, NULL, L);
            J->K.c[UIVR_ORC_SAMPLES]++;
            if (J->sample_L) memcpy(J->sample_L + 3 * (size_t) idx, L, sizeof(L));
            if (!J->backward) {
                for (int c = 0; c < 3; ++c) acc[c] += (double) L[c];
            } else {
                /* batched.py:272-306: box film => dL = grad_image[pixel] / spp */
                float dL[3];
                for (int c = 0; c < 3; ++c) dL[c] = J->grad_image[3 * (size_t) pix + c] * inv_spp;
                if (J->C->nerf) nerf_sample(J->C, &J->K, 1, J->seed, idx, J->spp, dL, L);
                else sample_from_camera(J->C, &J->K, 1, J->seed, J->alt_seed, idx, J->spp, dL, L);
            }
        }
        if (!J->backward)
            for (int c = 0; c < 3; ++c) J->image_acc[3 * (size_t) pix + c] = acc[c];
    }
    return NULL;
}

static int setup_ctx(ctx_t* C, const uivr_oracle_scene* sc, const float* sigma_t, const float* albedo) {
    memset(C, 0, sizeof(*C));
    C->sc = sc;
    C->sigma_t = sigma_t;
    C->albedo = albedo;
    int32_t m[3];
    for (int a = 0; a < 3; ++a) {
        m[a] = (sc->majorant_factor > 1) ? sc->res[a] / sc->majorant_factor : 1;
        if (m[a] < 1) m[a] = 1;
    }
    C->majorant = (float*) malloc(sizeof(float) * (size_t) m[0] * m[1] * m[2]);
    if (!C->majorant) return -1;
    uivr_oracle_build_majorant(sigma_t, sc->res, sc->scale, sc->majorant_factor, C->mres, C->majorant);
    C->exit_mask = (uint8_t*) malloc((size_t) m[0] * m[1] * m[2]);
    if (!C->exit_mask) { free(C->majorant); return -1; }
    uivr_oracle_build_exit_mask(C->majorant, C->mres, C->exit_mask);
    for (int a = 0; a < 3; ++a) {
        C->mcs[a] = 1.0f / (float) C->mres[a];
        C->half_le[a] = 0.5f * sc->radiance[a];
    }
    return 0;
}

static int run(ctx_t* C, int backward, uint32_t seed, uint32_t spp, const uivr_oracle_shard* shard,
               int nthreads, const float* grad_image, double* image_acc, float* sample_L,
               uint64_t* counters) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    uint32_t next_pixel = 0;
    job_t* jobs = (job_t*) calloc((size_t) nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*) calloc((size_t) nthreads, sizeof(pthread_t));
    if (!jobs || !th) return -1;
    uint32_t alt_seed = backward ? alt_seed_after(seed, C->batch ? 1 : 3) : 0;
    for (int i = 0; i < nthreads; ++i) {
        jobs[i].C = C;
        jobs[i].backward = backward;
        jobs[i].seed = seed;
        jobs[i].alt_seed = alt_seed;
        jobs[i].spp = spp;
        jobs[i].shard = shard;
        jobs[i].grad_image = grad_image;
        jobs[i].image_acc = image_acc;
        jobs[i].sample_L = sample_L;
        jobs[i].next_pixel = &next_pixel;
        if (i > 0) pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    worker(&jobs[0]);
    for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
    if (backward) {
        memset(g_last_backward_primal, 0, sizeof(g_last_backward_primal));
        memset(g_last_backward_replay, 0, sizeof(g_last_backward_replay));
    }
    for (int i = 0; i < nthreads; ++i)
        for (int k = 0; k < UIVR_ORC_NUM_COUNTERS; ++k) {
            if (counters) counters[k] += jobs[i].K.c[k] + jobs[i].Kp.c[k];
            if (backward) {
                g_last_backward_primal[k] += jobs[i].Kp.c[k];
                g_last_backward_replay[k] += jobs[i].K.r[k];
            }
        }
    free(jobs);
    free(th);
    return 0;
}

int uivr_oracle_render_forward(const uivr_oracle_scene* scene, const float* sigma_t,
                               const float* albedo, uint32_t seed, int32_t spp,
                               const uivr_oracle_shard* shard, int nthreads,
                               float* image_out, float* sample_L_out, uint64_t* counters) {
    if (!scene || !sigma_t || !albedo || !image_out || spp < 1) return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, albedo)) return -1;
    size_t n = (size_t) scene->width * scene->height * 3;
    double* acc = (double*) calloc(n, sizeof(double));
    if (sample_L_out) memset(sample_L_out, 0, sizeof(float) * n * (size_t) spp);
    int rc = acc ? run(&C, 0, seed, (uint32_t) spp, shard, nthreads, NULL, acc, sample_L_out, counters) : -1;
    if (!rc) {
        float inv_spp = 1.0f / (float) spp;
        for (size_t i = 0; i < n; ++i) image_out[i] = (float) acc[i] * inv_spp;
    }
    free(acc);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

int uivr_oracle_render_backward(const uivr_oracle_scene* scene, const float* sigma_t,
                                const float* albedo, const float* grad_image,
                                uint32_t seed_grad, int32_t spp_grad,
                                const uivr_oracle_shard* shard, int nthreads,
                                double* dsigma_out, double* dalbedo_out,
                                float* sample_L_out, uint64_t* counters) {
    if (!scene || !sigma_t || !albedo || !grad_image || !dsigma_out || !dalbedo_out || spp_grad < 1)
        return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, albedo)) return -1;
    size_t nvox = (size_t) scene->res[0] * scene->res[1] * scene->res[2];
    memset(dsigma_out, 0, sizeof(double) * nvox);
    memset(dalbedo_out, 0, sizeof(double) * nvox * 3);
    C.dsigma = dsigma_out;
    C.dalbedo = dalbedo_out;
    if (sample_L_out)
        memset(sample_L_out, 0, sizeof(float) * 3 * (size_t) scene->width * scene->height * (size_t) spp_grad);
    int rc = run(&C, 1, seed_grad, (uint32_t) spp_grad, shard, nthreads, grad_image, NULL, sample_L_out, counters);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

/* ray-batch entry points: the same drivers with C.batch set (film = B x 1) */
int uivr_oracle_render_batch_forward(const uivr_oracle_scene* scene, const uivr_oracle_batch* batch,
                                     const float* sigma_t, const float* albedo, uint32_t seed,
                                     int32_t spp, int nthreads, float* image_out,
                                     float* sample_L_out, uint64_t* counters) {
    if (!scene || !batch || !batch->sensors || batch->n_sensors < 1 || scene->height != 1 || !sigma_t || !albedo ||
        !image_out || spp < 1)
        return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, albedo)) return -1;
    C.batch = batch;
    size_t n = (size_t) scene->width * 3;
    double* acc = (double*) calloc(n, sizeof(double));
    if (sample_L_out) memset(sample_L_out, 0, sizeof(float) * n * (size_t) spp);
    int rc = acc ? run(&C, 0, seed, (uint32_t) spp, NULL, nthreads, NULL, acc, sample_L_out, counters) : -1;
    if (!rc) {
        float inv_spp = 1.0f / (float) spp;
        for (size_t i = 0; i < n; ++i) image_out[i] = (float) acc[i] * inv_spp;
    }
    free(acc);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

int uivr_oracle_render_batch_backward(const uivr_oracle_scene* scene, const uivr_oracle_batch* batch,
                                      const float* sigma_t, const float* albedo,
                                      const float* grad_image, uint32_t seed_grad, int32_t spp_grad,
                                      int nthreads, double* dsigma_out, double* dalbedo_out,
                                      float* sample_L_out, uint64_t* counters) {
    if (!scene || !batch || !batch->sensors || batch->n_sensors < 1 || scene->height != 1 || !sigma_t || !albedo ||
        !grad_image || !dsigma_out || !dalbedo_out || spp_grad < 1)
        return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, albedo)) return -1;
    C.batch = batch;
    size_t nvox = (size_t) scene->res[0] * scene->res[1] * scene->res[2];
    memset(dsigma_out, 0, sizeof(double) * nvox);
    memset(dalbedo_out, 0, sizeof(double) * nvox * 3);
    C.dsigma = dsigma_out;
    C.dalbedo = dalbedo_out;
    if (sample_L_out) memset(sample_L_out, 0, sizeof(float) * 3 * (size_t) scene->width * (size_t) spp_grad);
    int rc = run(&C, 1, seed_grad, (uint32_t) spp_grad, NULL, nthreads, grad_image, NULL, sample_L_out, counters);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

/* nerf entry points: the same drivers with C.nerf set; `emission` takes the albedo slot */
int uivr_oracle_nerf_forward(const uivr_oracle_scene* scene, const uivr_oracle_nerf* nerf,
                             const uivr_oracle_batch* batch,
                             const float* sigma_t, const float* emission, uint32_t seed, int32_t spp,
                             const uivr_oracle_shard* shard, int nthreads, float* image_out,
                             float* sample_L_out, uint64_t* counters) {
    if (!scene || !nerf || nerf->queries_per_ray < 2 || !sigma_t || !emission || !image_out || spp < 1) return -1;
    if (batch && (!batch->sensors || batch->n_sensors < 1 || scene->height != 1)) return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, emission)) return -1;
    C.nerf = nerf;
    C.batch = batch;
    size_t n = (size_t) scene->width * scene->height * 3;
    double* acc = (double*) calloc(n, sizeof(double));
    if (sample_L_out) memset(sample_L_out, 0, sizeof(float) * n * (size_t) spp);
    int rc = acc ? run(&C, 0, seed, (uint32_t) spp, shard, nthreads, NULL, acc, sample_L_out, counters) : -1;
    if (!rc) {
        float inv_spp = 1.0f / (float) spp;
        for (size_t i = 0; i < n; ++i) image_out[i] = (float) acc[i] * inv_spp;
    }
    free(acc);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

int uivr_oracle_nerf_backward(const uivr_oracle_scene* scene, const uivr_oracle_nerf* nerf,
                              const uivr_oracle_batch* batch,
                              const float* sigma_t, const float* emission, const float* grad_image,
                              uint32_t seed_grad, int32_t spp_grad, const uivr_oracle_shard* shard,
                              int nthreads, double* dsigma_out, double* demission_out,
                              float* sample_L_out, uint64_t* counters) {
    if (!scene || !nerf || nerf->queries_per_ray < 2 || !sigma_t || !emission || !grad_image || !dsigma_out ||
        !demission_out || spp_grad < 1)
        return -1;
    if (batch && (!batch->sensors || batch->n_sensors < 1 || scene->height != 1)) return -1;
    ctx_t C;
    if (setup_ctx(&C, scene, sigma_t, emission)) return -1;
    C.nerf = nerf;
    C.batch = batch;
    size_t nvox = (size_t) scene->res[0] * scene->res[1] * scene->res[2];
    memset(dsigma_out, 0, sizeof(double) * nvox);
    memset(demission_out, 0, sizeof(double) * nvox * 3);
    C.dsigma = dsigma_out;
    C.dalbedo = demission_out;
    if (sample_L_out)
        memset(sample_L_out, 0, sizeof(float) * 3 * (size_t) scene->width * scene->height * (size_t) spp_grad);
    int rc = run(&C, 1, seed_grad, (uint32_t) spp_grad, shard, nthreads, grad_image, NULL, sample_L_out, counters);
    free(C.majorant);
    free(C.exit_mask);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* upstream primitives for oracle/refshim.py (see uivr_oracle.h)                          */
/* ------------------------------------------------------------------------------------ */

struct uivr_oracle_shim {
    ctx_t C;
    uivr_oracle_scene sc;
};

uivr_oracle_shim* uivr_oracle_shim_create(const uivr_oracle_scene* scene, const float* sigma_t, const float* albedo) {
    uivr_oracle_shim* h = (uivr_oracle_shim*) calloc(1, sizeof(*h));
    if (!h) return NULL;
    h->sc = *scene;
    if (setup_ctx(&h->C, &h->sc, sigma_t, albedo)) { free(h); return NULL; }
    return h;
}

void uivr_oracle_shim_destroy(uivr_oracle_shim* h) {
    if (!h) return;
    free(h->C.majorant);
    free(h->C.exit_mask);
    free(h);
}

void uivr_oracle_shim_camera_ray(const uivr_oracle_shim* h, int n, const float* frames, const int32_t* frame_idx,
                                 const float* u, const float* v, float* o, float* d) {
    const uivr_oracle_scene* sc = &h->sc;
    float F0[16] = {0};
    for (int a = 0; a < 3; ++a) {
        F0[a] = sc->cam_origin[a]; F0[3 + a] = sc->cam_left[a]; F0[6 + a] = sc->cam_up[a]; F0[9 + a] = sc->cam_dir[a];
    }
    F0[12] = sc->tan_x; F0[13] = sc->tan_y; F0[14] = sc->near_clip;
    for (int i = 0; i < n; ++i) {
        const float* F = frames ? frames + 16 * (size_t) frame_idx[i] : F0;
        camera_ray_local(&h->C, F, u[i], v[i], o + 3 * i, d + 3 * i);
    }
}

void uivr_oracle_shim_film_uv(const uivr_oracle_shim* h, int n, const uint32_t* pix, const float* jx,
                              const float* jy, float* u, float* v) {
    const uivr_oracle_scene* sc = &h->sc;
    for (int i = 0; i < n; ++i) {
        uint32_t px = pix[i] % (uint32_t) sc->width, py = pix[i] / (uint32_t) sc->width;
        u[i] = ((float) px + jx[i]) * (1.0f / (float) sc->width);
        v[i] = ((float) py + jy[i]) * (1.0f / (float) sc->height);
    }
}

void uivr_oracle_shim_box_entry(int n, const float* o, const float* d, float* t, int32_t* kind) {
    for (int i = 0; i < n; ++i) {
        t[i] = UIVR_INF;
        kind[i] = box_entry(o + 3 * i, d + 3 * i, &t[i]);
    }
}

void uivr_oracle_shim_entry_spawn(int n, const float* o, const float* d, const float* t, float* o_new) {
    for (int i = 0; i < n; ++i) entry_spawn(o + 3 * i, d + 3 * i, t[i], o_new + 3 * i);
}

void uivr_oracle_shim_exit(int n, const float* o, const float* d, float* t, uint8_t* ok) {
    for (int i = 0; i < n; ++i) {
        seg_t s;
        for (int a = 0; a < 3; ++a) { s.o[a] = o[3 * i + a]; s.d[a] = d[3 * i + a]; }
        t[i] = exit_distance(&s);
        ok[i] = (uint8_t) (t[i] > 0.0f && t[i] < UIVR_INF);
    }
}

void uivr_oracle_shim_dir_to_local(const uivr_oracle_shim* h, int n, const float* w, float* d) {
    for (int i = 0; i < n; ++i) dir_to_local(h->sc.to_local, w + 3 * i, d + 3 * i);
}

static void shim_seg(const float* o, const float* d, float maxt, seg_t* s) {
    for (int a = 0; a < 3; ++a) {
        s->o[a] = o[a];
        s->d[a] = d[a];
        s->inv_d[a] = d[a] != 0.0f ? 1.0f / d[a] : UIVR_INF;
    }
    s->tmax = maxt;
}

void uivr_oracle_shim_sample_interaction(const uivr_oracle_shim* h, int n, const float* o, const float* d,
                                         const float* maxt, const float* u, const uint8_t* active,
                                         float* t, float* sigma_t, float* sigma_bar, uint8_t* valid) {
    counters_t K;
    memset(&K, 0, sizeof(K));
    for (int i = 0; i < n; ++i) {
        valid[i] = 0;
        t[i] = UIVR_INF;
        sigma_t[i] = 0.0f;
        sigma_bar[i] = 0.0f;
        if (!active[i]) continue;
        seg_t s;
        shim_seg(o + 3 * i, d + 3 * i, maxt[i], &s);
        walk_t w;
        walk_init(&h->C, &K, &s, &w);
        float tt, sb, p[3];
        if (!walk_next(&h->C, &K, &w, u[i], &tt, &sb)) continue;
        seg_point(&s, tt, p);
        t[i] = tt;
        sigma_t[i] = eval_sigma_t(&h->C, &K, p);
        sigma_bar[i] = sb;
        valid[i] = 1;
    }
}

void uivr_oracle_shim_sample_interaction_drt(const uivr_oracle_shim* h, int n, const float* o, const float* d,
                                             const float* maxt, uint64_t* rng_state, const uint64_t* rng_inc,
                                             const uint8_t* active, float* t, float* sigma_t, float* weight,
                                             uint8_t* valid) {
    counters_t K;
    memset(&K, 0, sizeof(K));
    for (int i = 0; i < n; ++i) {
        valid[i] = 0;
        t[i] = UIVR_INF;
        sigma_t[i] = 0.0f;
        weight[i] = 0.0f;
        if (!active[i]) continue;
        seg_t s;
        shim_seg(o + 3 * i, d + 3 * i, maxt[i], &s);
        rng_t r;
        r.state = rng_state[i];
        r.inc = rng_inc[i];
        r.draws = 0;
        float ts = 0.0f, st = 0.0f, D = 0.0f;
        valid[i] = (uint8_t) drt_sample(&h->C, &K, &s, &r, &ts, &st, &D);
        rng_state[i] = r.state;
        weight[i] = D;
        if (valid[i]) { t[i] = ts; sigma_t[i] = st; }
    }
}

void uivr_oracle_shim_lookup(const uivr_oracle_shim* h, int which, int n, const float* p, float* out) {
    counters_t K;
    memset(&K, 0, sizeof(K));
    for (int i = 0; i < n; ++i) {
        if (which == 0) {
            out[i] = eval_sigma_t(&h->C, &K, p + 3 * i);
        } else if (which == 1) {
            eval_albedo(&h->C, &K, p + 3 * i, out + 3 * i);
        } else {
            int cell[3];
            for (int a = 0; a < 3; ++a)
                cell[a] = clampi((int) floorf(p[3 * i + a] * (float) h->C.mres[a]), 0, h->C.mres[a] - 1);
            out[i] = majorant_at(&h->C, &K, cell);
        }
    }
}

void uivr_oracle_shim_scatter(const uivr_oracle_shim* h, int which, int n, const float* p, const float* g,
                              const uint8_t* mask, double* dgrid) {
    for (int i = 0; i < n; ++i) {
        if (!mask[i]) continue;
        if (which == 0) {
            float gs = h->sc.scale * g[i];
            scatter(dgrid, h->sc.res, 1, p + 3 * i, &gs);
        } else {
            scatter(dgrid, h->sc.res, 3, p + 3 * i, g + 3 * i);
        }
    }
}

void uivr_oracle_shim_env_eval(const uivr_oracle_shim* h, int n, const float* d, float* le, float* pdf) {
    for (int i = 0; i < n; ++i) env_eval(&h->sc, d + 3 * i, le + 3 * i, pdf + i);
}

void uivr_oracle_shim_env_sample(const uivr_oracle_shim* h, int n, const float* xi1, const float* xi2,
                                 float* d, float* pdf, float* le) {
    for (int i = 0; i < n; ++i) {
        float w[3];
        env_sample(&h->sc, xi1[i], xi2[i], w, pdf + i, le + 3 * i);
        dir_to_local(h->sc.to_local, w, d + 3 * i);
    }
}

void uivr_oracle_shim_uniform_sphere(int n, const float* xi1, const float* xi2, float* w) {
    for (int i = 0; i < n; ++i) uniform_sphere(xi1[i], xi2[i], w + 3 * i);
}

void uivr_oracle_shim_fma(int n, const float* a, const float* b, const float* c, float* out) {
    for (int i = 0; i < n; ++i) out[i] = FMA(a[i], b[i], c[i]);
}

/* ------------------------------------------------------------------------------------ */
/* optimiser step: mi.ad.Adam + enforce_valid_params (optimize.py:169-179, :352-353)    */
/* ------------------------------------------------------------------------------------ */

float uivr_oracle_adam_step_size(float lr, float beta1, float beta2, int32_t t) {
    const double b1t = pow((double) beta1, (double) t), b2t = pow((double) beta2, (double) t);
    return (float) ((double) lr * sqrt(1.0 - b2t) / (1.0 - b1t));
}

void uivr_oracle_adam_step(float* param, const float* grad, float* m, float* v, uint64_t n,
                           float lr, float beta1, float beta2, float eps, int32_t t,
                           float lo, float hi) {
    const float step = uivr_oracle_adam_step_size(lr, beta1, beta2, t);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    for (uint64_t i = 0; i < n; ++i) {
        const float g = grad[i];
        const float mi = FMA(beta1, m[i], omb1 * g);
        const float vi = FMA(beta2, v[i], (omb2 * g) * g);
        float p = param[i] - (step * mi) / (sqrtf(vi) + eps);
        p = p < lo ? lo : p;
        p = p > hi ? hi : p;
        m[i] = mi;
        v[i] = vi;
        param[i] = p;
    }
}
