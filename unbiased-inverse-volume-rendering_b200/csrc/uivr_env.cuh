// uivr_env.cuh -- lat-long environment emitter (Mitsuba `envmap`; python/integrators/volpathsimple.py
// :262-285 Emitter::eval + pdf_direction on escape, :419 Scene::sample_emitter_direction for NEE).
//
// Conventions and tables: scene.py EnvMap (vertices (H, W+1) x float4 = RGB radiance + sampling
// density of the bilinear patch (y, x); row marginal CDF; per-row conditional CDFs).  Same
// operation order as the oracle (oracle/uivr_oracle.c "envmap emitter"): bit-identical results.
#pragma once

#include "uivr_device.cuh"

namespace uivr {

#define UIVR_INV_4PI 0.07957747154594767f
#define UIVR_INV_2PI2 0.05066059182116889f
#define UIVR_ENV_EPS2 3.5527137e-15f
#define UIVR_ONE_MINUS_EPS 0.99999994f

// atan2(y, x) / (2 pi) in [-0.5, 0.5]: octant reduction + Cephes atanf polynomial
UIVR_DEV float atan2_turns(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float hi = ax > ay ? ax : ay, lo = ax > ay ? ay : ax;
    const float a = hi > 0.0f ? lo / hi : 0.0f;
    float off = 0.0f, t = a;
    if (a > 0.41421356f) {
        t = (a - 1.0f) / (a + 1.0f);
        off = 0.78539816f;
    }
    const float z = t * t;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    float r = (fmaf(p * z, t, t) + off) * 0.15915494f;
    if (ay > ax) r = 0.25f - r;
    if (x < 0.0f) r = 0.5f - r;
    if (y < 0.0f) r = -r;
    return r;
}

UIVR_DEV void mat3_apply(const float M[9], float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = fmaf(M[0], x, fmaf(M[1], y, M[2] * z));
    oy = fmaf(M[3], x, fmaf(M[4], y, M[5] * z));
    oz = fmaf(M[6], x, fmaf(M[7], y, M[8] * z));
}

// eval_spectrum(uv) * scale and the sampling density of the patch containing uv
UIVR_DEV void env_lookup(const Params& P, float tu, float tv, float le[3], float& pdf_uv) {
    const int W = P.env_w, H = P.env_h;
    const float fx = tu * (float) W, fy = tv * (float) (H - 1);
    int px = (int) fx, py = (int) fy;
    px = px > W - 1 ? W - 1 : (px < 0 ? 0 : px);
    py = py > H - 2 ? H - 2 : (py < 0 ? 0 : py);
    const float w1x = fx - (float) px, w1y = fy - (float) py, w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    const float4* row0 = P.env_data + (size_t) py * (W + 1) + px;
    const float4 v00 = __ldg(row0), v10 = __ldg(row0 + 1), v01 = __ldg(row0 + (W + 1)), v11 = __ldg(row0 + (W + 2));
    le[0] = fmaf(w0y, fmaf(w0x, v00.x, w1x * v10.x), w1y * fmaf(w0x, v01.x, w1x * v11.x)) * P.env_scale;
    le[1] = fmaf(w0y, fmaf(w0x, v00.y, w1x * v10.y), w1y * fmaf(w0x, v01.y, w1x * v11.y)) * P.env_scale;
    le[2] = fmaf(w0y, fmaf(w0x, v00.z, w1x * v10.z), w1y * fmaf(w0x, v01.z, w1x * v11.z)) * P.env_scale;
    pdf_uv = v00.w;
}

UIVR_DEV float env_inv_sin_theta(float dx, float dz) {
    const float s2 = fmaf(dx, dx, dz * dz);
    return 1.0f / sqrtf(s2 > UIVR_ENV_EPS2 ? s2 : UIVR_ENV_EPS2);
}

// Emitter::eval(si) and Emitter::pdf_direction for a ray leaving along the local direction (lx, ly, lz)
UIVR_DEV void env_eval(const Params& P, float lx, float ly, float lz, float le[3], float& pdf_dir) {
    float wx, wy, wz, dx, dy, dz;
    mat3_apply(P.local_to_world, lx, ly, lz, wx, wy, wz);
    mat3_apply(P.world_to_env, wx, wy, wz, dx, dy, dz);
    const float u = atan2_turns(dx, -dz);
    const float s2 = fmaf(-dy, dy, 1.0f);
    const float sy = sqrtf(s2 > 0.0f ? s2 : 0.0f);
    float tv = 2.0f * atan2_turns(sy, dy);
    float tu = u - 0.5f / (float) P.env_w;
    tu -= floorf(tu);
    tv = tv < 0.0f ? 0.0f : (tv > 1.0f ? 1.0f : tv);
    float pdf_uv;
    env_lookup(P, tu, tv, le, pdf_uv);
    pdf_dir = (pdf_uv * env_inv_sin_theta(dx, dz)) * UIVR_INV_2PI2;
}

// smallest i in [0, n) with cdf[i] > x (clamped to n - 1) and the position of x inside that bin
UIVR_DEV int cdf_find(const float* __restrict__ cdf, int n, float x, float& frac) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) > x) hi = mid; else lo = mid + 1;
    }
    const float a = lo > 0 ? __ldg(cdf + lo - 1) : 0.0f, den = __ldg(cdf + lo) - a;
    const float f = den > 0.0f ? (x - a) / den : 0.5f;
    frac = f < 0.0f ? 0.0f : (f > UIVR_ONE_MINUS_EPS ? UIVR_ONE_MINUS_EPS : f);
    return lo;
}

// Scene::sample_emitter_direction: world direction, solid-angle pdf, radiance
UIVR_DEV void env_sample(const Params& P, float xi1, float xi2, float& wx, float& wy, float& wz, float& pdf_dir,
                         float le[3]) {
    const int W = P.env_w, H = P.env_h;
    float f1, f2;
    const int r = cdf_find(P.env_marg, H - 1, xi2, f2);
    const int c = cdf_find(P.env_cond + (size_t) r * W, W, xi1, f1);
    const float tu = ((float) c + f1) / (float) W, tv = ((float) r + f2) / (float) (H - 1);
    float pdf_uv;
    env_lookup(P, tu, tv, le, pdf_uv);
    pdf_uv = __ldg(P.env_data + (size_t) r * (W + 1) + c).w;
    float du = tu + 0.5f / (float) W;
    du -= floorf(du);
    float sp, cp, st, ct;
    sincos2pi(du, sp, cp);
    sincos2pi(0.5f * tv, st, ct);
    const float dx = sp * st, dy = ct, dz = -(cp * st);
    pdf_dir = (pdf_uv * env_inv_sin_theta(dx, dz)) * UIVR_INV_2PI2;
    mat3_apply(P.env_to_world, dx, dy, dz, wx, wy, wz);
}

// mi.ad.common.mis_weight: power heuristic
UIVR_DEV float mis_power(float a, float b) {
    const float a2 = a * a;
    return a > 0.0f ? a2 / fmaf(b, b, a2) : 0.0f;
}

__global__ void k_test_atan2_turns(const float* y, const float* x, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = atan2_turns(y[i], x[i]);
}

}  // namespace uivr
