// uivr_nerf.cuh -- the `nerf` integrator (python/integrators/nerf.py): emission-absorption ray
// marching over the sigma_t grid (corner-octet layout) and an RGB emission grid (Z,Y,X,3).
//
// Every ray that enters the medium takes exactly `queries_per_ray` forward-looking steps, so one
// sample per lane has no divergence to compact away (unlike volpathsimple): warps pull 32-sample
// chunks from a global counter, consecutive samples belong to the same or the neighbouring pixel,
// and their taps share sectors in L1/L2.  Forward: 1 octet tap (32 B) + 1 emission tap (8 x 12 B)
// per step; the adjoint replays the steps with path replay (nerf.py:109-124) and adds one 8-voxel
// sigma_t scatter + one 8x3 emission scatter (red.global.add.f32) per step.  HBM/L2-bound gather /
// scatter work; arithmetic contract as everywhere (explicit fmaf, exact-op exp shared with the oracle).
#pragma once

#include "uivr_kernels.cuh"
#include "uivr_env.cuh"

namespace uivr {

// exp(x): n = rint(x log2 e), r = x - n ln2 (two parts), Cephes expf polynomial, 2^n by exponent bits
UIVR_DEV float exp_exact(float x) {
    x = x < -87.0f ? -87.0f : (x > 88.0f ? 88.0f : x);
    const float n = rintf(x * 0x1.715476p+0f);
    float r = fmaf(n, -0x1.62e400p-1f, x);
    r = fmaf(n, -0x1.7f7d1cp-20f, r);
    float q = 1.9875691500e-4f;
    q = fmaf(q, r, 1.3981999507e-3f);
    q = fmaf(q, r, 8.3334519073e-3f);
    q = fmaf(q, r, 4.1665795894e-2f);
    q = fmaf(q, r, 1.6666665459e-1f);
    q = fmaf(q, r, 5.0000001201e-1f);
    const float e = fmaf(q, r * r, r) + 1.0f;
    return e * __uint_as_float((uint32_t) ((int) n + 127) << 23);
}

// NeRFIntegrator.sample (nerf.py:47-147).  ADJ: R enters as the primal radiance (state_in).
template <bool ADJ, bool COUNT>
UIVR_DEV void nerf_sample(const Params& P, uint32_t pix, uint32_t idx, const float* dL, float R[3], Counters<COUNT>& K) {
    Rng rng;
    rng.seed_sampler(P.seed, idx);
    Seg seg;
    int status;
    if (P.sensors) {
        status = batch_segment(P, pix, idx, seg);
    } else {
        const float jx = draw(rng, K), jy = draw(rng, K);
        status = camera_segment(P, pix, jx, jy, seg);
    }
    const bool active = (status == 1), escaped = (status == 0);  // :69-77
    float wsum = 0.0f, T = 1.0f;
    if (active) {
        if (!ADJ) K.add(C_HITS, 1);
        const int Q = P.nerf_queries;
        const float step = P.nerf_jitter ? seg.tmax / (float) Q : seg.tmax / (float) (Q - 1);  // :6-10
        const float jit = draw(rng, K);                                                        // :87
        float t_a = 0.0f;
#pragma unroll 2
        for (int j = 0; j < Q; ++j) {
            const float sj = (float) (j + 1);
            const float t_b = P.nerf_jitter ? step * (sj + jit) : step * sj;  // :12-18, mint = 0
            const float dt = t_b - t_a;
            const float px = fmaf(t_b, seg.dx, seg.ox), py = fmaf(t_b, seg.dy, seg.oy), pz = fmaf(t_b, seg.dz, seg.oz);
            const float raw = sigma_tap(P, px, py, pz);
            const bool clipped = P.nerf_activation == 1 && !(raw > 0.0f);  // relu (:38-45)
            const float sigma = clipped ? 0.0f : raw;
            const bool last = !(j + 1 < Q);
            const float a = last ? 1.0f : exp_exact(-(sigma * dt));  // :103-105
            const float weight = (1.0f - a) * T;
            const float safe = a + 1e-10f;
            K.add(C_SIGMA, 1);
            K.add(C_ALBEDO, 1);
            float em[3] = {0.0f, 0.0f, 0.0f};
            // a zero weight makes the emission irrelevant to the primal sum (R + 0 == R)
            if (ADJ || weight != 0.0f) albedo_tap(P, px, py, pz, em);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float we = weight * em[c];
                R[c] = ADJ ? R[c] - we : R[c] + we;  // :109-112
            }
            if (ADJ && !last) {
                // :117-124  d/d emission_c = dL_c weight;  d/d sigma = sum_c dL_c (em_c dt a T - (R_c/safe) dt a)
                const float da = dt * a;
                float gs = 0.0f, ge[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    ge[c] = dL[c] * weight;
                    const float inner = fmaf(em[c], da * T, -((R[c] / safe) * da));
                    gs = fmaf(dL[c], inner, gs);
                }
                if (clipped) gs = 0.0f;
                K.add(C_SSCAT, 1);
                K.add(C_ASCAT, 1);
                if (gs != 0.0f) scatter_sigma(P, px, py, pz, gs);
                if (ge[0] != 0.0f || ge[1] != 0.0f || ge[2] != 0.0f) scatter_albedo(P, px, py, pz, ge);
            }
            t_a = t_b;
            if (!last) {  // :114-120, masked by the updated still_walking
                T *= safe;
                wsum += weight;
            }
        }
    }
    // :134-143 composite with the background emitter (both modes)
    bool active_e = escaped || active;
    if (P.hide_emitters) active_e = active_e && (wsum > 0.0f);
    if (active_e) {
        float le[3] = {P.radiance[0], P.radiance[1], P.radiance[2]}, pdf;
        if (P.env_data) env_eval(P, seg.dx, seg.dy, seg.dz, le, pdf);
#pragma unroll
        for (int c = 0; c < 3; ++c) R[c] += (1.0f - wsum) * le[c];
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_nerf_forward(const Params P) {
    Counters<COUNT> K;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    uint32_t item;
    while (next_chunk(P, total, item)) {
        uint32_t pix = 0;
        const bool live = (uint64_t) item < total && slot_to_pixel(P, item / P.spp, pix);
        float L[3] = {0.0f, 0.0f, 0.0f};
        if (live) {
            const uint32_t idx = pix * P.spp + item % P.spp;
            nerf_sample<false, COUNT>(P, pix, idx, nullptr, L, K);
            K.add(C_SAMPLES, 1);
            if (P.sample_L) {
                P.sample_L[3 * (size_t) idx + 0] = L[0];
                P.sample_L[3 * (size_t) idx + 1] = L[1];
                P.sample_L[3 * (size_t) idx + 2] = L[2];
            }
        }
        __syncwarp();
        if ((P.spp & 31u) == 0) {  // the whole chunk belongs to one pixel
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = L[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0 && live) atomicAdd(P.image + 3 * (size_t) pix + c, v);
            }
        } else if (live) {
            atomicAdd(P.image + 3 * (size_t) pix + 0, L[0]);
            atomicAdd(P.image + 3 * (size_t) pix + 1, L[1]);
            atomicAdd(P.image + 3 * (size_t) pix + 2, L[2]);
        }
    }
    K.flush(P.counters);
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_nerf_backward(const Params P) {
    Counters<COUNT> K;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    uint32_t item;
    while (next_chunk(P, total, item)) {
        uint32_t pix = 0;
        const bool live = (uint64_t) item < total && slot_to_pixel(P, item / P.spp, pix);
        if (live) {
            const uint32_t idx = pix * P.spp + item % P.spp;
            float L[3] = {0.0f, 0.0f, 0.0f};
            nerf_sample<false, COUNT>(P, pix, idx, nullptr, L, K);  // detached primal pass -> state_in
            K.add(C_SAMPLES, 1);
            if (P.sample_L) {
                P.sample_L[3 * (size_t) idx + 0] = L[0];
                P.sample_L[3 * (size_t) idx + 1] = L[1];
                P.sample_L[3 * (size_t) idx + 2] = L[2];
            }
            float dL[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) dL[c] = __ldg(P.grad_image + 3 * (size_t) pix + c) * P.inv_spp;
            nerf_sample<true, COUNT>(P, pix, idx, dL, L, K);
        }
        __syncwarp();
    }
    K.flush(P.counters);
}

__global__ void k_test_exp(const float* x, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp_exact(x[i]);
}

}  // namespace uivr
