// uivr_path.cuh -- per-sample path logic (variant 1: one sample per lane, run to completion).
//
// Follows python/integrators/volpathsimple.py; every function cites the lines it replaces.
// The persistent slot-pool kernels (uivr_pool.cuh) re-express the same logic as a state machine
// and must produce identical per-sample results.
#pragma once

#include "uivr_device.cuh"
#include "uivr_env.cuh"

namespace uivr {

UIVR_DEV float mean3(const float v[3]) { return ((v[0] + v[1]) + v[2]) * (1.0f / 3.0f); }

// estimate_transmittance: ratio tracking (volpathsimple.py:436-504).  ADJ: per tentative
// collision with tr > 0, d sigma_t(p) += -sum(adjoint)/sigma_n  (:483-492).
template <bool ADJ, bool COUNT>
UIVR_DEV float ratio_track(const Params& P, const Seg& s, Rng& rng, float asum, Counters<COUNT>& K) {
    Walk w;
    walk_init<COUNT>(P, s, w, K);
    float T = 1.0f;
    for (;;) {
        if (!walk_next<COUNT>(P, s, w, draw(rng, K), K)) break;
        const float px = fmaf(w.t, s.dx, s.ox), py = fmaf(w.t, s.dy, s.oy), pz = fmaf(w.t, s.dz, s.oz);
        const float st = sigma_tap(P, px, py, pz);
        K.add(C_SIGMA, 1);
        const float sn = w.sb - st;
        const float tr = sn / w.sb;
        if (ADJ && tr > 0.0f) {
            scatter_sigma(P, px, py, pz, -asum / sn);
            K.add(C_SSCAT, 1);
        }
        T *= tr;
        if (T == 0.0f) break;
    }
    return T;
}

// sample_emitter_for_nee + sample_emitter (volpathsimple.py:380-433): constant emitter,
// isotropic phase => contribution = beta * (0.5 Le) * T.  ADJ replays the walk from a cloned
// sampler with adjoint dL * contribution (:393-401).
template <bool ADJ, bool COUNT>
UIVR_DEV void nee(const Params& P, float px, float py, float pz, const float beta[3], Rng& rng,
                  const float dL[3], float contrib[3], Counters<COUNT>& K) {
    const float xi1 = draw(rng, K), xi2 = draw(rng, K);
    float wx, wy, wz, wgt[3];
    bool worked = true;
    if (P.env_data) {
        // envmap: throughput * phase_val * mis_weight(ds.pdf, phase_pdf) * (Le / ds.pdf) * T  (:385-391, :419-423)
        float pdf, le[3];
        env_sample(P, xi1, xi2, wx, wy, wz, pdf, le);
        worked = pdf != 0.0f;  // sampling_worked (:421-423): no shadow ray, no draws
        const float mis = mis_power(pdf, UIVR_INV_4PI);
#pragma unroll
        for (int c = 0; c < 3; ++c) wgt[c] = worked ? ((beta[c] * UIVR_INV_4PI) * mis) * (le[c] / pdf) : 0.0f;
    } else {
        uniform_sphere(xi1, xi2, wx, wy, wz);
#pragma unroll
        for (int c = 0; c < 3; ++c) wgt[c] = beta[c] * P.half_le[c];
    }
    Seg s;
    const bool valid = worked && make_segment(P, px, py, pz, wx, wy, wz, s);
    Rng clone = rng;
    const float T = valid ? ratio_track<false, COUNT>(P, s, rng, 0.0f, K) : 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) contrib[c] = wgt[c] * T;
    if (ADJ && valid) {
        const float asum = (dL[0] * contrib[0] + dL[1] * contrib[1]) + dL[2] * contrib[2];
        ratio_track<true, COUNT>(P, s, clone, asum, K);
    }
}

// Medium::sample_interaction_drt
template <bool COUNT>
UIVR_DEV bool drt_sample(const Params& P, const Seg& s, Rng& alt, float& t_sel, float& st_sel, float& D,
                         Counters<COUNT>& K) {
    Walk w;
    walk_init<COUNT>(P, s, w, K);
    float T = 1.0f;
    D = 0.0f;
    bool found = false;
    for (;;) {
        if (!walk_next<COUNT>(P, s, w, draw(alt, K), K)) break;
        const float u2 = draw(alt, K);
        const float st = sigma_tap(P, fmaf(w.t, s.dx, s.ox), fmaf(w.t, s.dy, s.oy), fmaf(w.t, s.dz, s.oz));
        K.add(C_SIGMA, 1);
        const float wi = T / w.sb;
        D += wi;
        if (u2 <= wi / D) {
            t_sel = w.t;
            st_sel = st;
            found = true;
        }
        T *= (w.sb - st) / w.sb;
        if (!(T > 0.0f)) break;
    }
    return found;
}

struct Reservoir {
    float wsum[3], wcur[3];
    Seg seg;
    int depth;
    bool valid;
};

template <bool COUNT>
UIVR_DEV void drt_backprop(const Params& P, const Seg& s, int depth, const float adjoint[3], Rng& alt,
                           Counters<COUNT>& K);

// VolpathSimpleIntegrator.sample main loop (volpathsimple.py:110-285)
template <bool ADJ, bool COUNT>
UIVR_DEV void path_loop(const Params& P, Rng& rng, Rng& alt, Seg seg, int depth, bool active, bool escaped,
                        bool has_scattered, const float dL[3], float R[3], Counters<COUNT>& K) {
    float beta[3] = {1.0f, 1.0f, 1.0f};
    Reservoir rsv;
    const bool use_rsv = ADJ && P.use_drt && P.use_drt_subsampling;
    if (ADJ) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rsv.wsum[c] = rsv.wcur[c] = 0.0f;
        rsv.valid = false;
        rsv.depth = 0;
    }

    while (active) {
        draw(rng, K);  // :120 Russian-roulette draw (never used: rr_depth > max_depth)
        if (beta[0] == 0.0f && beta[1] == 0.0f && beta[2] == 0.0f) break;

        // :126 sample_real_interaction (:323-377): analog delta tracking
        Walk w;
        walk_init<COUNT>(P, seg, w, K);
        bool did_scatter = false;
        float sigma_t = 0.0f, px = 0.0f, py = 0.0f, pz = 0.0f;
        for (;;) {
            if (!walk_next<COUNT>(P, seg, w, draw(rng, K), K)) break;
            px = fmaf(w.t, seg.dx, seg.ox); py = fmaf(w.t, seg.dy, seg.oy); pz = fmaf(w.t, seg.dz, seg.oz);
            const float st = sigma_tap(P, px, py, pz);
            K.add(C_SIGMA, 1);
            const float r = st / w.sb;
            if (draw(rng, K) >= r) continue;
            did_scatter = true;
            sigma_t = st;
            break;
        }
        const bool did_escape = !did_scatter;
        const float t_real = w.t;
        if (did_scatter) {
            has_scattered = true;
            K.add(C_REAL, 1);
        }

        float albedo[3] = {1.0f, 1.0f, 1.0f};
        if (did_scatter) {
            albedo_tap(P, px, py, pz, albedo);
            K.add(C_ALBEDO, 1);
        }

        if (ADJ) {
            if (P.use_drt) {
                if (use_rsv) {
                    // :521-539 + DRTReservoir.update (:745-753)
                    const float u = draw(alt, K);
                    float ratio[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        rsv.wsum[c] += beta[c];
                        ratio[c] = beta[c] / rsv.wsum[c];
                    }
                    if (u <= mean3(ratio)) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) rsv.wcur[c] = beta[c];
                        rsv.seg = seg;
                        rsv.depth = depth;
                        rsv.valid = true;
                    }
                } else {
                    float adj[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) adj[c] = dL[c] * beta[c];
                    drt_backprop<COUNT>(P, seg, depth, adj, alt, K);
                }
            }
            // :152-172 free-flight scattering gradient
            if ((!P.use_drt || P.use_drt_mis) && did_scatter) {
                float m = 1.0f;
                if (P.use_drt && P.use_drt_mis) {
                    const float s2 = sigma_t * sigma_t;
                    m = s2 / (1.0f + s2);
                }
                const float inv_pdf = 1.0f / sigma_t;
                float gs = 0.0f, ga[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float Li = R[c] / (albedo[c] > 1e-8f ? albedo[c] : 1e-8f);
                    const float term = ((m * dL[c]) * Li) * inv_pdf;
                    gs = fmaf(term, albedo[c], gs);
                    ga[c] = term * sigma_t;
                }
                scatter_sigma(P, px, py, pz, gs);
                scatter_albedo(P, px, py, pz, ga);
                K.add(C_SSCAT, 1);
                K.add(C_ASCAT, 1);
            }
            // :181-189, :584-607 transmittance gradient
            {
                const float interval = did_escape ? seg.tmax : t_real;
                const float aw = fmaf(dL[2], R[2], fmaf(dL[1], R[1], dL[0] * R[0]));
                const float g = -(aw * (interval * 0.25f));
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const float tk = draw(alt, K) * interval;
                    scatter_sigma(P, fmaf(tk, seg.dx, seg.ox), fmaf(tk, seg.dy, seg.oy), fmaf(tk, seg.dz, seg.oz), g);
                    K.add(C_SSCAT, 1);
                }
            }
        }

        // :193-200
#pragma unroll
        for (int c = 0; c < 3; ++c) beta[c] *= albedo[c];
        if (did_scatter) depth += 1;
        active = did_scatter && (depth < P.max_depth);

        // :206-215 emitter sampling
        if (P.use_nee && did_scatter && active) {
            float contrib[3];
            nee<ADJ, COUNT>(P, px, py, pz, beta, rng, dL, contrib, K);
#pragma unroll
            for (int c = 0; c < 3; ++c) R[c] = ADJ ? R[c] - contrib[c] : R[c] + contrib[c];
        }

        // :221-235 phase sampling
        if (did_scatter) {
            draw(rng, K);
            const float xi1 = draw(rng, K), xi2 = draw(rng, K);
            float wx, wy, wz;
            uniform_sphere(xi1, xi2, wx, wy, wz);
            if (!make_segment(P, px, py, pz, wx, wy, wz, seg)) active = false;  // :240-241
        }
        if (did_escape) escaped = true;  // :244-245
    }

    // :249-259 DRT on the reservoir's segment
    if (ADJ && use_rsv && rsv.valid) {
        const float d = mean3(rsv.wcur), ws = mean3(rsv.wsum);
        float adj[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float W = (d != 0.0f) ? (ws * rsv.wcur[c]) / d : 0.0f;
            adj[c] = W * dL[c];
        }
        drt_backprop<COUNT>(P, rsv.seg, rsv.depth, adj, alt, K);
    }

    // :263-285 envmap (primal only)
    if (!ADJ && escaped && !(depth <= 0 && P.hide_emitters)) {
        if (P.env_data) {
            // :270-285 emitter.eval(si), hit_mis_weight = mis_weight(last_scatter_direction_pdf,
            // has_scattered ? emitter.pdf_direction : 0)
            float le[3], pdf;
            env_eval(P, seg.dx, seg.dy, seg.dz, le, pdf);
            float wmis = 1.0f;
            if (P.use_nee) wmis = mis_power(has_scattered ? UIVR_INV_4PI : 1.0f, has_scattered ? pdf : 0.0f);
#pragma unroll
            for (int c = 0; c < 3; ++c) R[c] = R[c] + (beta[c] * wmis) * le[c];
        } else {
            const float wmis = (P.use_nee && has_scattered) ? 0.5f : 1.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) R[c] = fmaf(beta[c] * wmis, P.radiance[c], R[c]);
        }
    }
}

// backpropagate_scattering_drt (:543-581) + sample_recursive (:610-655)
template <bool COUNT>
UIVR_DEV void drt_backprop(const Params& P, const Seg& s, int depth, const float adjoint[3], Rng& alt,
                           Counters<COUNT>& K) {
    float t_sel = 0.0f, st = 0.0f, D = 0.0f;
    if (!drt_sample<COUNT>(P, s, alt, t_sel, st, D, K)) return;
    const float px = fmaf(t_sel, s.dx, s.ox), py = fmaf(t_sel, s.dy, s.oy), pz = fmaf(t_sel, s.dz, s.oz);
    float albedo[3];
    albedo_tap(P, px, py, pz, albedo);
    K.add(C_ALBEDO, 1);

    float Li[3] = {0.0f, 0.0f, 0.0f};
    const float one[3] = {1.0f, 1.0f, 1.0f};
    if (P.use_nee) nee<false, COUNT>(P, px, py, pz, one, alt, nullptr, Li, K);
    draw(alt, K);
    const float xi1 = draw(alt, K), xi2 = draw(alt, K);
    float wx, wy, wz;
    uniform_sphere(xi1, xi2, wx, wy, wz);
    Seg rs;
    const bool ok = make_segment(P, px, py, pz, wx, wy, wz, rs);
    const bool active = ok && (depth + 1 < P.max_depth);
    if (active) {
        draw(alt, K);  // :99 alt_seed_rnd of the recursive sample()
        float Lr[3] = {0.0f, 0.0f, 0.0f};
        path_loop<false, COUNT>(P, alt, alt, rs, depth + 1, true, false, true, nullptr, Lr, K);
#pragma unroll
        for (int c = 0; c < 3; ++c) Li[c] += Lr[c];
    }

    const float m = P.use_drt_mis ? 1.0f / (1.0f + st * st) : 1.0f;
    float gs = 0.0f, ga[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float term = ((m * D) * adjoint[c]) * Li[c];
        gs = fmaf(term, albedo[c], gs);
        ga[c] = term * st;
    }
    scatter_sigma(P, px, py, pz, gs);
    scatter_albedo(P, px, py, pz, ga);
    K.add(C_SSCAT, 1);
    K.add(C_ASCAT, 1);
}

// one full sample() call from the camera (ray generation + reach_medium + loop)
template <bool ADJ, bool COUNT>
UIVR_DEV void sample_from_camera(const Params& P, uint32_t idx, const float dL[3], float R[3], Counters<COUNT>& K) {
    Rng rng, alt;
    rng.seed_sampler(P.seed, idx);
    if (ADJ) alt.seed_sampler(P.alt_seed, idx);
    else alt = rng;
    const float jx = draw(rng, K), jy = draw(rng, K);
    Seg seg;
    const int status = camera_segment(P, idx / P.spp, jx, jy, seg);
    draw(rng, K);  // :71
    const bool active = status == 1;
    if (active) {
        draw(rng, K);  // :99
        if (!ADJ) K.add(C_HITS, 1);
    }
    path_loop<ADJ, COUNT>(P, rng, alt, seg, 0, active, status == 0, false, dL, R, K);
}

// local pixel slot -> global pixel (uivr_shard); returns false for padding slots
UIVR_DEV bool slot_to_pixel(const Params& P, uint32_t slot, uint32_t& pix) {
    if (P.shard_count <= 1) {
        pix = slot;
    } else {
        const uint32_t b = (uint32_t) P.shard_block;
        pix = ((slot / b) * (uint32_t) P.shard_count + (uint32_t) P.shard_rank) * b + slot % b;
    }
    return pix < P.npix;
}

}  // namespace uivr
