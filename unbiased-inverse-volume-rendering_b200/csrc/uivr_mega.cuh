// uivr_mega.cuh -- variant 0: persistent lane-refill megakernel (sm_100a).
//
// One CTA set per SM stays resident; every LANE owns one sample at a time and is refilled
// from a global counter the moment its sample finishes, so warps stay full no matter how
// path lengths vary.  The per-sample logic of uivr_path.cuh (= volpathsimple.py) is
// re-expressed as a per-lane state machine with FIVE states:
//
//   S_WALK     free-flight walk over the majorant supergrid (Medium::sample_interaction and
//              its ratio-tracking / DRT siblings).  This is where >80 % of the time goes, so
//              it is one fused handler: a branch-free DDA cell step executed by all walking
//              lanes, with the tentative collisions (sigma_t tap + accept / ratio / DRT
//              update) deferred until kTapBatch lanes are waiting, so that the tap code and
//              its memory latency are paid once per batch instead of once per lane.
//   S_VERTEX   end of a delta-tracking segment: albedo, reservoir, gradient estimators,
//              then straight on to emitter sampling or phase sampling
//   S_NEE_END  end of a shadow walk: contribution, adjoint replay, phase sampling
//   S_DRT_END  end of the DRT walk: set up the DRT vertex
//   S_PATH_END end of a path: envmap, pass switch (primal -> adjoint -> DRT -> recursive)
//   S_FETCH    ray generation: new sample from the global queue, or the adjoint re-start
//
// The scheduler is a fixed sweep WALK -> DRT_END -> VERTEX -> NEE_END -> PATH_END -> FETCH (a
// lane can pass through several handlers per sweep and be walking again at the end of it);
// WALK runs until half of its lanes have left.  Results are per-sample identical to variant 1 /
// the oracle because RNG streams are keyed by the global sample index and every handler keeps
// the operation order of the oracle.
//
// passes : PRIMAL (radiance) -> [backward only] ADJ (path replay adjoint) -> DRT walk ->
//          DRTV (DRT vertex: NEE + phase sample) -> REC (detached recursive path for Li)
// The O(n^2) `use_drt_subsampling=False` mode keeps its sub-paths on a stack and is served by
// variant 1.
#pragma once

#include "uivr_kernels.cuh"

namespace uivr {

enum : int { S_WALK = 0, S_VERTEX, S_NEE_END, S_DRT_END, S_PATH_END, S_FETCH, S_IDLE };
enum : int { M_DELTA = 0, M_NEE, M_NEE_ADJ, M_DRT };
enum : int { P_PRIMAL = 0, P_ADJ, P_DRTV, P_REC };

constexpr int kMegaBlock = 256;
constexpr int kTapBatch = 8;  // tentative collisions are evaluated when this many lanes wait

template <bool BWD, bool COUNT>
__global__ void __launch_bounds__(kMegaBlock, 2) k_mega(const Params P) {
    Counters<COUNT> K;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    const bool use_rsv = BWD && P.use_drt && P.use_drt_subsampling;

    // ---- per-lane state ----
    int state = S_FETCH, mode = M_DELTA, pass = P_PRIMAL;
    bool need_init = false;    // S_WALK: the walk over `seg` has not been set up yet
    bool restart = false;      // S_FETCH: re-generate the camera ray of `idx` for the adjoint pass
    uint32_t idx = 0, pix = 0;
    Rng rng, alt;              // rng = stream the current walk draws from; alt = adjoint side stream
    rng.state = rng.inc = alt.state = alt.inc = 0;
    uint64_t clone_state = 0;  // sampler.clone() before the NEE walk (:383)
    float ox = 0.0f, oy = 0.0f, oz = 0.0f, dx = 0.0f, dy = 0.0f, dz = 0.0f, tmax = 0.0f;  // segment
    float wt = 0.0f, tnx = 0.0f, tny = 0.0f, tnz = 0.0f, sb = 0.0f, tau = 0.0f;            // walk
    float adx = 0.0f, ady = 0.0f, adz = 0.0f;                                              // |cell / d|
    int cx = 0, cy = 0, cz = 0;
    float vpx = 0.0f, vpy = 0.0f, vpz = 0.0f;  // current vertex (local space)
    float beta[3] = {1.0f, 1.0f, 1.0f}, R[3] = {0.0f, 0.0f, 0.0f}, dL[3] = {0.0f, 0.0f, 0.0f};
    float sigma_t = 0.0f;      // sigma_t at the real collision / DRT vertex
    float T = 1.0f;            // ratio-tracking transmittance of the current side walk
    float asum = 0.0f;         // sum_c adjoint_c of the NEE replay
    int depth = 0;
    bool did_scatter = false, escaped = false, has_scattered = false, active = false, nee_valid = false;
    // adjoint-only state (dead code in the forward instance)
    float rs_wsum[3] = {0.0f, 0.0f, 0.0f}, rs_wcur[3] = {0.0f, 0.0f, 0.0f};
    float rs_ox = 0.0f, rs_oy = 0.0f, rs_oz = 0.0f, rs_dx = 0.0f, rs_dy = 0.0f, rs_dz = 0.0f, rs_tmax = 0.0f;
    int rs_depth = 0;
    bool rs_valid = false;
    float drt_D = 0.0f, drt_t = 0.0f, drt_st = 0.0f;
    bool drt_found = false;
    float aux_Li[3] = {0.0f, 0.0f, 0.0f}, aux_alb[3] = {1.0f, 1.0f, 1.0f};

    bool queue_empty = false;

// sample_emitter (:406-433): emitter direction + shadow segment from the vertex
#define UIVR_NEE_START()                                                                           \
    do {                                                                                           \
        const float xi1_ = draw(rng, K), xi2_ = draw(rng, K);                                      \
        float wx_, wy_, wz_;                                                                       \
        uniform_sphere(xi1_, xi2_, wx_, wy_, wz_);                                                 \
        Seg s_;                                                                                    \
        nee_valid = make_segment(P, vpx, vpy, vpz, wx_, wy_, wz_, s_);                             \
        ox = s_.ox; oy = s_.oy; oz = s_.oz; dx = s_.dx; dy = s_.dy; dz = s_.dz; tmax = s_.tmax;    \
        clone_state = rng.state; /* sampler.clone() position for the adjoint replay (:383) */     \
        T = nee_valid ? 1.0f : 0.0f;                                                               \
        if (nee_valid) { mode = M_NEE; need_init = true; state = S_WALK; }                         \
        else state = S_NEE_END;                                                                    \
    } while (0)

// phase sampling (:221-245) / DRT-vertex continuation (:626-652)
#define UIVR_PHASE()                                                                               \
    do {                                                                                           \
        draw(rng, K);                                                                              \
        const float xi1_ = draw(rng, K), xi2_ = draw(rng, K);                                      \
        float wx_, wy_, wz_;                                                                       \
        uniform_sphere(xi1_, xi2_, wx_, wy_, wz_);                                                 \
        Seg s_;                                                                                    \
        const bool ok_ = make_segment(P, vpx, vpy, vpz, wx_, wy_, wz_, s_);                        \
        ox = s_.ox; oy = s_.oy; oz = s_.oz; dx = s_.dx; dy = s_.dy; dz = s_.dz; tmax = s_.tmax;    \
        mode = M_DELTA;                                                                            \
        if (BWD && pass == P_DRTV) {                                                               \
            depth += 1;                                                                            \
            active = ok_ && (depth < P.max_depth);                                                 \
            pass = P_REC;                                                                          \
            beta[0] = beta[1] = beta[2] = 1.0f;                                                    \
            R[0] = R[1] = R[2] = 0.0f;                                                             \
            escaped = false;                                                                       \
            has_scattered = true;                                                                  \
            if (active) draw(rng, K); /* :99 of the recursive sample() */                          \
        } else if (!ok_) {                                                                         \
            active = false; /* :240-241 accidental escape */                                       \
        }                                                                                          \
        if (active) { need_init = true; state = S_WALK; }                                          \
        else state = S_PATH_END;                                                                   \
    } while (0)

    for (;;) {
        // ==============================================================================
        // S_WALK
        // ==============================================================================
        {
            const unsigned m0 = __ballot_sync(FULL, state == S_WALK);
            if (m0) {
                // ---- set up freshly started walks (loop top :114-121 for the main path) ----
                if (state == S_WALK && need_init) {
                    need_init = false;
                    bool go = true;
                    if (mode == M_DELTA) {
                        draw(rng, K);  // :120 Russian-roulette draw
                        if (beta[0] == 0.0f && beta[1] == 0.0f && beta[2] == 0.0f) {
                            active = false;
                            state = S_PATH_END;
                            go = false;
                        }
                    }
                    if (go) {
                        // walk_init (Medium::sample_interaction set-up, App. B.5)
                        const float ix = dx != 0.0f ? 1.0f / dx : UIVR_INF;
                        const float iy = dy != 0.0f ? 1.0f / dy : UIVR_INF;
                        const float iz = dz != 0.0f ? 1.0f / dz : UIVR_INF;
                        wt = 0.0f;
                        walk_axis_init(ox, dx, ix, P.fmres[0], P.mcs[0], P.mres[0], cx, tnx);
                        walk_axis_init(oy, dy, iy, P.fmres[1], P.mcs[1], P.mres[1], cy, tny);
                        walk_axis_init(oz, dz, iz, P.fmres[2], P.mcs[2], P.mres[2], cz, tnz);
                        adx = fabsf(P.mcs[0] * ix);
                        ady = fabsf(P.mcs[1] * iy);
                        adz = fabsf(P.mcs[2] * iz);
                        sb = majorant_at<COUNT>(P, cx, cy, cz, K);
                        tau = neg_log1m(draw(rng, K));
                    }
                }
                const int n0 = __popc(m0);
                bool pending = false;  // a tentative collision at `wt` waits for its tap
                for (;;) {
                    // ---- one supergrid cell per iteration (branch-free DDA) ----
                    if (state == S_WALK && !pending) {
                        const bool yx = tny < tnx;
                        float tn = yx ? tny : tnx;
                        const bool zb = tnz < tn;
                        tn = zb ? tnz : tn;
                        const float t_end = tn < tmax ? tn : tmax;
                        float len = t_end - wt;
                        len = len < 0.0f ? 0.0f : len;
                        const float dtau = sb * len;
                        if (sb > 0.0f && tau < dtau) {
                            float t = wt + tau / sb;
                            wt = t > t_end ? t_end : t;
                            pending = true;
                        } else {
                            if (sb > 0.0f) tau -= dtau;
                            wt = t_end > wt ? t_end : wt;
                            bool end = !(tn < tmax);
                            const bool a0 = !zb && !yx, a1 = !zb && yx;
                            const int sx = dx > 0.0f ? 1 : -1, sy = dy > 0.0f ? 1 : -1, sz = dz > 0.0f ? 1 : -1;
                            cx += a0 ? sx : 0;
                            cy += a1 ? sy : 0;
                            cz += zb ? sz : 0;
                            tnx = a0 ? tnx + adx : tnx;
                            tny = a1 ? tny + ady : tny;
                            tnz = zb ? tnz + adz : tnz;
                            end = end || (unsigned) cx >= (unsigned) P.mres[0] || (unsigned) cy >= (unsigned) P.mres[1] ||
                                  (unsigned) cz >= (unsigned) P.mres[2];
                            if (end) {
                                // segment end: no (further) collision
                                if (mode == M_DELTA) { did_scatter = false; state = S_VERTEX; }
                                else if (mode == M_DRT) state = S_DRT_END;
                                else state = S_NEE_END;
                            } else {
                                sb = majorant_at<COUNT>(P, cx, cy, cz, K);
                            }
                        }
                    }
                    const int n_walk = __popc(__ballot_sync(FULL, state == S_WALK));
                    const int n_pend = __popc(__ballot_sync(FULL, pending));
                    const bool leave = 2 * n_walk <= n0;
                    // ---- tentative collisions: sigma_t tap + per-mode decision ----
                    if (n_pend >= kTapBatch || (n_pend > 0 && (n_pend == n_walk || leave))) {
                        if (pending) {
                            pending = false;
                            const float px = fmaf(wt, dx, ox), py = fmaf(wt, dy, oy), pz = fmaf(wt, dz, oz);
                            float u2 = 0.0f;
                            if (mode == M_DRT) u2 = draw(rng, K);
                            const float st = sigma_tap(P, px, py, pz);
                            K.add(C_SIGMA, 1);
                            bool cont = true;
                            if (mode == M_DELTA) {
                                // :354-361 real vs null collision
                                const float r = st / sb;
                                if (!(draw(rng, K) >= r)) {
                                    did_scatter = true;
                                    sigma_t = st;
                                    vpx = px; vpy = py; vpz = pz;
                                    state = S_VERTEX;
                                    cont = false;
                                }
                            } else if (mode == M_DRT) {
                                // sample_interaction_drt (App. B.6): candidate weight T/sigma_bar, size-1 reservoir
                                const float wi = T / sb;
                                drt_D += wi;
                                if (u2 <= wi / drt_D) {
                                    drt_t = wt;
                                    drt_st = st;
                                    drt_found = true;
                                }
                                T *= (sb - st) / sb;
                                if (!(T > 0.0f)) { state = S_DRT_END; cont = false; }
                            } else {
                                // ratio tracking (:461-502); M_NEE_ADJ scatters -sum(adj)/sigma_n (:483-492)
                                const float sn = sb - st;
                                const float tr = sn / sb;
                                if (BWD && mode == M_NEE_ADJ && tr > 0.0f) {
                                    scatter_sigma(P, px, py, pz, -asum / sn);
                                    K.add(C_SSCAT, 1);
                                }
                                T *= tr;
                                if (T == 0.0f) { state = S_NEE_END; cont = false; }
                            }
                            if (cont) tau = neg_log1m(draw(rng, K));
                        }
                        if (leave || __ballot_sync(FULL, state == S_WALK) == 0u) break;
                    } else if (leave) {
                        break;
                    }
                }
            }
        }

        // ==============================================================================
        // S_DRT_END: DRT walk finished (:550-558); set up the DRT vertex
        // ==============================================================================
        if (BWD && __ballot_sync(FULL, state == S_DRT_END)) {
            if (state == S_DRT_END) {
                if (drt_found) {
                    vpx = fmaf(drt_t, rs_dx, rs_ox); vpy = fmaf(drt_t, rs_dy, rs_oy); vpz = fmaf(drt_t, rs_dz, rs_oz);
                    albedo_tap(P, vpx, vpy, vpz, aux_alb);
                    K.add(C_ALBEDO, 1);
                    aux_Li[0] = aux_Li[1] = aux_Li[2] = 0.0f;
                    beta[0] = beta[1] = beta[2] = 1.0f;
                    pass = P_DRTV;
                    if (P.use_nee) UIVR_NEE_START();
                    else UIVR_PHASE();
                } else {
                    state = S_FETCH;
                }
            }
        }

        // ==============================================================================
        // S_VERTEX: end of a delta-tracking segment, real collision or escape (:130-245)
        // ==============================================================================
        if (__ballot_sync(FULL, state == S_VERTEX)) {
            if (state == S_VERTEX) {
                float albedo[3] = {1.0f, 1.0f, 1.0f};
                if (did_scatter) {
                    has_scattered = true;
                    K.add(C_REAL, 1);
                    albedo_tap(P, vpx, vpy, vpz, albedo);
                    K.add(C_ALBEDO, 1);
                }
                if (BWD && pass == P_ADJ) {
                    if (use_rsv) {
                        // DRTReservoir.update (:745-753), weight = throughput before this vertex
                        const float u = draw(alt, K);
                        float ratio[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            rs_wsum[c] += beta[c];
                            ratio[c] = beta[c] / rs_wsum[c];
                        }
                        if (u <= mean3(ratio)) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) rs_wcur[c] = beta[c];
                            rs_ox = ox; rs_oy = oy; rs_oz = oz;
                            rs_dx = dx; rs_dy = dy; rs_dz = dz;
                            rs_tmax = tmax;
                            rs_depth = depth;
                            rs_valid = true;
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
                        scatter_sigma(P, vpx, vpy, vpz, gs);
                        scatter_albedo(P, vpx, vpy, vpz, ga);
                        K.add(C_SSCAT, 1);
                        K.add(C_ASCAT, 1);
                    }
                    // :181-189, :584-607 transmittance gradient: 4 uniform taps on the segment
                    {
                        const float interval = did_scatter ? wt : tmax;
                        const float aw = fmaf(dL[2], R[2], fmaf(dL[1], R[1], dL[0] * R[0]));
                        const float g = -(aw * (interval * 0.25f));
#pragma unroll 1
                        for (int k = 0; k < 4; ++k) {
                            const float tk = draw(alt, K) * interval;
                            scatter_sigma(P, fmaf(tk, dx, ox), fmaf(tk, dy, oy), fmaf(tk, dz, oz), g);
                            K.add(C_SSCAT, 1);
                        }
                    }
                }
                // :193-200
#pragma unroll
                for (int c = 0; c < 3; ++c) beta[c] *= albedo[c];
                if (did_scatter) depth += 1;
                active = did_scatter && (depth < P.max_depth);
                if (!did_scatter) {
                    escaped = true;  // :244-245
                    state = S_PATH_END;
                } else if (P.use_nee && active) {
                    UIVR_NEE_START();
                } else {
                    UIVR_PHASE();
                }
            }
        }

        // ==============================================================================
        // S_NEE_END: sample_emitter_for_nee (:380-403): contribution, adjoint replay, phase
        // ==============================================================================
        if (__ballot_sync(FULL, state == S_NEE_END)) {
            if (state == S_NEE_END) {
                // (after the adjoint replay, M_NEE_ADJ, the gradients are scattered: just carry on)
                bool replay = false;
                if (!(BWD && mode == M_NEE_ADJ)) {
                    float contrib[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) contrib[c] = (beta[c] * P.half_le[c]) * T;
                    if (BWD && pass == P_DRTV) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) aux_Li[c] = contrib[c];
                    } else if (BWD && pass == P_ADJ) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) R[c] = R[c] - contrib[c];  // path replay (:214)
                        if (nee_valid) {
                            asum = (dL[0] * contrib[0] + dL[1] * contrib[1]) + dL[2] * contrib[2];
                            rng.state = clone_state;  // the replay consumes exactly the same draws again
                            T = 1.0f;
                            mode = M_NEE_ADJ;
                            need_init = true;
                            state = S_WALK;
                            replay = true;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) R[c] = R[c] + contrib[c];
                    }
                }
                if (!replay) UIVR_PHASE();
            }
        }

        // ==============================================================================
        // S_PATH_END
        // ==============================================================================
        if (__ballot_sync(FULL, state == S_PATH_END)) {
            if (state == S_PATH_END) {
                if (pass == P_PRIMAL || (BWD && pass == P_REC)) {
                    // :263-285 envmap
                    if (escaped && !(depth <= 0 && P.hide_emitters)) {
                        const float wmis = (P.use_nee && has_scattered) ? 0.5f : 1.0f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) R[c] = fmaf(beta[c] * wmis, P.radiance[c], R[c]);
                    }
                }
                if (pass == P_PRIMAL) {
                    if (P.sample_L) {
                        P.sample_L[3 * (size_t) idx + 0] = R[0];
                        P.sample_L[3 * (size_t) idx + 1] = R[1];
                        P.sample_L[3 * (size_t) idx + 2] = R[2];
                    }
                    if (!BWD) {
                        atomicAdd(P.image + 3 * (size_t) pix + 0, R[0]);
                        atomicAdd(P.image + 3 * (size_t) pix + 1, R[1]);
                        atomicAdd(P.image + 3 * (size_t) pix + 2, R[2]);
                        state = S_FETCH;
                    } else {
                        pass = P_ADJ;  // batched.py:309-318: sample(Backward, state_in = L)
                        restart = true;
                        state = S_FETCH;
                    }
                } else if (BWD && pass == P_ADJ) {
                    state = S_FETCH;
                    if (use_rsv && rs_valid) {
                        // DRTReservoir.get (:756-760) and adjoint = weight * dL (:255)
                        const float d = mean3(rs_wcur), ws = mean3(rs_wsum);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float W = (d != 0.0f) ? (ws * rs_wcur[c]) / d : 0.0f;
                            dL[c] = W * dL[c];
                        }
                        ox = rs_ox; oy = rs_oy; oz = rs_oz;
                        dx = rs_dx; dy = rs_dy; dz = rs_dz;
                        tmax = rs_tmax;
                        depth = rs_depth;
                        rng = alt;  // everything from here on draws from the alt stream
                        mode = M_DRT;
                        T = 1.0f;
                        drt_D = 0.0f;
                        drt_found = false;
                        need_init = true;
                        state = S_WALK;
                    }
                } else if (BWD) {  // P_REC: Li complete -> DRT gradient (:571-581)
                    const float m = P.use_drt_mis ? 1.0f / (1.0f + drt_st * drt_st) : 1.0f;
                    float gs = 0.0f, ga[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float Li = aux_Li[c] + R[c];
                        const float term = ((m * drt_D) * dL[c]) * Li;
                        gs = fmaf(term, aux_alb[c], gs);
                        ga[c] = term * drt_st;
                    }
                    const float px = fmaf(drt_t, rs_dx, rs_ox), py = fmaf(drt_t, rs_dy, rs_oy), pz = fmaf(drt_t, rs_dz, rs_oz);
                    scatter_sigma(P, px, py, pz, gs);
                    scatter_albedo(P, px, py, pz, ga);
                    K.add(C_SSCAT, 1);
                    K.add(C_ASCAT, 1);
                    state = S_FETCH;
                }
            }
        }

        // ==============================================================================
        // S_FETCH: next sample from the global queue (or adjoint re-start of the current one),
        // ray generation + reach_medium (batched.py:426-467, volpathsimple.py:292-319)
        // ==============================================================================
        {
            unsigned todo = __ballot_sync(FULL, state == S_FETCH);
            while (todo) {
                const unsigned fresh = __ballot_sync(FULL, state == S_FETCH && !restart && !queue_empty);
                if (fresh) {
                    const int leader = __ffs(fresh) - 1;
                    unsigned base = 0;
                    if ((int) lane == leader) base = atomicAdd(P.work_counter, (unsigned) __popc(fresh));
                    base = __shfl_sync(FULL, base, leader);
                    if (state == S_FETCH && !restart && !queue_empty) {
                        const uint64_t item = (uint64_t) base + __popc(fresh & ((1u << lane) - 1u));
                        if (item < total) {
                            const uint32_t it = (uint32_t) item;
                            // (padding slots of a shard stay in S_FETCH and simply try again)
                            if (slot_to_pixel(P, it / P.spp, pix)) {
                                idx = pix * P.spp + it % P.spp;
                                pass = P_PRIMAL;
                                K.add(C_SAMPLES, 1);
                                restart = true;  // generate its camera ray below
                            }
                        } else {
                            queue_empty = true;
                        }
                    }
                    if ((uint64_t) base + __popc(fresh) >= total) queue_empty = true;
                }
                if (state == S_FETCH && restart) {
                    restart = false;
                    rng.seed_sampler(P.seed, idx);
                    if (BWD && pass == P_ADJ) {
                        alt.seed_sampler(P.alt_seed, idx);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            dL[c] = __ldg(P.grad_image + 3 * (size_t) pix + c) * P.inv_spp;
                            rs_wsum[c] = 0.0f;
                            rs_wcur[c] = 0.0f;
                        }
                        rs_valid = false;
                    } else {
                        R[0] = R[1] = R[2] = 0.0f;
                    }
                    const float jx = draw(rng, K), jy = draw(rng, K);
                    Seg s;
                    const int status = camera_segment(P, pix, jx, jy, s);
                    ox = s.ox; oy = s.oy; oz = s.oz; dx = s.dx; dy = s.dy; dz = s.dz; tmax = s.tmax;
                    draw(rng, K);  // :71
                    active = status == 1;
                    escaped = status == 0;
                    has_scattered = false;
                    depth = 0;
                    beta[0] = beta[1] = beta[2] = 1.0f;
                    mode = M_DELTA;
                    if (active) {
                        draw(rng, K);  // :99 alt_seed_rnd
                        if (pass == P_PRIMAL) K.add(C_HITS, 1);
                        need_init = true;
                        state = S_WALK;
                    } else if (pass == P_PRIMAL) {
                        // the ray misses the medium: finish the sample right here
                        if (escaped && !P.hide_emitters) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) R[c] = fmaf(1.0f, P.radiance[c], 0.0f);
                        }
                        if (P.sample_L) {
                            P.sample_L[3 * (size_t) idx + 0] = R[0];
                            P.sample_L[3 * (size_t) idx + 1] = R[1];
                            P.sample_L[3 * (size_t) idx + 2] = R[2];
                        }
                        if (!BWD) {
                            atomicAdd(P.image + 3 * (size_t) pix + 0, R[0]);
                            atomicAdd(P.image + 3 * (size_t) pix + 1, R[1]);
                            atomicAdd(P.image + 3 * (size_t) pix + 2, R[2]);
                        } else if (COUNT) {
                            // the adjoint pass of a missed ray draws jitter + :71 and nothing else
                            K.add(C_DRAWS, 3);
                        }
                        // state stays S_FETCH: pick up the next sample in the next round
                    } else {
                        state = S_FETCH;  // adjoint pass of a ray that cannot enter: nothing to do
                    }
                }
                todo = __ballot_sync(FULL, state == S_FETCH && !queue_empty);
            }
        }

        // every lane is idle and the queue is empty
        if (__ballot_sync(FULL, state != S_FETCH) == 0u) break;
    }
#undef UIVR_NEE_START
#undef UIVR_PHASE
    K.flush(P.counters);
}

inline int launch_mega(int num_sms, bool backward, bool counting, const Params& P, cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e;
#define UIVR_MEGA_LAUNCH(B, C)                                                                          \
    do {                                                                                                \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mega<B, C>, kMegaBlock, 0);        \
        if (e != cudaSuccess) return -2;                                                                \
        if (per_sm < 1) per_sm = 1;                                                                     \
        k_mega<B, C><<<num_sms * per_sm, kMegaBlock, 0, st>>>(P);                                       \
    } while (0)
    if (backward) {
        if (counting) UIVR_MEGA_LAUNCH(true, true); else UIVR_MEGA_LAUNCH(true, false);
    } else {
        if (counting) UIVR_MEGA_LAUNCH(false, true); else UIVR_MEGA_LAUNCH(false, false);
    }
#undef UIVR_MEGA_LAUNCH
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace uivr
