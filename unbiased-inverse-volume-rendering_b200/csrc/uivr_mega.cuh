// uivr_mega.cuh -- variant 0: persistent lane-refill megakernel (sm_100a).
//
// One CTA set per SM stays resident; every LANE owns one sample at a time and is refilled
// from a global counter the moment its sample finishes, so warps stay full no matter how
// path lengths vary.  The per-sample logic of uivr_path.cuh (= volpathsimple.py) is
// re-expressed as a per-lane state machine; each warp iteration runs the handler of the state
// that currently holds the most lanes ("majority scheduling"), which keeps the hot handlers
// (supergrid DDA cell steps, sigma_t taps) near-converged while rarer transitions are batched
// until enough lanes wait for them.  Results are per-sample identical to variant 1 / the
// oracle because RNG streams are keyed by the global sample index.
//
// passes : PRIMAL (radiance) -> [backward only] ADJ (path replay adjoint) -> DRT walk ->
//          DRTV (DRT vertex: NEE + phase sample) -> REC (detached recursive path for Li)
// The O(n^2) `use_drt_subsampling=False` mode keeps its sub-paths on a stack and is served by
// variant 1.
#pragma once

#include "uivr_kernels.cuh"

namespace uivr {

enum : int { S_STEP = 0, S_HIT, S_VERTEX, S_NEE_START, S_NEE_DONE, S_PHASE, S_WINIT, S_PATH_END, S_DRT_DONE, S_INIT, S_DONE, S_NUM };
enum : int { M_DELTA = 0, M_NEE, M_NEE_ADJ, M_DRT };
enum : int { P_PRIMAL = 0, P_ADJ, P_DRTV, P_REC };

constexpr int kMegaBlock = 256;
constexpr int kStepBurst = 8;

template <bool BWD, bool COUNT>
__global__ void __launch_bounds__(kMegaBlock, 2) k_mega(const Params P) {
    Counters<COUNT> K;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    const bool use_rsv = BWD && P.use_drt && P.use_drt_subsampling;

    // ---- per-lane state ----
    int state = S_DONE, mode = M_DELTA, pass = P_PRIMAL;
    uint32_t idx = 0, pix = 0;
    Rng rng, alt;              // rng = stream the current walk draws from; alt = adjoint side stream
    rng.state = rng.inc = alt.state = alt.inc = 0;
    uint64_t clone_state = 0;  // sampler.clone() before the NEE walk (:383)
    Seg seg;
    seg.ox = seg.oy = seg.oz = seg.dx = seg.dy = seg.dz = seg.ix = seg.iy = seg.iz = seg.tmax = 0.0f;
    Walk w;
    w.t = w.tmax = w.tnx = w.tny = w.tnz = w.sb = 0.0f;
    w.cx = w.cy = w.cz = 0;
    float tau = 0.0f;
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

    for (;;) {
        // ---------------- majority scheduling ----------------
        int sel;
        {
            const int st_eff = (state == S_DONE && queue_empty) ? S_NUM : state;
            const unsigned peers = __match_any_sync(0xffffffffu, st_eff);
            const int key = (st_eff == S_NUM) ? 0 : ((__popc(peers) << 4) | (15 - st_eff));
            const int best = __reduce_max_sync(0xffffffffu, key);
            if (best == 0) break;  // every lane is done and the queue is empty
            sel = 15 - (best & 15);
        }

        switch (sel) {
        // ------------------------------------------------------------------------------
        case S_DONE: {  // refill idle lanes from the global sample queue
            const unsigned idle = __ballot_sync(0xffffffffu, state == S_DONE);
            unsigned base = 0;
            if (lane == (unsigned) (__ffs(idle) - 1)) base = atomicAdd(P.work_counter, (unsigned) __popc(idle));
            base = __shfl_sync(0xffffffffu, base, __ffs(idle) - 1);
            if (state == S_DONE) {
                const uint64_t item = (uint64_t) base + __popc(idle & ((1u << lane) - 1u));
                if (item < total) {
                    const uint32_t it = (uint32_t) item;
                    if (slot_to_pixel(P, it / P.spp, pix)) {
                        idx = pix * P.spp + it % P.spp;
                        pass = P_PRIMAL;
                        state = S_INIT;
                        K.add(C_SAMPLES, 1);
                    }
                } else {
                    queue_empty = true;
                }
            }
            if ((uint64_t) base + __popc(idle) >= total) queue_empty = true;
            break;
        }
        // ------------------------------------------------------------------------------
        case S_INIT: {  // ray generation + reach_medium (batched.py:426-467, volpathsimple.py:292-319)
            if (state == S_INIT) {
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
                const int status = camera_segment(P, pix, jx, jy, seg);
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
                    state = S_WINIT;
                } else {
                    state = S_PATH_END;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_WINIT: {  // start walking `seg` (loop top :114-121 for the main path)
            if (state == S_WINIT) {
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
                    walk_init<COUNT>(P, seg, w, K);
                    tau = neg_log1m(draw(rng, K));
                    state = S_STEP;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_STEP: {  // supergrid DDA: one cell per step (Medium::sample_interaction)
            const int n0 = __popc(__ballot_sync(0xffffffffu, state == S_STEP));
#pragma unroll 1
            for (int burst = 0; burst < kStepBurst; ++burst) {
                if (state == S_STEP) {
                    int ax = 0;
                    float tn = w.tnx;
                    if (w.tny < tn) { ax = 1; tn = w.tny; }
                    if (w.tnz < tn) { ax = 2; tn = w.tnz; }
                    const float t_end = tn < w.tmax ? tn : w.tmax;
                    float len = t_end - w.t;
                    if (len < 0.0f) len = 0.0f;
                    bool hit = false;
                    if (w.sb > 0.0f) {
                        const float dtau = w.sb * len;
                        if (tau < dtau) {
                            float t = w.t + tau / w.sb;
                            if (t > t_end) t = t_end;
                            w.t = t;
                            hit = true;
                            state = S_HIT;
                        } else {
                            tau -= dtau;
                        }
                    }
                    if (!hit) {
                        if (t_end > w.t) w.t = t_end;
                        bool end = !(tn < w.tmax);
                        if (!end) {
                            if (ax == 0) {
                                w.cx += seg.dx > 0.0f ? 1 : -1;
                                end = w.cx < 0 || w.cx >= P.mres[0];
                                w.tnx += fabsf(P.mcs[0] * seg.ix);
                            } else if (ax == 1) {
                                w.cy += seg.dy > 0.0f ? 1 : -1;
                                end = w.cy < 0 || w.cy >= P.mres[1];
                                w.tny += fabsf(P.mcs[1] * seg.iy);
                            } else {
                                w.cz += seg.dz > 0.0f ? 1 : -1;
                                end = w.cz < 0 || w.cz >= P.mres[2];
                                w.tnz += fabsf(P.mcs[2] * seg.iz);
                            }
                        }
                        if (end) {
                            // segment end: no collision
                            if (mode == M_DELTA) { did_scatter = false; state = S_VERTEX; }
                            else if (mode == M_NEE) state = S_NEE_DONE;
                            else if (mode == M_NEE_ADJ) state = S_PHASE;
                            else state = S_DRT_DONE;
                        } else {
                            w.sb = majorant_at<COUNT>(P, w.cx, w.cy, w.cz, K);
                        }
                    }
                }
                const int n = __popc(__ballot_sync(0xffffffffu, state == S_STEP));
                if (2 * n < n0 || n == 0) break;
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_HIT: {  // tentative collision: sigma_t tap + per-mode decision
            if (state == S_HIT) {
                const float px = fmaf(w.t, seg.dx, seg.ox), py = fmaf(w.t, seg.dy, seg.oy), pz = fmaf(w.t, seg.dz, seg.oz);
                float u2 = 0.0f;
                if (mode == M_DRT) u2 = draw(rng, K);
                const float st = sigma_tap(P, px, py, pz);
                K.add(C_SIGMA, 1);
                bool cont = true;
                if (mode == M_DELTA) {
                    // :354-361 real vs null collision
                    const float r = st / w.sb;
                    if (!(draw(rng, K) >= r)) {
                        did_scatter = true;
                        sigma_t = st;
                        vpx = px; vpy = py; vpz = pz;
                        state = S_VERTEX;
                        cont = false;
                    }
                } else if (mode == M_DRT) {
                    // sample_interaction_drt (App. B.6): candidate weight T/sigma_bar, size-1 reservoir
                    const float wi = T / w.sb;
                    drt_D += wi;
                    if (u2 <= wi / drt_D) {
                        drt_t = w.t;
                        drt_st = st;
                        drt_found = true;
                    }
                    T *= (w.sb - st) / w.sb;
                    if (!(T > 0.0f)) { state = S_DRT_DONE; cont = false; }
                } else {
                    // ratio tracking (:461-502); M_NEE_ADJ scatters -sum(adj)/sigma_n (:483-492)
                    const float sn = w.sb - st;
                    const float tr = sn / w.sb;
                    if (BWD && mode == M_NEE_ADJ && tr > 0.0f) {
                        scatter_sigma(P, px, py, pz, -asum / sn);
                        K.add(C_SSCAT, 1);
                    }
                    T *= tr;
                    if (T == 0.0f) { state = (mode == M_NEE) ? S_NEE_DONE : S_PHASE; cont = false; }
                }
                if (cont) {
                    tau = neg_log1m(draw(rng, K));
                    state = S_STEP;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_VERTEX: {  // end of a delta-tracking segment: real collision or escape (:130-200)
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
                            rs_ox = seg.ox; rs_oy = seg.oy; rs_oz = seg.oz;
                            rs_dx = seg.dx; rs_dy = seg.dy; rs_dz = seg.dz;
                            rs_tmax = seg.tmax;
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
                        const float interval = did_scatter ? w.t : seg.tmax;
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
                if (!did_scatter) {
                    escaped = true;  // :244-245
                    state = S_PATH_END;
                } else if (P.use_nee && active) {
                    state = S_NEE_START;
                } else {
                    state = S_PHASE;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_NEE_START: {  // sample_emitter (:406-433): direction + shadow segment
            if (state == S_NEE_START) {
                const float xi1 = draw(rng, K), xi2 = draw(rng, K);
                float wx, wy, wz;
                uniform_sphere(xi1, xi2, wx, wy, wz);
                nee_valid = make_segment(P, vpx, vpy, vpz, wx, wy, wz, seg);
                clone_state = rng.state;  // sampler.clone() position for the adjoint replay
                T = nee_valid ? 1.0f : 0.0f;
                if (nee_valid) {
                    mode = M_NEE;
                    state = S_WINIT;
                } else {
                    state = S_NEE_DONE;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_NEE_DONE: {  // sample_emitter_for_nee (:380-403): contribution, then replay in the adjoint
            if (state == S_NEE_DONE) {
                float contrib[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) contrib[c] = (beta[c] * P.half_le[c]) * T;
                state = S_PHASE;
                if (BWD && pass == P_DRTV) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) aux_Li[c] = contrib[c];
                } else if (BWD && pass == P_ADJ) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) R[c] = R[c] - contrib[c];
                    if (nee_valid) {
                        asum = (dL[0] * contrib[0] + dL[1] * contrib[1]) + dL[2] * contrib[2];
                        rng.state = clone_state;  // the replay consumes exactly the same draws again
                        T = 1.0f;
                        mode = M_NEE_ADJ;
                        state = S_WINIT;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) R[c] = R[c] + contrib[c];
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_PHASE: {  // phase sampling (:221-245) / DRT-vertex continuation (:626-652)
            if (state == S_PHASE) {
                draw(rng, K);
                const float xi1 = draw(rng, K), xi2 = draw(rng, K);
                float wx, wy, wz;
                uniform_sphere(xi1, xi2, wx, wy, wz);
                const bool ok = make_segment(P, vpx, vpy, vpz, wx, wy, wz, seg);
                mode = M_DELTA;
                if (BWD && pass == P_DRTV) {
                    depth += 1;
                    active = ok && (depth < P.max_depth);
                    pass = P_REC;
                    beta[0] = beta[1] = beta[2] = 1.0f;
                    R[0] = R[1] = R[2] = 0.0f;
                    escaped = false;
                    has_scattered = true;
                    if (active) draw(rng, K);  // :99 of the recursive sample()
                } else if (!ok) {
                    active = false;  // :240-241 accidental escape
                }
                state = active ? S_WINIT : S_PATH_END;
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_PATH_END: {
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
                        state = S_DONE;
                    } else {
                        pass = P_ADJ;  // batched.py:309-318: sample(Backward, state_in = L)
                        state = S_INIT;
                    }
                } else if (BWD && pass == P_ADJ) {
                    state = S_DONE;
                    if (use_rsv && rs_valid) {
                        // DRTReservoir.get (:756-760) and adjoint = weight * dL (:255)
                        const float d = mean3(rs_wcur), ws = mean3(rs_wsum);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float W = (d != 0.0f) ? (ws * rs_wcur[c]) / d : 0.0f;
                            dL[c] = W * dL[c];
                        }
                        seg.ox = rs_ox; seg.oy = rs_oy; seg.oz = rs_oz;
                        seg.dx = rs_dx; seg.dy = rs_dy; seg.dz = rs_dz;
                        exit_distance(seg);  // refresh 1/d (bitwise the same values as when stored)
                        seg.tmax = rs_tmax;
                        depth = rs_depth;
                        rng = alt;  // everything from here on draws from the alt stream
                        mode = M_DRT;
                        T = 1.0f;
                        drt_D = 0.0f;
                        drt_found = false;
                        state = S_WINIT;
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
                    state = S_DONE;
                }
            }
            break;
        }
        // ------------------------------------------------------------------------------
        case S_DRT_DONE: {  // DRT walk finished (:550-558); set up the DRT vertex
            if (state == S_DRT_DONE) {
                if (BWD && drt_found) {
                    vpx = fmaf(drt_t, rs_dx, rs_ox); vpy = fmaf(drt_t, rs_dy, rs_oy); vpz = fmaf(drt_t, rs_dz, rs_oz);
                    albedo_tap(P, vpx, vpy, vpz, aux_alb);
                    K.add(C_ALBEDO, 1);
                    aux_Li[0] = aux_Li[1] = aux_Li[2] = 0.0f;
                    beta[0] = beta[1] = beta[2] = 1.0f;
                    pass = P_DRTV;
                    state = P.use_nee ? S_NEE_START : S_PHASE;
                } else {
                    state = S_DONE;
                }
            }
            break;
        }
        default: break;
        }
    }
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
