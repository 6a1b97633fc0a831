/*
 * uivr.h -- C-ABI of the B200-native differential volumetric path tracer (libuivr.so).
 *
 * Drop-in boundary for the hot path of rgl-epfl/unbiased-inverse-volume-rendering:
 * everything `mi.render(scene, params, integrator='volpathsimple', ...)` + `dr.backward(loss)`
 * execute on the device for one sensor (python/optimize.py:345-350), i.e.
 *   RBIntegrator.render / render_backward   (restated in python/batched.py:134-197, 212-326)
 *   VolpathSimpleIntegrator.sample          (python/integrators/volpathsimple.py:38-290)
 *   Medium::sample_interaction(_drt), GridVolume lookup/adjoint, majorant supergrid,
 *   PCG32 `independent` sampler, perspective sensor, box hdrfilm   (un-vendored Mitsuba 3
 *   branch; call sites python/integrators/volpathsimple.py:348,469,550,375,141,...).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types; every call returns 0 on success,
 *    a negative uivr_status otherwise (uivr_last_error(ctx) gives the message).  No C++
 *    exception crosses the boundary.
 *  - "d_" pointers are DEVICE pointers owned by the caller (e.g. torch tensors); the library
 *    never frees them.  Tensors are contiguous float32: sigma_t (Z,Y,X[,1]), albedo (Z,Y,X,3),
 *    image / grad_image (H,W,3).  "h_" pointers are HOST pointers (the *_host entry points
 *    stage them through context-owned device buffers, copies on `stream`).
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream unless
 *    stated otherwise.  A context is bound to one device and is not thread-safe; multi-GPU =
 *    one context per rank (pixel sharding via uivr_shard, gradients summed by the caller,
 *    e.g. ncclAllReduce).
 */
#ifndef UIVR_H
#define UIVR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uivr_ctx uivr_ctx;

typedef enum {
    UIVR_OK = 0,
    UIVR_ERR_INVALID = -1,    /* bad argument / precondition (reference: Python assert / raise) */
    UIVR_ERR_CUDA = -2,       /* CUDA runtime error */
    UIVR_ERR_STATE = -3,      /* call order (scene / medium not set) */
    UIVR_ERR_NOMEM = -4,
    UIVR_ERR_WATCHDOG = -5    /* the persistent kernel made no progress and aborted (results invalid) */
} uivr_status;

/* Scene: one medium box + perspective sensor + constant emitter.
 * Replaces the Mitsuba scene dict of tests/test_integrators.py:19-116 (cube_test_scene). */
typedef struct {
    int32_t res[3];          /* grid resolution X, Y, Z */
    float   to_local[12];    /* world->local affine (row-major 3x4); medium = local [0,1]^3 */
    float   scale;           /* medium 'scale' (tests/test_integrators.py:83) */
    int32_t majorant_factor; /* supergrid factor (scene_config.py:36; optimize.py:182-199); <=1: global */
    float   cam_origin[3], cam_left[3], cam_up[3], cam_dir[3]; /* look_at frame (:46-53) */
    float   tan_x, tan_y;    /* tan(fov_x/2), tan_x*H/W */
    float   near_clip;
    int32_t width, height;   /* box-filter hdrfilm (:58-66) */
    float   radiance[3];     /* constant emitter (:73-77) */
} uivr_scene_desc;

/* Integrator properties: python/integrators/volpathsimple.py:19-34 (+ max_depth of the base
 * class; rr_depth is always > max_depth, python/opt_config.py:105-106). */
typedef struct {
    int32_t max_depth;
    int32_t hide_emitters;
    int32_t use_nee;
    int32_t use_drt;
    int32_t use_drt_subsampling;
    int32_t use_drt_mis;
} uivr_integrator_props;

/* Ray-batch rendering: render_batch(batch_size, scene, sensors, film_size, ..., seed, seed_grad, spp,
 * spp_grad) of python/batched.py:88-131 (the reference's production mode, python/optimize.py:334).
 * The wavefront is a batch of B (sensor, pixel) pairs drawn uniformly -- element b: stream
 * (tea32(seed, 5), b): sensor = uint(n_sensors u), pixel = uint((W, H) (u, u)) (batched.py:409-421) --
 * with spp samples each; sub-pixel offsets come from stream (tea32(seed, 22), k) in the primal and
 * (tea32(seed, 39), k) in the adjoint, k = b spp + j (batched.py:409-413, :437-439); the film is
 * (B x 1), box filter: image / grad_image are [B, 3].  All sensors share the film size. */
typedef struct {
    int32_t n_sensors;
    const float* sensors;   /* HOST, n_sensors x 16 floats: origin[3] left[3] up[3] dir[3] tan_x tan_y near_clip 0 */
    int32_t film_w, film_h;
    int32_t batch_size;     /* B */
    uint32_t seed;          /* the `seed` of render_batch: selects the pixels; uivr_render_forward must get the same */
} uivr_batch_desc;

/* Pixel p is rendered by this call iff (p / block) % count == rank.  count<=1: all pixels.
 * RNG streams are keyed by the GLOBAL sample index, so results do not depend on the split. */
typedef struct {
    int32_t rank, count, block;
} uivr_shard;

enum {
    UIVR_CNT_SIGMA_TAPS = 0,  /* trilinear sigma_t lookups                 (32 B each)   */
    UIVR_CNT_ALBEDO_TAPS,     /* trilinear albedo lookups                  (96 B each)   */
    UIVR_CNT_MAJORANT_READS,  /* supergrid cell reads                      (4 B each)    */
    UIVR_CNT_SIGMA_SCATTERS,  /* gradient scatters into d sigma_t          (64 B each)   */
    UIVR_CNT_ALBEDO_SCATTERS, /* gradient scatters into d albedo           (192 B each)  */
    UIVR_CNT_CAMERA_HITS,
    UIVR_CNT_REAL_COLLISIONS,
    UIVR_CNT_RNG_DRAWS,
    UIVR_CNT_SAMPLES,
    UIVR_NUM_COUNTERS
};

/* ---- lifetime ---- */
int         uivr_create(int device, uivr_ctx** out);
int         uivr_destroy(uivr_ctx* ctx);
const char* uivr_last_error(const uivr_ctx* ctx);
int         uivr_version(void);

/* ---- configuration (host side, cheap) ---- */
int uivr_set_scene(uivr_ctx* ctx, const uivr_scene_desc* scene);              /* mi.load_dict(scene) */
int uivr_set_integrator(uivr_ctx* ctx, const uivr_integrator_props* props);   /* IntegratorConfig.create, opt_config.py:97-108 */

/* Enter (batch != NULL) or leave (NULL) ray-batch mode.  In batch mode uivr_render_forward /
 * uivr_render_backward render the batch: d_image / d_grad_image are [B, 3]; the sensor and film of
 * uivr_set_scene are ignored; shards split the batch elements.  Needs the slot-pool kernels (variant 3). */
int uivr_set_batch(uivr_ctx* ctx, const uivr_batch_desc* batch);

/* The scene's emitter is an `envmap` (lat-long environment map; every scene of
 * python/scene_config.py:102-340) instead of the `constant` emitter of uivr_scene_desc.radiance:
 * Emitter::eval + pdf_direction on escape (python/integrators/volpathsimple.py:262-285),
 * Scene::sample_emitter_direction for next-event estimation (:419).  All pointers are HOST memory,
 * copied by the call; tables as produced by the host mirror's EnvMap.tables() (scene.py):
 *   data  [env_h][env_w + 1][4]  vertex radiance RGB (first column repeated at the end) and, in .w,
 *                                the sampling density over [0,1]^2 of the bilinear patch (y, x)
 *   marg  [env_h - 1]            CDF over patch rows;   cond [env_h - 1][env_w]  per-row CDFs
 * Rotations are row-major 3x3; local_to_world is the linear part of the inverse of
 * uivr_scene_desc.to_local.  NULL returns to the constant emitter.  The slot-pool kernels have
 * envmap template instances; the constant-emitter instances are unaffected. */
typedef struct {
    int32_t env_w, env_h;     /* resolution of the source image (H >= 2) */
    float   scale;
    const float* data;
    const float* marg;
    const float* cond;
    float   env_to_world[9], world_to_env[9], local_to_world[9];
} uivr_envmap_desc;
int uivr_set_envmap(uivr_ctx* ctx, const uivr_envmap_desc* env);

/* params.update(): rebuild the device-side lookup structures derived from sigma_t -- the
 * corner-octet tap layout and the majorant supergrid (upstream does the latter on
 * params.update(), triggered at optimize.py:165, :251, :354).  Must be called after every
 * change of sigma_t and before rendering. */
int uivr_update_medium(uivr_ctx* ctx, const float* d_sigma_t, void* stream);

/* ---- the path ---- */
/* mi.render(...) primal: image[H,W,3] = box-film mean over spp of sample(Primal) (optimize.py:345).
 * d_sample_L (optional, may be NULL): per-sample radiance [W*H*spp,3] for bit-level parity tests. */
int uivr_render_forward(uivr_ctx* ctx, const float* d_albedo, uint32_t seed, int32_t spp,
                        const uivr_shard* shard, float* d_image, float* d_sample_L, void* stream);

/* dr.backward(loss) -> RBIntegrator.render_backward (batched.py:212-326): the primal radiance at
 * seed_grad (gathered inside the adjoint replay by the slot-pool kernels, a separate primal pass in the
 * one-sample-per-lane kernels), the path-replay adjoint with the three gradient estimators, then DRT.
 * d_dsigma_t [Z,Y,X] and d_dalbedo [Z,Y,X,3] are OVERWRITTEN with this call's gradients. */
int uivr_render_backward(uivr_ctx* ctx, const float* d_albedo, const float* d_grad_image,
                         uint32_t seed_grad, int32_t spp_grad, const uivr_shard* shard,
                         float* d_dsigma_t, float* d_dalbedo, float* d_sample_L, void* stream);

/* Host-buffer variants (end-to-end path: H2D of the parameters / grad_image, D2H of the
 * results inside the call; synchronises `stream` before returning).  They include
 * uivr_update_medium.  uivr_render_backward_host accepts h_sigma_t == h_albedo == NULL to
 * reuse the parameters staged by the previous *_host call (dr.backward follows mi.render on
 * unchanged parameters, optimize.py:345-350). */
int uivr_render_forward_host(uivr_ctx* ctx, const float* h_sigma_t, const float* h_albedo,
                             uint32_t seed, int32_t spp, const uivr_shard* shard,
                             float* h_image, void* stream);
int uivr_render_backward_host(uivr_ctx* ctx, const float* h_sigma_t, const float* h_albedo,
                              const float* h_grad_image, uint32_t seed_grad, int32_t spp_grad,
                              const uivr_shard* shard, float* h_dsigma_t, float* h_dalbedo,
                              void* stream);

/* ---- `nerf` integrator ("next" row): python/integrators/nerf.py ----
 * NeRFIntegrator(props) (nerf.py:27-35, registered at :168, registry entry opt_config.py:162-169):
 * emission-absorption ray marching, `queries_per_ray` forward-looking steps per ray with one jitter
 * draw, over the sigma_t grid of uivr_update_medium and an RGB emission grid d_emission [Z,Y,X,3]
 * (medium.get_emission).  Same film, seeds, shards and ray-batch mode as the calls above;
 * uivr_set_integrator is not needed.  `density_noise_std` is not offered: the reference marks it
 * incorrect (nerf.py:157). */
enum { UIVR_NERF_IDENTITY = 0, UIVR_NERF_RELU = 1 };  /* 'activation' (nerf.py:38-45) */
typedef struct {
    int32_t queries_per_ray;    /* 128 */
    int32_t jittering_enabled;  /* True */
    int32_t activation;         /* UIVR_NERF_IDENTITY */
    int32_t hide_emitters;      /* False */
} uivr_nerf_props;
/* NeRFIntegrator.sample(Primal) under mi.render: image[H,W,3] (nerf.py:47-147) */
int uivr_nerf_forward(uivr_ctx* ctx, const uivr_nerf_props* props, const float* d_emission, uint32_t seed,
                      int32_t spp, const uivr_shard* shard, float* d_image, float* d_sample_L, void* stream);
/* dr.backward(loss): primal replay at seed_grad + sample(Backward) (nerf.py:109-124).
 * d_dsigma_t [Z,Y,X] and d_demission [Z,Y,X,3] are OVERWRITTEN. */
int uivr_nerf_backward(uivr_ctx* ctx, const uivr_nerf_props* props, const float* d_emission,
                       const float* d_grad_image, uint32_t seed_grad, int32_t spp_grad, const uivr_shard* shard,
                       float* d_dsigma_t, float* d_demission, float* d_sample_L, void* stream);

/* ---- optimisation step ("next" row after the path itself) ----
 * opt.step() of mi.ad.Adam (python/opt_config.py:46-48, python/optimize.py:352) fused with
 * enforce_valid_params (python/optimize.py:169-179, :353): for every element
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p = clip(p - lr_t m / (sqrt(v) + eps), lo, hi),
 *   lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),  t = 1, 2, ...
 * All arrays are DEVICE pointers of n floats (16-byte aligned); param / m / v are updated in
 * place.  Call uivr_update_medium afterwards when the tensor is sigma_t (params.update(),
 * python/optimize.py:354). */
int uivr_adam_step(uivr_ctx* ctx, float* d_param, const float* d_grad, float* d_m, float* d_v,
                   uint64_t n, float lr, float beta1, float beta2, float eps, int32_t t,
                   float lo, float hi, void* stream);

/* Multires step: upsample_grid(values, old_res, 2 * old_res) of python/optimize.py:203-252 (first-order
 * scipy zoom with grid_mode=True, mode='nearest').  d_in (Z,Y,X,C) -> d_out (2Z,2Y,2X,C), both DEVICE. */
int uivr_upsample2x(uivr_ctx* ctx, const float* d_in, const int32_t res[3], int32_t channels, float* d_out,
                    void* stream);

/* ---- instrumentation ---- */
/* Event counters (SURVEY §8d algorithmic bytes).  Counting kernels are separate template
 * instances; enable only for accounting passes, not for timing. */
int uivr_set_counting(uivr_ctx* ctx, int enable);
int uivr_reset_counters(uivr_ctx* ctx, void* stream);
int uivr_get_counters(uivr_ctx* ctx, uint64_t out[UIVR_NUM_COUNTERS], void* stream); /* synchronises */
/* CUDA-event duration (ms) of the most recent path megakernel launched by this context:
 * which = 0 forward (sample(Primal) kernel), 1 backward (adjoint replay + DRT launches + gradient folds; the one-sample-per-lane route: its single kernel).
 * Events are recorded on the stream the kernel was launched on; synchronises on the end event. */
int uivr_get_kernel_ms(uivr_ctx* ctx, int which, float* ms);
/* Synchronises `stream` and reports whether a persistent path kernel launched by this context
 * has tripped its progress watchdog since the last check (UIVR_ERR_WATCHDOG; the record is
 * cleared).  out (optional): 64 raw words of the record.  A healthy run never trips it. */
int uivr_check_watchdog(uivr_ctx* ctx, uint32_t out[64], void* stream);
/* number of kernel launches issued by this context so far */
int uivr_get_launch_count(const uivr_ctx* ctx, uint64_t* out);
/* kernel variant: 3 (default) = persistent slot-pool kernels (CTA-wide compaction of live rays through
 * shared-memory queues; walker warps run the supergrid DDA, handler warps everything else in full batches;
 * the backward runs as adjoint replay (which gathers the primal radiance itself: no primal pass) -> DRT launch,
 * handing per-sample state through context-owned HBM scratch); 1 = one sample per lane, run to completion
 * (in-GPU cross-check, the O(n^2) `use_drt_subsampling = False` mode, and backward calls with max_depth > 255).
 * Other values: UIVR_ERR_INVALID. */
int uivr_set_variant(uivr_ctx* ctx, int variant);
/* Test hook for the watchdog of the slot-pool kernels: a walker warp that steps more than `limit` supergrid cells
 * in one go trips it (default 2^24: never on a sane scene; limit <= 0 restores the default).  With a tiny limit
 * every render call aborts in bounded time and uivr_check_watchdog returns UIVR_ERR_WATCHDOG -- which is what
 * tests/test_gpu_parity.py checks: a scheduling bug must end as an error, not as a hung GPU. */
int uivr_debug_set_walk_limit(uivr_ctx* ctx, int limit);

/* ---- device primitives exposed for bit-exactness tests (all arrays are DEVICE pointers) ---- */
/* out[0:n] = -ln(1-u);  s,c = sin/cos(2 pi x);  sampler floats of stream (seed, idx) */
int uivr_test_neg_log1m(uivr_ctx* ctx, const float* d_u, int n, float* d_out, void* stream);
int uivr_test_exp(uivr_ctx* ctx, const float* d_x, int n, float* d_out, void* stream); /* exp(x), nerf.py:104 */
int uivr_test_atan2_turns(uivr_ctx* ctx, const float* d_y, const float* d_x, int n, float* d_out, void* stream); /* atan2/(2 pi) */
int uivr_test_sincos2pi(uivr_ctx* ctx, const float* d_x, int n, float* d_s, float* d_c, void* stream);
int uivr_test_sampler(uivr_ctx* ctx, uint32_t seed, uint32_t idx0, int nstreams, int ndraws,
                      float* d_out, void* stream);
/* sigma_t(p) through the octet layout (after uivr_update_medium), p = n x 3 local points */
int uivr_test_sigma_lookup(uivr_ctx* ctx, const float* d_p, int n, float* d_out, void* stream);
/* copy the current majorant supergrid to d_out (mres[0]*mres[1]*mres[2] floats) */
int uivr_get_majorant(uivr_ctx* ctx, int32_t mres[3], float* d_out, void* stream);
/* copy the walk table to d_out: (mres[2]+2)*(mres[1]+2)*(mres[0]+2) words, the supergrid with a one-cell
 * border.  Non-empty cell: bits of its majorant; empty cell: 0x80000000 | exit mask (bit o: only empty
 * cells ahead in octant o, octant bit a = direction negative along axis a); border: 0x800001FF. */
int uivr_get_walk_table(uivr_ctx* ctx, int32_t mres[3], uint32_t* d_out, void* stream);
uint32_t uivr_tea32(uint32_t v0, uint32_t v1);          /* mi.sample_tea_32(v0, v1)[0] */
uint32_t uivr_alt_seed(uint32_t seed_grad);             /* volpathsimple.py:99-107 under mi.render */
uint32_t uivr_alt_seed_batch(uint32_t seed_grad);       /* same under render_batch (no jitter draws, batched.py:390) */

#ifdef __cplusplus
}
#endif
#endif
