/*
 * uivr_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the hot path of rgl-epfl/unbiased-inverse-volume-rendering:
 * the `volpathsimple` integrator (python/integrators/volpathsimple.py) driven the way
 * RBIntegrator.render / render_backward drive it (restated in python/batched.py:134-326).
 *
 * How it is pinned: the reference cannot be installed here (un-vendored, un-pinned Mitsuba 3
 * branch + Dr.Jit, README.md:97-104) and its tests hold no golden vectors for this path
 * (tests/test_integrators.py:343-347 is disabled), but its Python FILES can be imported:
 * oracle/refshim.py runs volpathsimple.py / nerf.py / batched.py / opt_config.py unmodified on a
 * numpy stand-in for the Mitsuba / Dr.Jit API and tests/golden/refshim_*.npz hold their outputs.
 * Those vectors pin everything the reference's files decide (state machine, masks, RNG draw order,
 * gradient formulae, film, sub-seeds); this file is tested against them to 1e-6.
 * PARITY UNPINNED only for the arithmetic INSIDE the upstream Mitsuba branch (Medium::
 * sample_interaction(_drt), GridVolume lookup, PCG32 stream assignment, perspective sensor ...):
 * this oracle DEFINES it (DESIGN.md "Arithmetic contract") and the stand-in borrows it from the
 * uivr_oracle_shim_* wrappers below.  Further pins: PCG32 public known-answer vectors, analytic
 * known answers (KA1-KA3) and finite differences (KA4, the reference's own methodology,
 * python/fd.py) -- see tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (csrc/) never links or calls it.
 */
#ifndef UIVR_ORACLE_H
#define UIVR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scene + integrator description (all host memory). */
typedef struct {
    int32_t res[3];          /* grid resolution X, Y, Z; tensors are (Z,Y,X,C), x fastest     */
    float   to_local[12];    /* world->local affine, row-major 3x4; medium = local [0,1]^3    */
    float   scale;           /* medium 'scale' (tests/test_integrators.py:83)                 */
    int32_t majorant_factor; /* supergrid resolution factor; <=1 -> one global majorant       */
    float   cam_origin[3];   /* perspective sensor (tests/test_integrators.py:46-53)          */
    float   cam_left[3];
    float   cam_up[3];
    float   cam_dir[3];
    float   tan_x, tan_y;    /* tan(fov_x/2), tan_x*H/W                                       */
    float   near_clip;
    int32_t width, height;   /* box-filter hdrfilm                                            */
    float   radiance[3];     /* constant emitter (tests/test_integrators.py:73-77)            */
    int32_t max_depth;       /* integrator props (volpathsimple.py:19-34)                     */
    int32_t hide_emitters;
    int32_t use_nee;
    int32_t use_drt;
    int32_t use_drt_subsampling;
    int32_t use_drt_mis;
    /* emitter = lat-long environment map instead of the constant `radiance` when env_data != NULL
     * ("next" row SURVEY 8f rank 4; Mitsuba `envmap`, volpathsimple.py:262-285, :419).  Tables as
     * built by the host (scene.py EnvMap.tables): vertices (env_h, env_w + 1, 4) = RGB radiance +
     * sampling density of the bilinear patch (y, x); row marginal CDF (env_h - 1); per-row
     * conditional CDFs (env_h - 1, env_w).  Rotations are row-major 3x3. */
    float   local_to_world[9];   /* linear part of medium-local -> world */
    const float* env_data;
    int32_t env_w, env_h;
    float   env_scale;
    const float* env_marg;
    const float* env_cond;
    float   env_to_world[9], world_to_env[9];
} uivr_oracle_scene;

/* Ray-batch rendering (python/batched.py:88-131, "next" row SURVEY 8f rank 2): instead of the
 * pixels of one sensor, the wavefront is a batch of B (sensor, pixel) pairs drawn uniformly, spp
 * samples each; the film is (B x 1) with a box filter.  Set scene->width = B, scene->height = 1.
 *   element b:  stream (seed_pixels, b):  sensor = uint(n_sensors * u), pixel = uint((W,H) * (u,u))
 *                                                                      (batched.py:416-421)
 *   sample  k = b*spp + j:  sub-pixel offset = two draws of stream (seed_offsets, k)  (:437-439);
 *   the path sampler (seed, k) draws NO jitter (prepare_batch seeds it for integrator.sample only).
 * seed_pixels = tea32(seed, 5); seed_offsets = tea32(seed, 22) for the primal, tea32(seed, 39) for
 * the adjoint (batched.py:409-413: sub_seed_i = tea32(seed, 17 i + 5)). */
typedef struct {
    int32_t n_sensors;
    const float* sensors;    /* n_sensors x 16: origin[3] left[3] up[3] dir[3] tan_x tan_y near_clip pad */
    int32_t film_w, film_h;  /* film size shared by all sensors (batched.py:428) */
    uint32_t seed_pixels, seed_offsets;
} uivr_oracle_batch;

/* NeRFIntegrator properties (python/integrators/nerf.py:27-35; "next" row SURVEY 8f rank 4).
 * density_noise_std is not restated: the reference flags it incorrect itself (nerf.py:157). */
typedef struct {
    int32_t queries_per_ray;    /* 128 */
    int32_t jittering_enabled;  /* True */
    int32_t activation;         /* 0 identity, 1 relu */
    int32_t hide_emitters;
} uivr_oracle_nerf;

/* Pixel sharding: pixel p belongs to this call iff (p / shard_block) % shard_count == shard_rank. */
typedef struct {
    int32_t shard_rank, shard_count, shard_block;
} uivr_oracle_shard;

enum {
    UIVR_ORC_SIGMA_TAPS = 0,   /* trilinear sigma_t lookups (8 voxels each)          */
    UIVR_ORC_ALBEDO_TAPS,      /* trilinear albedo lookups (8 voxels x 3 ch)         */
    UIVR_ORC_MAJORANT_READS,   /* supergrid cell reads                               */
    UIVR_ORC_SIGMA_SCATTERS,   /* gradient scatter events into d sigma_t (8 voxels)  */
    UIVR_ORC_ALBEDO_SCATTERS,  /* gradient scatter events into d albedo (8 x 3)      */
    UIVR_ORC_CAMERA_HITS,      /* primary rays entering the medium                   */
    UIVR_ORC_REAL_COLLISIONS,  /* real scattering events (all path kinds)            */
    UIVR_ORC_RNG_DRAWS,        /* PCG32 outputs consumed (primary + alt streams)     */
    UIVR_ORC_SAMPLES,          /* samples processed                                  */
    UIVR_ORC_NUM_COUNTERS
};

/* ---- primitives (KA5 + bitwise GPU-vs-oracle checks) ---- */
void  uivr_oracle_tea(uint32_t v0, uint32_t v1, uint32_t out[2]);
void  uivr_oracle_pcg32_stream(uint64_t initstate, uint64_t initseq, int n, uint32_t* out);
void  uivr_oracle_sampler_floats(uint32_t seed, uint32_t idx, int n, float* out);
void  uivr_oracle_neg_log1m(const float* u, int n, float* out);
void  uivr_oracle_sincos2pi(const float* x, int n, float* s, float* c);
uint32_t uivr_oracle_alt_seed(uint32_t seed_grad);        /* mi.render: 4th float of lane 0 */
uint32_t uivr_oracle_alt_seed_batch(uint32_t seed_grad);  /* render_batch: 2nd float (no jitter draws) */
/* trilinear lookup at local points p (n x 3); grid (Z,Y,X,C) */
void  uivr_oracle_trilinear(const float* grid, const int32_t res[3], int channels,
                            const float* p, int n, float* out);

/* supergrid of local majorants; out has M[0]*M[1]*M[2] floats, M written to mres */
void  uivr_oracle_build_majorant(const float* sigma_t, const int32_t res[3], float scale,
                                 int32_t factor, int32_t mres[3], float* out);

/* exit mask of the supergrid (one byte per cell, bit o: only empty cells ahead in octant o; octant bit a =
 * ray direction negative along axis a): lets a walk stop as soon as nothing but empty space is ahead */
void  uivr_oracle_build_exit_mask(const float* majorant, const int32_t mres[3], uint8_t* out);
void  uivr_oracle_set_exit_mask(int enable);  /* test hook (default on) */
void  uivr_oracle_set_remaining_by_difference(int enable);  /* test hook (default off), see path_loop */
/* counters only: book as "replay" events just the second shadow walks a collision log of this capacity makes
 * unnecessary (0: all of them); _overflows = shadow walks of the last backward calls that exceeded it */
void  uivr_oracle_set_nee_log_capacity(int capacity);
uint64_t uivr_oracle_nee_log_overflows(void);

/* events of the primal pass inside the most recent backward call (they are part of its `counters`) */
void  uivr_oracle_last_backward_primal_counters(uint64_t* out);
/* ... and of the second walk the NEE adjoint makes over every shadow segment (volpathsimple.py:393-401) */
void  uivr_oracle_last_backward_replay_counters(uint64_t* out);

/* ---- the path ---- */
/* image_out: H*W*3 (overwritten; pixels outside the shard are zero).
 * sample_L_out: optional (NULL) S*3 per-sample radiance, S = W*H*spp.
 * counters: optional, UIVR_ORC_NUM_COUNTERS uint64 (accumulated into). */
int uivr_oracle_render_forward(const uivr_oracle_scene* scene, const float* sigma_t,
                               const float* albedo, uint32_t seed, int32_t spp,
                               const uivr_oracle_shard* shard, int nthreads,
                               float* image_out, float* sample_L_out, uint64_t* counters);

/* grad_image: H*W*3.  dsigma_out: Z*Y*X doubles, dalbedo_out: Z*Y*X*3 doubles (overwritten).
 * sample_L_out: optional per-sample primal radiance of the seed_grad pass. */
int uivr_oracle_render_backward(const uivr_oracle_scene* scene, const float* sigma_t,
                                const float* albedo, const float* grad_image,
                                uint32_t seed_grad, int32_t spp_grad,
                                const uivr_oracle_shard* shard, int nthreads,
                                double* dsigma_out, double* dalbedo_out,
                                float* sample_L_out, uint64_t* counters);

/* same as the two calls above with the wavefront drawn per `batch` (image: B x 1 x 3) */
int uivr_oracle_render_batch_forward(const uivr_oracle_scene* scene, const uivr_oracle_batch* batch,
                                     const float* sigma_t, const float* albedo, uint32_t seed,
                                     int32_t spp, int nthreads, float* image_out,
                                     float* sample_L_out, uint64_t* counters);
int uivr_oracle_render_batch_backward(const uivr_oracle_scene* scene, const uivr_oracle_batch* batch,
                                      const float* sigma_t, const float* albedo,
                                      const float* grad_image, uint32_t seed_grad, int32_t spp_grad,
                                      int nthreads, double* dsigma_out, double* dalbedo_out,
                                      float* sample_L_out, uint64_t* counters);
/* the (sensor, pixel x, pixel y) triple of batch element b (host-side check of the index sampler) */
void uivr_oracle_batch_element(const uivr_oracle_batch* batch, uint32_t b, uint32_t out[3]);

/* ---- nerf integrator (python/integrators/nerf.py): emission-absorption ray marching over the
 * sigma_t grid and an RGB emission grid (Z,Y,X,3); same drivers, film and seeds as above ---- */
int uivr_oracle_nerf_forward(const uivr_oracle_scene* scene, const uivr_oracle_nerf* nerf,
                             const uivr_oracle_batch* batch, /* NULL: the scene's sensor; else ray-batch mode */
                             const float* sigma_t, const float* emission, uint32_t seed, int32_t spp,
                             const uivr_oracle_shard* shard, int nthreads, float* image_out,
                             float* sample_L_out, uint64_t* counters);
int uivr_oracle_nerf_backward(const uivr_oracle_scene* scene, const uivr_oracle_nerf* nerf,
                              const uivr_oracle_batch* batch,
                              const float* sigma_t, const float* emission, const float* grad_image,
                              uint32_t seed_grad, int32_t spp_grad, const uivr_oracle_shard* shard,
                              int nthreads, double* dsigma_out, double* demission_out,
                              float* sample_L_out, uint64_t* counters);
void uivr_oracle_exp(const float* x, int n, float* out); /* the exact-op exp both sides use */

/* ---- upstream primitives for oracle/refshim.py ----
 * refshim runs the reference's UNMODIFIED python/integrators/volpathsimple.py (and the drivers of
 * python/batched.py) on a small numpy stand-in for the Mitsuba 3 / Dr.Jit API.  The control flow,
 * masks, RNG draw order and gradient formulae then come from the reference files themselves; the
 * upstream arithmetic (un-vendored Mitsuba branch: Medium::sample_interaction(_drt), GridVolume
 * lookup + adjoint, box intersection, sensor, sphere warp) is supplied by the functions below,
 * which wrap the same static functions the oracle's own path code uses.  Lane arrays of length n;
 * masks are bytes; vectors are [n][3] row-major; positions and directions are in the medium's
 * local space with distances in world units (the oracle's segment convention). */
typedef struct uivr_oracle_shim uivr_oracle_shim;
uivr_oracle_shim* uivr_oracle_shim_create(const uivr_oracle_scene* scene, const float* sigma_t, const float* albedo);
void uivr_oracle_shim_destroy(uivr_oracle_shim* h);
/* film position of pixel pix with jitter (Sensor::sample_ray's sample2) */
void uivr_oracle_shim_film_uv(const uivr_oracle_shim* h, int n, const uint32_t* pix, const float* jx,
                              const float* jy, float* u, float* v);
/* Sensor::sample_ray_differential: primary ray through (u, v); frames == NULL: the scene's sensor,
 * else frames[frame_idx[i]] (16 floats each, uivr_oracle_batch layout) */
void uivr_oracle_shim_camera_ray(const uivr_oracle_shim* h, int n, const float* frames, const int32_t* frame_idx,
                                 const float* u, const float* v, float* o, float* d);
/* scene.ray_intersect for an origin not known to be inside: kind 0 miss, 1 entry at t, 2 far wall at t */
void uivr_oracle_shim_box_entry(int n, const float* o, const float* d, float* t, int32_t* kind);
/* si.spawn_ray into the medium: origin just inside the entry point */
void uivr_oracle_shim_entry_spawn(int n, const float* o, const float* d, const float* t, float* o_new);
/* scene.ray_intersect from inside the medium: distance to the box boundary; ok iff 0 < t < inf */
void uivr_oracle_shim_exit(int n, const float* o, const float* d, float* t, uint8_t* ok);
void uivr_oracle_shim_dir_to_local(const uivr_oracle_shim* h, int n, const float* w, float* d);
/* Medium::sample_interaction(ray, u): DDA from the ray origin (restarted per call, as the
 * reference does, volpathsimple.py:331-334); t relative to o; valid iff a collision precedes maxt */
void uivr_oracle_shim_sample_interaction(const uivr_oracle_shim* h, int n, const float* o, const float* d,
                                         const float* maxt, const float* u, const uint8_t* active,
                                         float* t, float* sigma_t, float* sigma_bar, uint8_t* valid);
/* Medium::sample_interaction_drt(ray, sampler): advances the lanes' PCG32 states in place */
void uivr_oracle_shim_sample_interaction_drt(const uivr_oracle_shim* h, int n, const float* o, const float* d,
                                             const float* maxt, uint64_t* rng_state, const uint64_t* rng_inc,
                                             const uint8_t* active, float* t, float* sigma_t, float* weight,
                                             uint8_t* valid);
/* which = 0: sigma_t(p) = scale * grid (1 value), 1: albedo(p) (3 values), 2: majorant at p (1 value) */
void uivr_oracle_shim_lookup(const uivr_oracle_shim* h, int which, int n, const float* p, float* out);
/* adjoint of lookup 0 / 1: scatter-add g (1 or 3 floats per lane) into dgrid (doubles) where mask */
void uivr_oracle_shim_scatter(const uivr_oracle_shim* h, int which, int n, const float* p, const float* g,
                              const uint8_t* mask, double* dgrid);
/* envmap emitter: Emitter::eval + pdf_direction for a ray leaving along local direction d;
 * Scene::sample_emitter_direction (direction in local space, solid-angle pdf, radiance) */
void uivr_oracle_shim_env_eval(const uivr_oracle_shim* h, int n, const float* d, float* le, float* pdf);
void uivr_oracle_shim_env_sample(const uivr_oracle_shim* h, int n, const float* xi1, const float* xi2,
                                 float* d, float* pdf, float* le);
void uivr_oracle_atan2_turns(const float* y, const float* x, int n, float* out); /* atan2(y,x)/(2 pi) */
void uivr_oracle_shim_uniform_sphere(int n, const float* xi1, const float* xi2, float* w);
void uivr_oracle_shim_fma(int n, const float* a, const float* b, const float* c, float* out); /* fmaf */

/* ---- optimiser step ("next" row) ----
 * mi.ad.Adam.step() (python/opt_config.py:46-48, python/optimize.py:352; update rule SURVEY App.
 * B.10) followed by enforce_valid_params (python/optimize.py:169-179): clip to [lo, hi].
 *   step = lr * sqrt(1 - beta2^t) / (1 - beta1^t)           (double on the host, then float)
 *   m = fma(beta1, m, (1 - beta1) * g);  v = fma(beta2, v, ((1 - beta2) * g) * g)
 *   p = clip(p - (step * m) / (sqrt(v) + eps), lo, hi) */
float uivr_oracle_adam_step_size(float lr, float beta1, float beta2, int32_t t);
void  uivr_oracle_adam_step(float* param, const float* grad, float* m, float* v, uint64_t n,
                            float lr, float beta1, float beta2, float eps, int32_t t,
                            float lo, float hi);

#ifdef __cplusplus
}
#endif
#endif
