// uivr_api.cu -- C-ABI of libuivr.so (see include/uivr.h).  Host side: context, launch
// configuration, derived-layout management.  No torch types, no exceptions across the ABI.
#include "../../include/uivr.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "uivr_kernels.cuh"
#include "uivr_pool.cuh"
#include "uivr_nerf.cuh"

using namespace uivr;

struct uivr_ctx {
    int device = 0;
    int num_sms = 148;
    std::string err;
    bool have_scene = false, have_props = false, have_medium = false;
    uivr_scene_desc scene{};
    uivr_integrator_props props{};
    int mres[3] = {1, 1, 1};
    float4* oct = nullptr;
    size_t oct_cells = 0;
    float* maj = nullptr;
    size_t maj_cells = 0;
    uint32_t* wtab_alloc = nullptr; // walk table: padded supergrid + exit masks, with `wtab_slack` border words on
    uint32_t* wtab = nullptr;       // either side (speculative look-ahead reads); wtab = cell 0 (Params::wtab)
    int wtab_slack = 0, wtab_words = 0;
    uint8_t* emask[2] = {nullptr, nullptr};  // ping-pong buffers of the exit-mask sweeps
    unsigned long long* counters = nullptr;
    unsigned int* work_counter = nullptr;
    unsigned int* debug = nullptr;  // [64] watchdog record of the slot-pool kernel
    int walk_limit = kPoolWalkLimit;  // (uivr_debug_set_walk_limit)
    // scratch of the backward pipeline: the reservoir records handed from the adjoint launch to the DRT launch
    uint32_t* records = nullptr;
    size_t records_cap = 0;
    uint4* desc = nullptr;          // vertex descriptors of the adjoint launch: [SM][slot][max_depth + 1][4]
    size_t desc_vecs = 0;
    float2* neelog = nullptr;       // NEE collision log of the adjoint launch: [SM][slot][kNeeLog + 1]
    float4* dalbedo4 = nullptr;     // (UIVR_DALBEDO_V4 builds) RGBA-padded d albedo
    size_t dalbedo4_vox = 0;
    float4* dsigma4 = nullptr;      // (UIVR_DSIGMA_TILED builds) d sigma_t in 2 x 2 tiles, four copies
    size_t dsigma4_tiles = 0;
    int variant = 3;
    // ray-batch mode (uivr_set_batch)
    bool batch_on = false;
    uivr_batch_desc batch{};
    float* d_sensors = nullptr;
    int d_sensors_cap = 0;
    // envmap emitter (uivr_set_envmap)
    bool env_on = false;
    uivr_envmap_desc env{};
    float4* d_env_data = nullptr;
    float* d_env_marg = nullptr;
    float* d_env_cond = nullptr;
    int counting = 0;
    uint64_t launches = 0;
    // staging for the *_host entry points
    float* st_sigma = nullptr;
    float* st_albedo = nullptr;
    float* st_image = nullptr;
    float* st_gimage = nullptr;
    float* st_dsigma = nullptr;
    float* st_dalbedo = nullptr;
    size_t st_vox = 0, st_pix = 0;
    bool st_params_valid = false;  // st_sigma / st_albedo hold the parameters of the last *_host call
    cudaStream_t copy_stream = nullptr;   // *_host: the albedo upload runs beside the supergrid rebuild
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    // CUDA events bracketing the most recent path kernel: [0] forward, [1] backward
    cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool ev_valid[2] = {false, false};
};

namespace {

int fail(uivr_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define UIVR_CUDA(ctx, call)                                                                   \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(ctx, UIVR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int check_ready(uivr_ctx* ctx) {
    if (!ctx) return UIVR_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, UIVR_ERR_STATE, "uivr_set_scene has not been called");
    if (!ctx->have_props) return fail(ctx, UIVR_ERR_STATE, "uivr_set_integrator has not been called");
    if (!ctx->have_medium) return fail(ctx, UIVR_ERR_STATE, "uivr_update_medium has not been called");
    return UIVR_OK;
}

// the slot-pool kernels pack depth into 16 bits and the padded supergrid cell index into 28 bits
bool pool_ok(const uivr_ctx* ctx) {
    return ctx->variant == 3 && ctx->props.max_depth < 65536 && ctx->mres[0] <= 512 && ctx->mres[1] <= 512 &&
           ctx->mres[2] <= 512;
}

int fill_params(uivr_ctx* ctx, Params& P, const uivr_shard* shard, uint32_t seed, int32_t spp, bool volpath = true) {
    const uivr_scene_desc& s = ctx->scene;
    const uivr_integrator_props& ip = ctx->props;
    if (spp < 1) return fail(ctx, UIVR_ERR_INVALID, "spp must be >= 1");
    const uint64_t npix = ctx->batch_on ? (uint64_t) ctx->batch.batch_size : (uint64_t) s.width * (uint64_t) s.height;
    if (npix * (uint64_t) spp >= (1ull << 32) - (1ull << 24))
        return fail(ctx, UIVR_ERR_INVALID, "wavefront too large: W*H*spp must be < 2^32 (batched.py:378-388)");
    memset(&P, 0, sizeof(P));
    for (int a = 0; a < 3; ++a) {
        P.res[a] = s.res[a];
        P.ores[a] = s.res[a] + 1;
        P.fres[a] = (float) s.res[a];
        P.mres[a] = ctx->mres[a];
        P.fmres[a] = (float) ctx->mres[a];
        P.mcs[a] = 1.0f / (float) ctx->mres[a];
        P.cam_origin[a] = s.cam_origin[a];
        P.cam_left[a] = s.cam_left[a];
        P.cam_up[a] = s.cam_up[a];
        P.cam_dir[a] = s.cam_dir[a];
        P.radiance[a] = s.radiance[a];
        P.half_le[a] = 0.5f * s.radiance[a];
    }
    memcpy(P.to_local, s.to_local, sizeof(P.to_local));
    P.oct = ctx->oct;
    P.maj = ctx->maj;
    P.wtab = ctx->wtab;
    for (int a = 0; a < 3; ++a) P.pm[a] = ctx->mres[a] + 2;
    P.wtab_slack = ctx->wtab_slack;
    P.wtab_words = ctx->wtab_words;
    P.scale = s.scale;
    P.tan_x = s.tan_x;
    P.tan_y = s.tan_y;
    P.near_clip = s.near_clip;
    P.width = ctx->batch_on ? ctx->batch.batch_size : s.width;
    P.height = ctx->batch_on ? 1 : s.height;
    P.inv_w = 1.0f / (float) P.width;
    P.inv_h = 1.0f / (float) P.height;
    if (ctx->batch_on) {
        if (volpath && !pool_ok(ctx))
            return fail(ctx, UIVR_ERR_INVALID, "ray-batch rendering needs the slot-pool kernels (variant 3)");
        P.sensors = ctx->d_sensors;
        P.n_sensors = ctx->batch.n_sensors;
        P.film_w = ctx->batch.film_w;
        P.film_h = ctx->batch.film_h;
        P.seed_pixels = uivr_tea32(ctx->batch.seed, 5);    // batched.py:409-413: sub_seed_i = tea32(seed, 17 i + 5)
        P.seed_offsets = uivr_tea32(ctx->batch.seed, 22);  // primal offsets; the backward entry switches to i = 2
    }
    if (ctx->env_on) {
        P.env_data = ctx->d_env_data;
        P.env_marg = ctx->d_env_marg;
        P.env_cond = ctx->d_env_cond;
        P.env_w = ctx->env.env_w;
        P.env_h = ctx->env.env_h;
        P.env_scale = ctx->env.scale;
        memcpy(P.env_to_world, ctx->env.env_to_world, sizeof(P.env_to_world));
        memcpy(P.world_to_env, ctx->env.world_to_env, sizeof(P.world_to_env));
        memcpy(P.local_to_world, ctx->env.local_to_world, sizeof(P.local_to_world));
    }
    P.max_depth = ip.max_depth;
    P.hide_emitters = ip.hide_emitters;
    P.use_nee = ip.use_nee;
    P.use_drt = ip.use_drt;
    P.use_drt_subsampling = ip.use_drt_subsampling;
    P.use_drt_mis = ip.use_drt_mis;
    P.seed = seed;
    P.spp = (uint32_t) spp;
    P.inv_spp = 1.0f / (float) spp;
    P.npix = (uint32_t) npix;
    if (shard && shard->count > 1) {
        if (shard->rank < 0 || shard->rank >= shard->count || shard->block < 1)
            return fail(ctx, UIVR_ERR_INVALID, "invalid shard (need 0 <= rank < count, block >= 1)");
        P.shard_rank = shard->rank;
        P.shard_count = shard->count;
        P.shard_block = shard->block;
        const uint64_t nblocks = (npix + shard->block - 1) / shard->block;
        const uint64_t mine = nblocks > (uint64_t) shard->rank
                                  ? (nblocks - shard->rank + shard->count - 1) / shard->count : 0;
        P.n_slots = (uint32_t) (mine * shard->block);
    } else {
        P.shard_rank = 0;
        P.shard_count = 1;
        P.shard_block = 1;
        P.n_slots = (uint32_t) npix;
    }
    P.counters = ctx->counters;
    P.work_counter = ctx->work_counter;
    P.debug = ctx->debug;
    P.walk_limit = ctx->walk_limit;
    {
        // chunked work fetch of the forward kernel: at least ~32 chunks per SM, so that the last chunk of the slowest
        // SM stays a small share of a short launch
        const uint64_t total = (uint64_t) P.n_slots * (uint64_t) P.spp;
        uint64_t c = total / ((uint64_t) ctx->num_sms * 32u);
        c = c / 32u * 32u;
        P.fetch_chunk = (int) (c < 32u ? 32u : (c > (uint64_t) UIVR_POOL_CHUNK_FWD ? (uint64_t) UIVR_POOL_CHUNK_FWD : c));
    }
    return UIVR_OK;
}

template <typename K>
int persistent_grid(uivr_ctx* ctx, K kernel, int block, int* grid) {
    int per_sm = 0;
    UIVR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0));
    if (per_sm < 1) per_sm = 1;
    *grid = ctx->num_sms * per_sm;
    return UIVR_OK;
}

int ensure_staging(uivr_ctx* ctx) {
    const uivr_scene_desc& s = ctx->scene;
    const size_t vox = (size_t) s.res[0] * s.res[1] * s.res[2];
    const size_t pix = (size_t) s.width * s.height;
    if (vox != ctx->st_vox) {
        cudaFree(ctx->st_sigma); cudaFree(ctx->st_albedo); cudaFree(ctx->st_dsigma); cudaFree(ctx->st_dalbedo);
        ctx->st_sigma = ctx->st_albedo = ctx->st_dsigma = ctx->st_dalbedo = nullptr;
        ctx->st_vox = 0;
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_sigma, vox * sizeof(float)));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_albedo, vox * 3 * sizeof(float)));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_dsigma, vox * sizeof(float)));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_dalbedo, vox * 3 * sizeof(float)));
        ctx->st_vox = vox;
    }
    if (pix != ctx->st_pix) {
        cudaFree(ctx->st_image); cudaFree(ctx->st_gimage);
        ctx->st_image = ctx->st_gimage = nullptr;
        ctx->st_pix = 0;
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_image, pix * 3 * sizeof(float)));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->st_gimage, pix * 3 * sizeof(float)));
        ctx->st_pix = pix;
    }
    return UIVR_OK;
}

}  // namespace

extern "C" int uivr_update_medium(uivr_ctx* ctx, const float* d_sigma_t, void* stream);

namespace {

// Upload of the parameters for the *_host entry points: sigma_t first, then the rebuild of everything derived from
// it (corner octets, majorant supergrid, walk table) on the caller's stream WHILE the albedo, three times the bytes,
// is still crossing the link on a second stream.
int stage_params(uivr_ctx* ctx, const float* h_sigma_t, const float* h_albedo, cudaStream_t st) {
    if (!ctx->copy_stream) {
        UIVR_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        UIVR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_ev[0], cudaEventDisableTiming));
        UIVR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_ev[1], cudaEventDisableTiming));
    }
    ctx->st_params_valid = false;
    UIVR_CUDA(ctx, cudaMemcpyAsync(ctx->st_sigma, h_sigma_t, ctx->st_vox * sizeof(float), cudaMemcpyHostToDevice, st));
    // the albedo copy follows the sigma_t copy (and whatever the caller's stream still does with the staging buffer)
    UIVR_CUDA(ctx, cudaEventRecord(ctx->copy_ev[0], st));
    UIVR_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
    UIVR_CUDA(ctx, cudaMemcpyAsync(ctx->st_albedo, h_albedo, ctx->st_vox * 3 * sizeof(float), cudaMemcpyHostToDevice,
                                   ctx->copy_stream));
    UIVR_CUDA(ctx, cudaEventRecord(ctx->copy_ev[1], ctx->copy_stream));
    const int rc = uivr_update_medium(ctx, ctx->st_sigma, (void*) st);
    UIVR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->copy_ev[1], 0));
    if (rc) return rc;
    ctx->st_params_valid = true;
    return UIVR_OK;
}

}  // namespace

extern "C" {

int uivr_version(void) { return 100; }

int uivr_create(int device, uivr_ctx** out) {
    if (!out) return UIVR_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return UIVR_ERR_CUDA;
    uivr_ctx* ctx = new (std::nothrow) uivr_ctx();
    if (!ctx) return UIVR_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaMalloc(&ctx->counters, sizeof(unsigned long long) * UIVR_NUM_COUNTERS) != cudaSuccess ||
        cudaMalloc(&ctx->work_counter, sizeof(unsigned int) * 4) != cudaSuccess ||
        cudaMalloc(&ctx->debug, sizeof(unsigned int) * 64) != cudaSuccess ||
        cudaMemset(ctx->debug, 0, sizeof(unsigned int) * 64) != cudaSuccess ||
        cudaMemset(ctx->counters, 0, sizeof(unsigned long long) * UIVR_NUM_COUNTERS) != cudaSuccess ||
        cudaEventCreate(&ctx->ev[0][0]) != cudaSuccess || cudaEventCreate(&ctx->ev[0][1]) != cudaSuccess ||
        cudaEventCreate(&ctx->ev[1][0]) != cudaSuccess || cudaEventCreate(&ctx->ev[1][1]) != cudaSuccess) {
        delete ctx;
        return UIVR_ERR_CUDA;
    }
    *out = ctx;
    return UIVR_OK;
}

int uivr_destroy(uivr_ctx* ctx) {
    if (!ctx) return UIVR_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->oct); cudaFree(ctx->maj); cudaFree(ctx->wtab_alloc); cudaFree(ctx->emask[0]); cudaFree(ctx->emask[1]); cudaFree(ctx->counters); cudaFree(ctx->work_counter); cudaFree(ctx->debug); cudaFree(ctx->records); cudaFree(ctx->desc); cudaFree(ctx->neelog); cudaFree(ctx->dalbedo4); cudaFree(ctx->dsigma4); cudaFree(ctx->d_sensors);
    cudaFree(ctx->d_env_data); cudaFree(ctx->d_env_marg); cudaFree(ctx->d_env_cond);
    cudaFree(ctx->st_sigma); cudaFree(ctx->st_albedo); cudaFree(ctx->st_image);
    cudaFree(ctx->st_gimage); cudaFree(ctx->st_dsigma); cudaFree(ctx->st_dalbedo);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (ctx->ev[i][j]) cudaEventDestroy(ctx->ev[i][j]);
    for (int i = 0; i < 2; ++i)
        if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return UIVR_OK;
}

const char* uivr_last_error(const uivr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int uivr_set_scene(uivr_ctx* ctx, const uivr_scene_desc* scene) {
    if (!ctx || !scene) return UIVR_ERR_INVALID;
    for (int a = 0; a < 3; ++a)
        if (scene->res[a] < 1 || scene->res[a] > 2047)
            return fail(ctx, UIVR_ERR_INVALID, "grid resolution must be in [1, 2047] per axis");
    if (scene->width < 1 || scene->height < 1) return fail(ctx, UIVR_ERR_INVALID, "film size must be positive");
    if (!(scene->scale >= 0.0f)) return fail(ctx, UIVR_ERR_INVALID, "medium scale must be >= 0");
    const bool res_changed = !ctx->have_scene || memcmp(ctx->scene.res, scene->res, sizeof(scene->res)) != 0 ||
                             ctx->scene.majorant_factor != scene->majorant_factor ||
                             ctx->scene.scale != scene->scale;
    ctx->scene = *scene;
    ctx->have_scene = true;
    if (res_changed) ctx->have_medium = false;  // derived layouts are stale
    return UIVR_OK;
}

int uivr_set_batch(uivr_ctx* ctx, const uivr_batch_desc* batch) {
    if (!ctx) return UIVR_ERR_INVALID;
    if (!batch) {
        ctx->batch_on = false;
        return UIVR_OK;
    }
    if (batch->n_sensors < 1 || !batch->sensors || batch->film_w < 1 || batch->film_h < 1 || batch->batch_size < 1)
        return fail(ctx, UIVR_ERR_INVALID, "invalid batch descriptor");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (batch->n_sensors > ctx->d_sensors_cap) {
        cudaFree(ctx->d_sensors);
        ctx->d_sensors = nullptr;
        ctx->d_sensors_cap = 0;
        UIVR_CUDA(ctx, cudaMalloc(&ctx->d_sensors, (size_t) batch->n_sensors * 16 * sizeof(float)));
        ctx->d_sensors_cap = batch->n_sensors;
    }
    // (synchronous copy: a handful of KB, and the host array need not outlive the call)
    UIVR_CUDA(ctx, cudaMemcpy(ctx->d_sensors, batch->sensors, (size_t) batch->n_sensors * 16 * sizeof(float),
                              cudaMemcpyHostToDevice));
    ctx->batch = *batch;
    ctx->batch.sensors = nullptr;
    ctx->batch_on = true;
    return UIVR_OK;
}

int uivr_set_envmap(uivr_ctx* ctx, const uivr_envmap_desc* env) {
    if (!ctx) return UIVR_ERR_INVALID;
    if (!env) {
        ctx->env_on = false;
        return UIVR_OK;
    }
    if (env->env_w < 1 || env->env_h < 2 || env->env_w > 16384 || env->env_h > 8192 || !env->data || !env->marg || !env->cond)
        return fail(ctx, UIVR_ERR_INVALID, "invalid envmap descriptor");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->env_on = false;
    cudaFree(ctx->d_env_data); cudaFree(ctx->d_env_marg); cudaFree(ctx->d_env_cond);
    ctx->d_env_data = nullptr; ctx->d_env_marg = ctx->d_env_cond = nullptr;
    const size_t nv = (size_t) env->env_h * (env->env_w + 1), nr = (size_t) env->env_h - 1, nc = nr * env->env_w;
    UIVR_CUDA(ctx, cudaMalloc(&ctx->d_env_data, nv * sizeof(float4)));
    UIVR_CUDA(ctx, cudaMalloc(&ctx->d_env_marg, nr * sizeof(float)));
    UIVR_CUDA(ctx, cudaMalloc(&ctx->d_env_cond, nc * sizeof(float)));
    // (synchronous copies: set once per scene, and the host arrays need not outlive the call)
    UIVR_CUDA(ctx, cudaMemcpy(ctx->d_env_data, env->data, nv * sizeof(float4), cudaMemcpyHostToDevice));
    UIVR_CUDA(ctx, cudaMemcpy(ctx->d_env_marg, env->marg, nr * sizeof(float), cudaMemcpyHostToDevice));
    UIVR_CUDA(ctx, cudaMemcpy(ctx->d_env_cond, env->cond, nc * sizeof(float), cudaMemcpyHostToDevice));
    ctx->env = *env;
    ctx->env.data = ctx->env.marg = ctx->env.cond = nullptr;
    ctx->env_on = true;
    return UIVR_OK;
}

int uivr_set_integrator(uivr_ctx* ctx, const uivr_integrator_props* props) {
    if (!ctx || !props) return UIVR_ERR_INVALID;
    if (props->max_depth < 0) return fail(ctx, UIVR_ERR_INVALID, "max_depth must be >= 0 (opt_config.py:102)");
    ctx->props = *props;
    ctx->have_props = true;
    return UIVR_OK;
}

int uivr_set_counting(uivr_ctx* ctx, int enable) {
    if (!ctx) return UIVR_ERR_INVALID;
    ctx->counting = enable ? 1 : 0;
    return UIVR_OK;
}

int uivr_debug_set_walk_limit(uivr_ctx* ctx, int limit) {
    if (!ctx) return UIVR_ERR_INVALID;
    ctx->walk_limit = limit > 0 ? limit : kPoolWalkLimit;
    return UIVR_OK;
}

int uivr_set_variant(uivr_ctx* ctx, int variant) {
    if (!ctx || (variant != 1 && variant != 3)) return UIVR_ERR_INVALID;
    ctx->variant = variant;
    return UIVR_OK;
}

int uivr_reset_counters(uivr_ctx* ctx, void* stream) {
    if (!ctx) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long) * UIVR_NUM_COUNTERS, (cudaStream_t) stream));
    return UIVR_OK;
}

int uivr_get_counters(uivr_ctx* ctx, uint64_t out[UIVR_NUM_COUNTERS], void* stream) {
    if (!ctx || !out) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    UIVR_CUDA(ctx, cudaMemcpyAsync(out, ctx->counters, sizeof(uint64_t) * UIVR_NUM_COUNTERS, cudaMemcpyDeviceToHost,
                                   (cudaStream_t) stream));
    UIVR_CUDA(ctx, cudaStreamSynchronize((cudaStream_t) stream));
    return UIVR_OK;
}

int uivr_get_kernel_ms(uivr_ctx* ctx, int which, float* ms) {
    if (!ctx || !ms || which < 0 || which > 1) return UIVR_ERR_INVALID;
    if (!ctx->ev_valid[which]) return fail(ctx, UIVR_ERR_STATE, "no path kernel of that kind has been launched");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    UIVR_CUDA(ctx, cudaEventSynchronize(ctx->ev[which][1]));
    UIVR_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
    return UIVR_OK;
}

int uivr_check_watchdog(uivr_ctx* ctx, uint32_t out[64], void* stream) {
    if (!ctx) return UIVR_ERR_INVALID;
    uint32_t rec[64];
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    UIVR_CUDA(ctx, cudaMemcpyAsync(rec, ctx->debug, sizeof(rec), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    UIVR_CUDA(ctx, cudaStreamSynchronize((cudaStream_t) stream));
    if (out) memcpy(out, rec, sizeof(rec));
    if (rec[0] == 0u) return UIVR_OK;
    char msg[512];
    int n = snprintf(msg, sizeof(msg), "slot-pool kernel watchdog tripped: reason 0x%x, block %u thread %u; queues (head,tail,count):",
                     rec[0], rec[1], rec[2]);
    for (int q = 0; q < Q_NUM && n < (int) sizeof(msg) - 40; ++q)
        n += snprintf(msg + n, sizeof(msg) - n, " q%d(%u,%u,%d)", q, rec[4 + 3 * q], rec[5 + 3 * q], (int) rec[6 + 3 * q]);
    snprintf(msg + n, sizeof(msg) - n, " live %d exhausted %u", (int) rec[4 + 3 * Q_NUM], rec[5 + 3 * Q_NUM]);
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->debug, 0, sizeof(rec), (cudaStream_t) stream));
    return fail(ctx, UIVR_ERR_WATCHDOG, msg);
}

int uivr_get_launch_count(const uivr_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return UIVR_ERR_INVALID;
    *out = ctx->launches;
    return UIVR_OK;
}

int uivr_update_medium(uivr_ctx* ctx, const float* d_sigma_t, void* stream) {
    if (!ctx || !d_sigma_t) return UIVR_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, UIVR_ERR_STATE, "uivr_set_scene has not been called");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    const uivr_scene_desc& s = ctx->scene;
    const size_t cells = (size_t) (s.res[0] + 1) * (s.res[1] + 1) * (s.res[2] + 1);
    if (cells != ctx->oct_cells) {
        cudaFree(ctx->oct);
        ctx->oct = nullptr;
        ctx->oct_cells = 0;
        UIVR_CUDA(ctx, cudaMalloc(&ctx->oct, cells * 2 * sizeof(float4)));
        ctx->oct_cells = cells;
    }
    for (int a = 0; a < 3; ++a) {
        ctx->mres[a] = (s.majorant_factor > 1) ? s.res[a] / s.majorant_factor : 1;
        if (ctx->mres[a] < 1) ctx->mres[a] = 1;
    }
    const size_t mcells = (size_t) ctx->mres[0] * ctx->mres[1] * ctx->mres[2];
    const int mx = ctx->mres[0], my = ctx->mres[1], mz = ctx->mres[2];
    const size_t pcells = (size_t) (mx + 2) * (my + 2) * (mz + 2);
    if (mcells != ctx->maj_cells) {
        cudaFree(ctx->maj); cudaFree(ctx->wtab_alloc); cudaFree(ctx->emask[0]); cudaFree(ctx->emask[1]);
        ctx->maj = nullptr; ctx->wtab = ctx->wtab_alloc = nullptr; ctx->emask[0] = ctx->emask[1] = nullptr;
        ctx->maj_cells = 0;
        ctx->wtab_slack = ((mx + 2) * (my + 2) + 3) / 4 * 4;   // (a multiple of 4 words keeps cell 0 16-byte aligned)
        ctx->wtab_words = (int) ((pcells + 2 * (size_t) ctx->wtab_slack + 3) / 4 * 4);
        UIVR_CUDA(ctx, cudaMalloc(&ctx->maj, mcells * sizeof(float)));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->wtab_alloc, (size_t) ctx->wtab_words * sizeof(uint32_t)));
        UIVR_CUDA(ctx, cudaMemsetAsync(ctx->wtab_alloc, 0xFF, (size_t) ctx->wtab_words * sizeof(uint32_t), st));
        ctx->wtab = ctx->wtab_alloc + ctx->wtab_slack;
        UIVR_CUDA(ctx, cudaMalloc(&ctx->emask[0], mcells));
        UIVR_CUDA(ctx, cudaMalloc(&ctx->emask[1], mcells));
        ctx->maj_cells = mcells;
    }
    const int grid = ctx->num_sms * 8;
    k_build_octets<<<grid, kBlock, 0, st>>>(d_sigma_t, ctx->oct, s.res[0], s.res[1], s.res[2]);
    k_build_majorant<<<grid, kBlock, 0, st>>>(d_sigma_t, ctx->maj, s.res[0], s.res[1], s.res[2], mx, my, mz, s.scale);
    // exit mask: separable AND sweeps along x, y, z (one thread per grid line), then the padded walk table
    const size_t sxy = (size_t) mx * my;
    k_exit_sweep<<<(my * mz + kBlock - 1) / kBlock, kBlock, 0, st>>>(ctx->maj, nullptr, ctx->emask[0], mx, 1, my, (size_t) mx, mz, sxy, 1);
    k_exit_sweep<<<(mx * mz + kBlock - 1) / kBlock, kBlock, 0, st>>>(ctx->maj, ctx->emask[0], ctx->emask[1], my, (size_t) mx, mx, 1, mz, sxy, 2);
    k_exit_sweep<<<(mx * my + kBlock - 1) / kBlock, kBlock, 0, st>>>(ctx->maj, ctx->emask[1], ctx->emask[0], mz, sxy, mx, 1, my, (size_t) mx, 4);
    k_build_walk_table<<<grid, kBlock, 0, st>>>(ctx->maj, ctx->emask[0], ctx->wtab, mx, my, mz, ctx->wtab_slack);
    ctx->launches += 6;
    UIVR_CUDA(ctx, cudaGetLastError());
    ctx->have_medium = true;
    return UIVR_OK;
}

int uivr_render_forward(uivr_ctx* ctx, const float* d_albedo, uint32_t seed, int32_t spp, const uivr_shard* shard,
                        float* d_image, float* d_sample_L, void* stream) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!d_albedo || !d_image) return fail(ctx, UIVR_ERR_INVALID, "null device pointer");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    Params P;
    if ((rc = fill_params(ctx, P, shard, seed, spp))) return rc;
    if (ctx->batch_on && seed != ctx->batch.seed)
        return fail(ctx, UIVR_ERR_INVALID, "ray-batch mode: the forward seed must be the seed of uivr_set_batch");
    P.albedo = d_albedo;
    P.image = d_image;
    P.sample_L = d_sample_L;
    const size_t nimg = (size_t) P.npix * 3;
    UIVR_CUDA(ctx, cudaMemsetAsync(d_image, 0, nimg * sizeof(float), st));
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int) * 4, st));
    int grid = 0;
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[0][0], st));
    if (!pool_ok(ctx)) {  // variant 1, or a scene outside the slot-pool kernels' packing limits
        if (ctx->counting) {
            if ((rc = persistent_grid(ctx, k_forward_v1<true>, kBlock, &grid))) return rc;
            k_forward_v1<true><<<grid, kBlock, 0, st>>>(P);
        } else {
            if ((rc = persistent_grid(ctx, k_forward_v1<false>, kBlock, &grid))) return rc;
            k_forward_v1<false><<<grid, kBlock, 0, st>>>(P);
        }
    } else {
        if ((rc = launch_pool(ctx->num_sms, KIND_FWD, ctx->counting != 0, P, st))) return fail(ctx, rc, "pool kernel launch failed");
    }
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[0][1], st));
    ctx->ev_valid[0] = true;
    k_scale<<<ctx->num_sms * 4, kBlock, 0, st>>>(d_image, nimg, P.inv_spp);
    ctx->launches += 2;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_render_backward(uivr_ctx* ctx, const float* d_albedo, const float* d_grad_image, uint32_t seed_grad,
                         int32_t spp_grad, const uivr_shard* shard, float* d_dsigma_t, float* d_dalbedo,
                         float* d_sample_L, void* stream) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!d_albedo || !d_grad_image || !d_dsigma_t || !d_dalbedo) return fail(ctx, UIVR_ERR_INVALID, "null device pointer");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    Params P;
    if ((rc = fill_params(ctx, P, shard, seed_grad, spp_grad))) return rc;
    P.alt_seed = ctx->batch_on ? uivr_alt_seed_batch(seed_grad) : uivr_alt_seed(seed_grad);
    if (ctx->batch_on) P.seed_offsets = uivr_tea32(ctx->batch.seed, 39);  // decorrelated offsets, same pixels (batched.py:69-75)
    P.albedo = d_albedo;
    P.grad_image = d_grad_image;
    P.dsigma = d_dsigma_t;
    P.dalbedo = d_dalbedo;
    P.sample_L = d_sample_L;
    const size_t vox = (size_t) P.res[0] * P.res[1] * P.res[2];
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int) * 4, st));
    int grid = 0;
    // the O(n^2) mode (use_drt_subsampling = False) nests sub-paths: served by variant 1
    const bool quadratic = ctx->props.use_drt && !ctx->props.use_drt_subsampling;
    if (ctx->batch_on && quadratic)
        return fail(ctx, UIVR_ERR_INVALID, "ray-batch rendering is not available for use_drt_subsampling = False "
                                           "(the O(n^2) mode runs on the one-sample-per-lane kernels, which generate sensor rays only)");
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[1][0], st));
    // (the slot-pool adjoint keeps max_depth + 1 vertex descriptors per in-flight sample: deeper paths than 255
    // vertices go to the one-sample-per-lane kernels)
    // The slot-pool kernels scatter into accumulation buffers of their own (tiles / RGBA) that are folded into the
    // caller's tensors at the end: nothing else writes those, so they need no zeroing there.
    const bool pool_path = !(quadratic || !pool_ok(ctx) || ctx->props.max_depth > 255);
    const bool sigma_folded = pool_path && UIVR_DSIGMA_TILED && !UIVR_SCATTER_MATCH;
    const bool albedo_folded = pool_path && UIVR_DALBEDO_V4;
    if (!sigma_folded) UIVR_CUDA(ctx, cudaMemsetAsync(d_dsigma_t, 0, vox * sizeof(float), st));
    if (!albedo_folded) UIVR_CUDA(ctx, cudaMemsetAsync(d_dalbedo, 0, vox * 3 * sizeof(float), st));
    if (!pool_path) {
        if (ctx->counting) {
            if ((rc = persistent_grid(ctx, k_backward_v1<true>, kBlock, &grid))) return rc;
            k_backward_v1<true><<<grid, kBlock, 0, st>>>(P);
        } else {
            if ((rc = persistent_grid(ctx, k_backward_v1<false>, kBlock, &grid))) return rc;
            k_backward_v1<false><<<grid, kBlock, 0, st>>>(P);
        }
    } else {
        // pipeline: adjoint replay (gathers the primal radiance itself, scatters the free-flight / transmittance /
        // NEE gradients, reservoir records to HBM) -> DRT pass.  Work counters: [1] adjoint, [2] DRT, [3] records.
        const bool drt_pass = ctx->props.use_drt != 0;
        const size_t n_local = (size_t) P.n_slots * P.spp;
        if (drt_pass && ctx->records_cap < n_local) {
            cudaFree(ctx->records);
            ctx->records = nullptr;
            ctx->records_cap = 0;
            UIVR_CUDA(ctx, cudaMalloc(&ctx->records, n_local * kRecWords * sizeof(uint32_t)));
            ctx->records_cap = n_local;
        }
        const size_t desc_vecs = (size_t) ctx->num_sms * UIVR_POOL_SLOTS_ADJ * (size_t) (ctx->props.max_depth + 1) * kDescVec;
        if (ctx->desc_vecs < desc_vecs) {
            cudaFree(ctx->desc);
            ctx->desc = nullptr;
            ctx->desc_vecs = 0;
            UIVR_CUDA(ctx, cudaMalloc(&ctx->desc, desc_vecs * sizeof(uint4)));
            ctx->desc_vecs = desc_vecs;
        }
        if (!ctx->neelog)
            UIVR_CUDA(ctx, cudaMalloc(&ctx->neelog, (size_t) ctx->num_sms * UIVR_POOL_SLOTS_ADJ * kNeeLogStride * sizeof(float2)));
        P.desc = ctx->desc;
        P.desc_cap = ctx->props.max_depth + 1;
        P.neelog = ctx->neelog;
#if UIVR_DSIGMA_TILED
        {
            P.tile_x = (P.res[0] + 1) >> 1;
            P.tile_y = (P.res[1] + 1) >> 1;
#if UIVR_DSIGMA_TILED == 2
            P.tile_z = (P.res[2] + 1) >> 1;
            const size_t tiles = (size_t) 16 * P.tile_z * P.tile_y * P.tile_x;
#else
            P.tile_z = P.res[2];
            const size_t tiles = (size_t) 4 * P.res[2] * P.tile_y * P.tile_x;
#endif
            if (ctx->dsigma4_tiles < tiles) {
                cudaFree(ctx->dsigma4);
                ctx->dsigma4 = nullptr;
                ctx->dsigma4_tiles = 0;
                UIVR_CUDA(ctx, cudaMalloc(&ctx->dsigma4, tiles * sizeof(float4)));
                ctx->dsigma4_tiles = tiles;
            }
            UIVR_CUDA(ctx, cudaMemsetAsync(ctx->dsigma4, 0, tiles * sizeof(float4), st));
            P.dsigma4 = ctx->dsigma4;
        }
#endif
#if UIVR_DALBEDO_V4
        if (ctx->dalbedo4_vox < vox) {
            cudaFree(ctx->dalbedo4);
            ctx->dalbedo4 = nullptr;
            ctx->dalbedo4_vox = 0;
            UIVR_CUDA(ctx, cudaMalloc(&ctx->dalbedo4, vox * sizeof(float4)));
            ctx->dalbedo4_vox = vox;
        }
        UIVR_CUDA(ctx, cudaMemsetAsync(ctx->dalbedo4, 0, vox * sizeof(float4), st));
        P.dalbedo4 = ctx->dalbedo4;
#endif
        P.records = ctx->records;
        P.rec_count = ctx->work_counter + 3;
        P.work_counter = ctx->work_counter + 1;
        if ((rc = launch_pool(ctx->num_sms, KIND_ADJ, ctx->counting != 0, P, st))) return fail(ctx, rc, "pool kernel launch failed");
        if (drt_pass) {
            P.work_counter = ctx->work_counter + 2;
            if ((rc = launch_pool(ctx->num_sms, KIND_DRT, ctx->counting != 0, P, st))) return fail(ctx, rc, "pool kernel launch failed");
            ctx->launches += 1;
        }
#if UIVR_DSIGMA_TILED
#if UIVR_DSIGMA_TILED == 2
        k_tiles3_to_dsigma<<<ctx->num_sms * 8, kBlock, 0, st>>>(ctx->dsigma4, d_dsigma_t, P.res[0], P.res[1], P.res[2], P.tile_x, P.tile_y, P.tile_z);
#else
        k_tiles_to_dsigma<<<ctx->num_sms * 8, kBlock, 0, st>>>(ctx->dsigma4, d_dsigma_t, P.res[0], P.res[1], P.res[2], P.tile_x, P.tile_y);
#endif
        ctx->launches += 1;
#endif
#if UIVR_DALBEDO_V4
        k_rgba_to_rgb<<<ctx->num_sms * 8, kBlock, 0, st>>>(ctx->dalbedo4, d_dalbedo, vox);
        ctx->launches += 1;
#endif
    }
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[1][1], st));
    ctx->ev_valid[1] = true;
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

// ---------------------------------------------------------------------------------------
// `nerf` integrator (python/integrators/nerf.py)
// ---------------------------------------------------------------------------------------
static int nerf_params(uivr_ctx* ctx, const uivr_nerf_props* np, Params& P, const uivr_shard* shard, uint32_t seed,
                       int32_t spp) {
    if (!ctx) return UIVR_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, UIVR_ERR_STATE, "uivr_set_scene has not been called");
    if (!ctx->have_medium) return fail(ctx, UIVR_ERR_STATE, "uivr_update_medium has not been called");
    if (!np) return fail(ctx, UIVR_ERR_INVALID, "null nerf properties");
    if (np->queries_per_ray < 2) return fail(ctx, UIVR_ERR_INVALID, "queries_per_ray must be >= 2");
    if (np->activation != UIVR_NERF_IDENTITY && np->activation != UIVR_NERF_RELU)
        return fail(ctx, UIVR_ERR_INVALID, "Unsupported activation (nerf.py:44)");
    const int rc = fill_params(ctx, P, shard, seed, spp, false);
    if (rc) return rc;
    P.nerf_queries = np->queries_per_ray;
    P.nerf_jitter = np->jittering_enabled ? 1 : 0;
    P.nerf_activation = np->activation;
    P.hide_emitters = np->hide_emitters ? 1 : 0;
    return UIVR_OK;
}

int uivr_nerf_forward(uivr_ctx* ctx, const uivr_nerf_props* props, const float* d_emission, uint32_t seed, int32_t spp,
                      const uivr_shard* shard, float* d_image, float* d_sample_L, void* stream) {
    Params P;
    int rc = nerf_params(ctx, props, P, shard, seed, spp);
    if (rc) return rc;
    if (!d_emission || !d_image) return fail(ctx, UIVR_ERR_INVALID, "null device pointer");
    if (ctx->batch_on && seed != ctx->batch.seed)
        return fail(ctx, UIVR_ERR_INVALID, "ray-batch mode: the forward seed must be the seed of uivr_set_batch");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    P.albedo = d_emission;
    P.image = d_image;
    P.sample_L = d_sample_L;
    const size_t nimg = (size_t) P.npix * 3;
    UIVR_CUDA(ctx, cudaMemsetAsync(d_image, 0, nimg * sizeof(float), st));
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int) * 4, st));
    int grid = 0;
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[0][0], st));
    if (ctx->counting) {
        if ((rc = persistent_grid(ctx, k_nerf_forward<true>, kBlock, &grid))) return rc;
        k_nerf_forward<true><<<grid, kBlock, 0, st>>>(P);
    } else {
        if ((rc = persistent_grid(ctx, k_nerf_forward<false>, kBlock, &grid))) return rc;
        k_nerf_forward<false><<<grid, kBlock, 0, st>>>(P);
    }
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[0][1], st));
    ctx->ev_valid[0] = true;
    k_scale<<<ctx->num_sms * 4, kBlock, 0, st>>>(d_image, nimg, P.inv_spp);
    ctx->launches += 2;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_nerf_backward(uivr_ctx* ctx, const uivr_nerf_props* props, const float* d_emission, const float* d_grad_image,
                       uint32_t seed_grad, int32_t spp_grad, const uivr_shard* shard, float* d_dsigma_t,
                       float* d_demission, float* d_sample_L, void* stream) {
    Params P;
    int rc = nerf_params(ctx, props, P, shard, seed_grad, spp_grad);
    if (rc) return rc;
    if (!d_emission || !d_grad_image || !d_dsigma_t || !d_demission) return fail(ctx, UIVR_ERR_INVALID, "null device pointer");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    if (ctx->batch_on) P.seed_offsets = uivr_tea32(ctx->batch.seed, 39);  // decorrelated offsets (batched.py:69-75)
    P.albedo = d_emission;
    P.grad_image = d_grad_image;
    P.dsigma = d_dsigma_t;
    P.dalbedo = d_demission;
    P.sample_L = d_sample_L;
    const size_t vox = (size_t) P.res[0] * P.res[1] * P.res[2];
    UIVR_CUDA(ctx, cudaMemsetAsync(d_dsigma_t, 0, vox * sizeof(float), st));
    UIVR_CUDA(ctx, cudaMemsetAsync(d_demission, 0, vox * 3 * sizeof(float), st));
    UIVR_CUDA(ctx, cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int) * 4, st));
    int grid = 0;
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[1][0], st));
    if (ctx->counting) {
        if ((rc = persistent_grid(ctx, k_nerf_backward<true>, kBlock, &grid))) return rc;
        k_nerf_backward<true><<<grid, kBlock, 0, st>>>(P);
    } else {
        if ((rc = persistent_grid(ctx, k_nerf_backward<false>, kBlock, &grid))) return rc;
        k_nerf_backward<false><<<grid, kBlock, 0, st>>>(P);
    }
    UIVR_CUDA(ctx, cudaEventRecord(ctx->ev[1][1], st));
    ctx->ev_valid[1] = true;
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_test_atan2_turns(uivr_ctx* ctx, const float* d_y, const float* d_x, int n, float* d_out, void* stream) {
    if (!ctx || !d_y || !d_x || !d_out || n < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > 0) k_test_atan2_turns<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(d_y, d_x, n, d_out);
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_test_exp(uivr_ctx* ctx, const float* d_x, int n, float* d_out, void* stream) {
    if (!ctx || !d_x || !d_out || n < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n > 0) k_test_exp<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(d_x, n, d_out);
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_render_forward_host(uivr_ctx* ctx, const float* h_sigma_t, const float* h_albedo, uint32_t seed, int32_t spp,
                             const uivr_shard* shard, float* h_image, void* stream) {
    if (!ctx || !h_sigma_t || !h_albedo || !h_image) return UIVR_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, UIVR_ERR_STATE, "uivr_set_scene has not been called");
    if (ctx->batch_on) return fail(ctx, UIVR_ERR_STATE, "the *_host entry points render a sensor, not a ray batch");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    int rc = ensure_staging(ctx);
    if (rc) return rc;
    if ((rc = stage_params(ctx, h_sigma_t, h_albedo, st))) return rc;
    if ((rc = uivr_render_forward(ctx, ctx->st_albedo, seed, spp, shard, ctx->st_image, nullptr, stream))) return rc;
    UIVR_CUDA(ctx, cudaMemcpyAsync(h_image, ctx->st_image, ctx->st_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    UIVR_CUDA(ctx, cudaStreamSynchronize(st));
    return UIVR_OK;
}

int uivr_render_backward_host(uivr_ctx* ctx, const float* h_sigma_t, const float* h_albedo, const float* h_grad_image,
                              uint32_t seed_grad, int32_t spp_grad, const uivr_shard* shard, float* h_dsigma_t,
                              float* h_dalbedo, void* stream) {
    if (!ctx || !h_grad_image || !h_dsigma_t || !h_dalbedo || (!h_sigma_t != !h_albedo)) return UIVR_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, UIVR_ERR_STATE, "uivr_set_scene has not been called");
    if (ctx->batch_on) return fail(ctx, UIVR_ERR_STATE, "the *_host entry points render a sensor, not a ray batch");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t) stream;
    const size_t old_vox = ctx->st_vox;
    int rc = ensure_staging(ctx);
    if (rc) return rc;
    if (h_sigma_t) {
        if ((rc = stage_params(ctx, h_sigma_t, h_albedo, st))) return rc;
    } else if (!ctx->st_params_valid || old_vox != ctx->st_vox || !ctx->have_medium) {
        return fail(ctx, UIVR_ERR_STATE, "no staged parameters: pass h_sigma_t / h_albedo, or call uivr_render_forward_host first");
    }
    UIVR_CUDA(ctx, cudaMemcpyAsync(ctx->st_gimage, h_grad_image, ctx->st_pix * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = uivr_render_backward(ctx, ctx->st_albedo, ctx->st_gimage, seed_grad, spp_grad, shard, ctx->st_dsigma,
                                   ctx->st_dalbedo, nullptr, stream)))
        return rc;
    UIVR_CUDA(ctx, cudaMemcpyAsync(h_dsigma_t, ctx->st_dsigma, ctx->st_vox * sizeof(float), cudaMemcpyDeviceToHost, st));
    UIVR_CUDA(ctx, cudaMemcpyAsync(h_dalbedo, ctx->st_dalbedo, ctx->st_vox * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    UIVR_CUDA(ctx, cudaStreamSynchronize(st));
    return UIVR_OK;
}

int uivr_adam_step(uivr_ctx* ctx, float* d_param, const float* d_grad, float* d_m, float* d_v, uint64_t n, float lr,
                   float beta1, float beta2, float eps, int32_t t, float lo, float hi, void* stream) {
    if (!ctx || !d_param || !d_grad || !d_m || !d_v) return UIVR_ERR_INVALID;
    if (t < 1) return fail(ctx, UIVR_ERR_INVALID, "Adam step counter t starts at 1");
    if (!(lo <= hi)) return fail(ctx, UIVR_ERR_INVALID, "clip range must satisfy lo <= hi");
    if ((((uintptr_t) d_param | (uintptr_t) d_grad | (uintptr_t) d_m | (uintptr_t) d_v) & 15u) != 0)
        return fail(ctx, UIVR_ERR_INVALID, "Adam tensors must be 16-byte aligned");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    const double b1t = pow((double) beta1, (double) t), b2t = pow((double) beta2, (double) t);
    const float step = (float) ((double) lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    if (n) k_adam_step<<<ctx->num_sms * 8, kBlock, 0, (cudaStream_t) stream>>>(d_param, d_grad, d_m, d_v, (size_t) n, step, beta1,
                                                                             beta2, eps, lo, hi);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_upsample2x(uivr_ctx* ctx, const float* d_in, const int32_t res[3], int32_t channels, float* d_out, void* stream) {
    if (!ctx || !d_in || !d_out || !res) return UIVR_ERR_INVALID;
    if (res[0] < 1 || res[1] < 1 || res[2] < 1 || channels < 1) return fail(ctx, UIVR_ERR_INVALID, "invalid grid shape");
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    k_upsample2x<<<ctx->num_sms * 8, kBlock, 0, (cudaStream_t) stream>>>(d_in, d_out, res[0], res[1], res[2], channels);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

// ---- primitives ----

uint32_t uivr_tea32(uint32_t v0, uint32_t v1) {
    uint32_t a, b;
    tea(v0, v1, a, b);
    return a;
}

static uint32_t alt_seed_after(uint32_t seed_grad, int skipped) {
    // volpathsimple.py:99-107: bits of lane 0's `alt_seed_rnd` (the sampler float that follows the
    // `skipped` draws made before :99), scrambled by TEA(.,1)
    Rng r;
    r.seed_sampler(seed_grad, 0);
    for (int i = 0; i < skipped; ++i) r.next();
    const uint32_t x = r.next();
    union { uint32_t u; float f; } c;
    c.u = (x >> 9) | 0x3f800000u;
    c.f = c.f - 1.0f;
    return uivr_tea32(c.u, 1);
}

// mi.render: 2 jitter draws + the burned draw of :71 precede alt_seed_rnd
uint32_t uivr_alt_seed(uint32_t seed_grad) { return alt_seed_after(seed_grad, 3); }
// render_batch: the path sampler draws no jitter (batched.py:390, :437), only :71 precedes it
uint32_t uivr_alt_seed_batch(uint32_t seed_grad) { return alt_seed_after(seed_grad, 1); }

int uivr_test_neg_log1m(uivr_ctx* ctx, const float* d_u, int n, float* d_out, void* stream) {
    if (!ctx || !d_u || !d_out || n < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n) k_test_neg_log1m<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(d_u, n, d_out);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_test_sincos2pi(uivr_ctx* ctx, const float* d_x, int n, float* d_s, float* d_c, void* stream) {
    if (!ctx || !d_x || !d_s || !d_c || n < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n) k_test_sincos2pi<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(d_x, n, d_s, d_c);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_test_sampler(uivr_ctx* ctx, uint32_t seed, uint32_t idx0, int nstreams, int ndraws, float* d_out, void* stream) {
    if (!ctx || !d_out || nstreams < 0 || ndraws < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (nstreams) k_test_sampler<<<(nstreams + 255) / 256, 256, 0, (cudaStream_t) stream>>>(seed, idx0, nstreams, ndraws, d_out);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_test_sigma_lookup(uivr_ctx* ctx, const float* d_p, int n, float* d_out, void* stream) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!d_p || !d_out || n < 0) return UIVR_ERR_INVALID;
    UIVR_CUDA(ctx, cudaSetDevice(ctx->device));
    Params P;
    if ((rc = fill_params(ctx, P, nullptr, 0, 1))) return rc;
    if (n) k_test_sigma_lookup<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(P, d_p, n, d_out);
    ctx->launches += 1;
    UIVR_CUDA(ctx, cudaGetLastError());
    return UIVR_OK;
}

int uivr_get_majorant(uivr_ctx* ctx, int32_t mres[3], float* d_out, void* stream) {
    if (!ctx || !mres) return UIVR_ERR_INVALID;
    if (!ctx->have_medium) return fail(ctx, UIVR_ERR_STATE, "uivr_update_medium has not been called");
    for (int a = 0; a < 3; ++a) mres[a] = ctx->mres[a];
    if (d_out)
        UIVR_CUDA(ctx, cudaMemcpyAsync(d_out, ctx->maj, ctx->maj_cells * sizeof(float), cudaMemcpyDeviceToDevice,
                                       (cudaStream_t) stream));
    return UIVR_OK;
}

int uivr_get_walk_table(uivr_ctx* ctx, int32_t mres[3], uint32_t* d_out, void* stream) {
    if (!ctx || !mres) return UIVR_ERR_INVALID;
    if (!ctx->have_medium) return fail(ctx, UIVR_ERR_STATE, "uivr_update_medium has not been called");
    for (int a = 0; a < 3; ++a) mres[a] = ctx->mres[a];
    const size_t pcells = (size_t) (ctx->mres[0] + 2) * (ctx->mres[1] + 2) * (ctx->mres[2] + 2);
    if (d_out)
        UIVR_CUDA(ctx, cudaMemcpyAsync(d_out, ctx->wtab, pcells * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                       (cudaStream_t) stream));
    return UIVR_OK;
}

}  // extern "C"
