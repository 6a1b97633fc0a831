// uivr_device.cuh -- device-side primitives of the volpathsimple hot path (sm_100a).
//
// Arithmetic contract: only IEEE + - * / sqrt, explicit fmaf and integer ops;
// the translation unit is compiled with -fmad=false so nvcc never fuses a*b+c on its own.
// The CPU oracle follows the same operation order, which makes every branch decision of a
// path (collision accept/reject, escape tests, reservoir swaps) reproducible bit for bit.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace uivr {

#define UIVR_DEV __device__ __forceinline__
#define UIVR_INF __int_as_float(0x7f800000)
#define UIVR_ENTRY_EPS 1e-4f

// --------------------------------------------------------------------------------------
// Parameters shared by all kernels (passed by value -> constant bank)
// --------------------------------------------------------------------------------------
struct Params {
    // medium
    int   res[3];            // X, Y, Z
    int   ores[3];           // octet grid dims = res + 1
    float fres[3];           // (float) res
    const float4* __restrict__ oct;     // corner octets: 2 x float4 per cell, ((z*OY+y)*OX+x)
    const float*  __restrict__ albedo;  // (Z,Y,X,3)
    const float*  __restrict__ maj;     // supergrid (MZ,MY,MX)
    // walk table: the supergrid with a one-cell border, (MZ+2,MY+2,MX+2), one word per cell.  Non-empty cell:
    // the bits of its majorant (> 0).  Empty cell: kWalkEmpty | exit mask (bit o: only empty cells ahead in
    // octant o; octant bit a = direction negative along axis a).  Border: kWalkBorder (ends every walk).
    const uint32_t* __restrict__ wtab;
    int   pm[3];             // padded dims = mres + 2
    int   wtab_slack;        // border words in front of cell 0 (and behind the last cell)
    int   wtab_words;        // words of the whole allocation (a multiple of 4)
    int   mres[3];
    float fmres[3];          // (float) mres
    float mcs[3];            // 1 / mres
    float scale;
    float to_local[12];
    // sensor / film / emitter
    float cam_origin[3], cam_left[3], cam_up[3], cam_dir[3];
    float tan_x, tan_y, near_clip;
    int   width, height;
    float inv_w, inv_h;
    float radiance[3], half_le[3];
    // ray-batch mode (python/batched.py): the wavefront is a batch of (sensor, pixel) pairs, film = (B x 1)
    const float* __restrict__ sensors;  // n_sensors x 16: origin[3] left[3] up[3] dir[3] tan_x tan_y near pad; NULL = off
    int   n_sensors, film_w, film_h;
    uint32_t seed_pixels, seed_offsets;
    // integrator
    int max_depth, hide_emitters, use_nee, use_drt, use_drt_subsampling, use_drt_mis;
    int nerf_queries, nerf_jitter, nerf_activation;  // `nerf` integrator (nerf.py:27-35); albedo = emission grid
    // envmap emitter (uivr_set_envmap; NULL env_data = the constant emitter `radiance`); see uivr_env.cuh
    const float4* __restrict__ env_data;   // (env_h, env_w + 1): RGB vertex radiance, w = density of patch (y, x)
    const float* __restrict__ env_marg;    // env_h - 1
    const float* __restrict__ env_cond;    // (env_h - 1, env_w)
    int   env_w, env_h;
    float env_scale;
    float env_to_world[9], world_to_env[9], local_to_world[9];
    // launch
    uint32_t seed, alt_seed, spp;
    float inv_spp;
    int shard_rank, shard_count, shard_block;
    uint32_t npix;            // W*H
    uint32_t n_slots;         // local pixel slots of this shard
    // buffers
    float* image;             // [H,W,3] accumulated sums (forward)
    float* sample_L;          // optional [S,3]
    const float* __restrict__ grad_image;
    float* dsigma;            // [Z,Y,X]
    float* dalbedo;           // [Z,Y,X,3]
    float4* dalbedo4;         // (UIVR_DALBEDO_V4 builds) RGBA-padded accumulation buffer, else NULL
    float4* dsigma4;          // (UIVR_DSIGMA_TILED builds) 2x2 (x, y) tiles of d sigma_t, four copies by the parity of the
    int   tile_x, tile_y, tile_z; // tile origin: [copy][z][tile_y][tile_x] float4; else NULL
    unsigned long long* counters;
    unsigned int* work_counter;
    int fetch_chunk;          // work items a CTA of the forward kernel reserves at a time (uivr_pool.cuh, UIVR_POOL_CHUNK_FWD)
    int walk_limit;           // watchdog: supergrid cells a walker warp may step in one go (uivr_debug_set_walk_limit)
    uint4* desc;              // adjoint launch: vertex descriptors, [CTA][slot][desc_cap][4] (uivr_pool.cuh)
    int desc_cap;             // descriptors per slot = max_depth + 1
    float2* neelog;           // adjoint launch: NEE collision log, [CTA][slot][kNeeLog] (t, sigma_n)
    uint32_t* records;        // backward: reservoir records (kRecWords words each), adjoint -> DRT launch
    unsigned int* rec_count;  // number of records appended
    unsigned int* debug;      // [64] watchdog record of the slot-pool kernel (word 0 != 0: tripped)
};

// --------------------------------------------------------------------------------------
// RNG: TEA + PCG32 `independent` sampler  
// --------------------------------------------------------------------------------------
__host__ __device__ inline void tea(uint32_t v0, uint32_t v1, uint32_t& o0, uint32_t& o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    o0 = v0;
    o1 = v1;
}

struct Rng {
    uint64_t state, inc;

    __host__ __device__ inline uint32_t next() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dull + inc;
        uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xs >> rot) | (xs << ((0u - rot) & 31u));
    }
    __host__ __device__ inline void seed_stream(uint64_t initstate, uint64_t initseq) {
        state = 0;
        inc = (initseq << 1u) | 1u;
        next();
        state += initstate;
        next();
    }
    // sampler.seed(seed, wavefront): lane idx gets PCG32(TEA(seed, idx))
    __host__ __device__ inline void seed_sampler(uint32_t seed, uint32_t idx) {
        uint32_t v0, v1;
        tea(seed, idx, v0, v1);
        seed_stream(v0, v1);
    }
    UIVR_DEV float f() { return __uint_as_float((next() >> 9) | 0x3f800000u) - 1.0f; }
};

// --------------------------------------------------------------------------------------
// exact-op transcendental replacements (same coefficients / order as the oracle)
// --------------------------------------------------------------------------------------
UIVR_DEV float neg_log1m(float u) {
    float x = 1.0f - u;
    uint32_t ix = __float_as_uint(x) + (0x3f800000u - 0x3f3504f3u);
    int e = (int) (ix >> 23) - 127;
    float f = __uint_as_float((ix & 0x007fffffu) + 0x3f3504f3u) - 1.0f;
    float q = -0x1.2d9544p-4f;
    q = fmaf(q, f, 0x1.0276dcp-3f);
    q = fmaf(q, f, -0x1.0e610cp-3f);
    q = fmaf(q, f, 0x1.235f0ep-3f);
    q = fmaf(q, f, -0x1.5467dap-3f);
    q = fmaf(q, f, 0x1.9998b0p-3f);
    q = fmaf(q, f, -0x1.00023cp-2f);
    q = fmaf(q, f, 0x1.555564p-2f);
    float f2 = f * f;
    float lm = fmaf(f2 * f, q, fmaf(-0.5f, f2, f));
    float l = fmaf((float) e, 0x1.62e430p-1f, lm);
    return -l;
}

UIVR_DEV void sincos2pi(float x, float& s, float& c) {
    float y = 4.0f * x;
    int k = (int) (y + 0.5f);
    float r = y - (float) k;
    float z = r * r;
    float ps = -0x1.2d9b78p-8f;
    ps = fmaf(ps, z, 0x1.465ec4p-4f);
    ps = fmaf(ps, z, -0x1.4abbbap-1f);
    ps = fmaf(ps, z, 0x1.921fb6p+0f);
    ps = ps * r;
    float pc = 0x1.d9c322p-11f;
    pc = fmaf(pc, z, -0x1.55c57ap-6f);
    pc = fmaf(pc, z, 0x1.03c1dcp-2f);
    pc = fmaf(pc, z, -0x1.3bd3ccp+0f);
    pc = fmaf(pc, z, 1.0f);
    float ss = (k & 1) ? pc : ps;
    float cc = (k & 1) ? ps : pc;
    s = (k & 2) ? -ss : ss;
    c = (((k + 1) & 2) ? -cc : cc);
}

// warp::square_to_uniform_sphere
UIVR_DEV void uniform_sphere(float xi1, float xi2, float& wx, float& wy, float& wz) {
    float z = fmaf(-2.0f, xi2, 1.0f);
    float r2 = fmaf(-z, z, 1.0f);
    float r = sqrtf(r2 > 0.0f ? r2 : 0.0f);
    float s, c;
    sincos2pi(xi1, s, c);
    wx = r * c;
    wy = r * s;
    wz = z;
}

UIVR_DEV float lerpf(float a, float b, float w) { return fmaf(w, b - a, a); }
UIVR_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// --------------------------------------------------------------------------------------
// event counters (only touched by COUNT template instances)
// --------------------------------------------------------------------------------------
enum { C_SIGMA = 0, C_ALBEDO, C_MAJ, C_SSCAT, C_ASCAT, C_HITS, C_REAL, C_DRAWS, C_SAMPLES, C_NUM };

template <bool COUNT>
struct Counters;
template <>
struct Counters<false> {
    UIVR_DEV void add(int, unsigned) {}
    UIVR_DEV void flush(unsigned long long*) {}
};
template <>
struct Counters<true> {
    unsigned v[C_NUM];
    UIVR_DEV Counters() {
#pragma unroll
        for (int i = 0; i < C_NUM; ++i) v[i] = 0;
    }
    UIVR_DEV void add(int k, unsigned n) { v[k] += n; }
    UIVR_DEV void flush(unsigned long long* g) {
#pragma unroll
        for (int i = 0; i < C_NUM; ++i) {
            unsigned s = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(&g[i], (unsigned long long) s);
            v[i] = 0;
        }
    }
};

// RNG wrapper that optionally counts draws
template <bool COUNT>
UIVR_DEV float draw(Rng& r, Counters<COUNT>& K) {
    K.add(C_DRAWS, 1);
    return r.f();
}

// --------------------------------------------------------------------------------------
// grid lookups
// --------------------------------------------------------------------------------------
// Taps are gathers with little reuse (one 32-byte sector per sigma_t tap out of 537 MB, 8 x 12 bytes per albedo
// tap).  UIVR_TAP_NOALLOC=1 reads them through the non-coherent path WITHOUT allocating in L1 (to keep the L1 for
// the walk table): measured SLOWER on config 3 (460 vs 484 Msamples/s, profiles/r02_history.md), so plain __ldg
// is the default.
#ifndef UIVR_TAP_NOALLOC
#define UIVR_TAP_NOALLOC 0
#endif
UIVR_DEV float4 ldg_tap4(const float4* p) {
#if UIVR_TAP_NOALLOC
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
UIVR_DEV float ldg_tap(const float* p) {
#if UIVR_TAP_NOALLOC
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
UIVR_DEV bool inside_unit(float x, float y, float z) {
    return x >= 0.0f && x <= 1.0f && y >= 0.0f && y <= 1.0f && z >= 0.0f && z <= 1.0f;
}

// sigma_t(p) = scale * trilinear(grid, p) through the corner-octet layout: one 32-byte
// sector (2 x LDG.128) per tap, border clamping baked into the layout.
UIVR_DEV float sigma_tap(const Params& P, float px, float py, float pz) {
    if (!inside_unit(px, py, pz)) return 0.0f;
    float qx = fmaf(px, P.fres[0], -0.5f), qy = fmaf(py, P.fres[1], -0.5f), qz = fmaf(pz, P.fres[2], -0.5f);
    float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
    float wx = qx - fx, wy = qy - fy, wz = qz - fz;
    int ix = (int) fx + 1, iy = (int) fy + 1, iz = (int) fz + 1;
    size_t cell = ((size_t) iz * P.ores[1] + iy) * P.ores[0] + ix;
    const float4 a = ldg_tap4(P.oct + 2 * cell);
    const float4 b = ldg_tap4(P.oct + 2 * cell + 1);
    float c00 = lerpf(a.x, a.y, wx), c10 = lerpf(a.z, a.w, wx);
    float c01 = lerpf(b.x, b.y, wx), c11 = lerpf(b.z, b.w, wx);
    float c0 = lerpf(c00, c10, wy), c1 = lerpf(c01, c11, wy);
    return P.scale * lerpf(c0, c1, wz);
}

struct GridCell {
    int x0, x1, y0, y1, z0, z1;
    float wx, wy, wz;
};

UIVR_DEV bool grid_cell(const Params& P, float px, float py, float pz, GridCell& g) {
    if (!inside_unit(px, py, pz)) return false;
    float qx = fmaf(px, P.fres[0], -0.5f), qy = fmaf(py, P.fres[1], -0.5f), qz = fmaf(pz, P.fres[2], -0.5f);
    float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
    g.wx = qx - fx; g.wy = qy - fy; g.wz = qz - fz;
    int ix = (int) fx, iy = (int) fy, iz = (int) fz;
    g.x0 = max(ix, 0); g.x1 = min(ix + 1, P.res[0] - 1);
    g.y0 = max(iy, 0); g.y1 = min(iy + 1, P.res[1] - 1);
    g.z0 = max(iz, 0); g.z1 = min(iz + 1, P.res[2] - 1);
    return true;
}

UIVR_DEV void albedo_tap(const Params& P, float px, float py, float pz, float a[3]) {
    GridCell g;
    if (!grid_cell(P, px, py, pz, g)) { a[0] = a[1] = a[2] = 0.0f; return; }
    const size_t sy = (size_t) P.res[0], sz = (size_t) P.res[0] * P.res[1];
    const float* b000 = P.albedo + 3 * (g.z0 * sz + g.y0 * sy + g.x0);
    const float* b100 = P.albedo + 3 * (g.z0 * sz + g.y0 * sy + g.x1);
    const float* b010 = P.albedo + 3 * (g.z0 * sz + g.y1 * sy + g.x0);
    const float* b110 = P.albedo + 3 * (g.z0 * sz + g.y1 * sy + g.x1);
    const float* b001 = P.albedo + 3 * (g.z1 * sz + g.y0 * sy + g.x0);
    const float* b101 = P.albedo + 3 * (g.z1 * sz + g.y0 * sy + g.x1);
    const float* b011 = P.albedo + 3 * (g.z1 * sz + g.y1 * sy + g.x0);
    const float* b111 = P.albedo + 3 * (g.z1 * sz + g.y1 * sy + g.x1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float c00 = lerpf(ldg_tap(b000 + c), ldg_tap(b100 + c), g.wx);
        float c10 = lerpf(ldg_tap(b010 + c), ldg_tap(b110 + c), g.wx);
        float c01 = lerpf(ldg_tap(b001 + c), ldg_tap(b101 + c), g.wx);
        float c11 = lerpf(ldg_tap(b011 + c), ldg_tap(b111 + c), g.wx);
        a[c] = lerpf(lerpf(c00, c10, g.wy), lerpf(c01, c11, g.wy), g.wz);
    }
}

// adjoint of the lookups: scatter-add g*w_k into the 8 voxels.
// A/B switches of the two scatter variants BASELINE.json's north_star names (profiles/r02_scatter_ab.txt):
//   UIVR_SCATTER_MATCH  1: warp-aggregate with __match_any_sync -- lanes of the batch that hit the same voxel are
//                          summed by the lowest of them, one RED per distinct voxel
//   UIVR_DALBEDO_V4     1: d albedo accumulated RGBA-padded (Params::dalbedo4), one red.global.add.v4.f32 per corner
//                          instead of three scalar REDs; k_rgba_to_rgb folds it into the caller's (Z,Y,X,3) tensor
#ifndef UIVR_SCATTER_MATCH
#define UIVR_SCATTER_MATCH 0
#endif
#ifndef UIVR_DSIGMA_V2
#define UIVR_DSIGMA_V2 1   // the two x-neighbours of a d sigma_t scatter go out as ONE red.global.add.v2.f32 when they
#endif                     // are adjacent and 8-byte aligned (about half of the scatters)
#ifndef UIVR_DSIGMA_TILED
#define UIVR_DSIGMA_TILED 1 // d sigma_t is accumulated in 2 x 2 (x, y) tiles (Params::dsigma4): the four voxels a scatter touches
#endif                      // in one z plane are ONE red.global.add.v4.f32 (2 REDs per scatter instead of 8)
#ifndef UIVR_DALBEDO_V4
#define UIVR_DALBEDO_V4 1
#endif

UIVR_DEV void red_add(float* base, size_t idx, float v) {
#if UIVR_SCATTER_MATCH
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(active, (unsigned long long) idx);
    if (__any_sync(active, peers != (1u << lane))) {
        // some lanes share a voxel: the lowest lane of every group sums its group (uniform loop over the lanes)
        float sum = 0.0f;
#pragma unroll 1
        for (int b = 0; b < 32; ++b) {
            if (!((active >> b) & 1u)) continue;
            const float x = __shfl_sync(active, v, b);
            if ((peers >> b) & 1u) sum += x;
        }
        if ((unsigned) (__ffs(peers) - 1) == lane) atomicAdd(base + idx, sum);
        return;
    }
#endif
    atomicAdd(base + idx, v);
}

// ACC: the caller guarantees the accumulation buffers (Params::dsigma4 / dalbedo4; the slot-pool backward always has
// them), so only that path is compiled in -- the scatter is inlined at three sites of the adjoint kernel
template <bool ACC = false>
UIVR_DEV void scatter_sigma(const Params& P, float px, float py, float pz, float g) {
    GridCell c;
    if (!grid_cell(P, px, py, pz, c)) return;
    const float gs = P.scale * g;
    const size_t sy = (size_t) P.res[0], sz = (size_t) P.res[0] * P.res[1];
    const float ux = 1.0f - c.wx, uy = 1.0f - c.wy, uz = 1.0f - c.wz;
#if UIVR_DSIGMA_TILED && !UIVR_SCATTER_MATCH
    if (ACC || P.dsigma4) {
        // tile (x0 >> 1, y0 >> 1) of the copy picked by the parities of (x0, y0) holds (x0, y0), (x0+1, y0), (x0, y0+1),
        // (x0+1, y0+1) contiguously; a clamped neighbour (x1 == x0 at the border) folds into the slot of x0
        const int sx = c.x1 - c.x0, sy1 = c.y1 - c.y0;
#if UIVR_DSIGMA_TILED == 2
        // 2 x 2 x 2 tiles, eight copies: both z planes of a scatter lie in ONE 32-byte sector
        const size_t tile = (((size_t) (((c.x0 & 1) | ((c.y0 & 1) << 1) | ((c.z0 & 1) << 2)) * P.tile_z + (c.z0 >> 1)) * P.tile_y + (c.y0 >> 1)) * P.tile_x + (c.x0 >> 1)) * 2;
        const int sz1 = c.z1 - c.z0;
#else
        const size_t tile = ((size_t) (((c.x0 & 1) | ((c.y0 & 1) << 1)) * P.res[2]) * P.tile_y + (c.y0 >> 1)) * P.tile_x + (c.x0 >> 1);
        const size_t plane = (size_t) P.tile_y * P.tile_x;
#endif
#pragma unroll
        for (int kz = 0; kz < 2; ++kz) {
            const float wz = kz ? c.wz : uz;
            const float v00 = gs * ((ux * uy) * wz), v10 = gs * ((c.wx * uy) * wz);
            const float v01 = gs * ((ux * c.wy) * wz), v11 = gs * ((c.wx * c.wy) * wz);
            float t0 = v00, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;   // slots (0,0) (1,0) (0,1) (1,1)
            if (sx) t1 = v10; else t0 += v10;
            if (sy1) { t2 = v01; if (sx) t3 = v11; else t2 += v11; }
            else { t0 += v01; if (sx) t1 += v11; else t0 += v11; }
#if UIVR_DSIGMA_TILED == 2
            float4* dst = P.dsigma4 + tile + (kz ? sz1 : 0);
#else
            float4* dst = P.dsigma4 + tile + (size_t) (kz ? c.z1 : c.z0) * plane;
#endif
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst),
                         "f"(t0), "f"(t1), "f"(t2), "f"(t3) : "memory");
        }
        return;
    }
    if (ACC) return;
#endif
#if UIVR_DSIGMA_V2 && !UIVR_SCATTER_MATCH
    const bool adjacent = c.x1 == c.x0 + 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t row = ((k & 2) ? c.z1 : c.z0) * sz + ((k & 1) ? c.y1 : c.y0) * sy;
        const float wyz = ((k & 1) ? c.wy : uy), wz = ((k & 2) ? c.wz : uz);
        const float a = gs * ((ux * wyz) * wz), b = gs * ((c.wx * wyz) * wz);
        const size_t i0 = row + c.x0;
        if (adjacent && ((uintptr_t) (P.dsigma + i0) & 7u) == 0u) {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(P.dsigma + i0), "f"(a), "f"(b) : "memory");
        } else {
            atomicAdd(P.dsigma + i0, a);
            atomicAdd(P.dsigma + row + c.x1, b);
        }
    }
#else
    red_add(P.dsigma, c.z0 * sz + c.y0 * sy + c.x0, gs * ((ux * uy) * uz));
    red_add(P.dsigma, c.z0 * sz + c.y0 * sy + c.x1, gs * ((c.wx * uy) * uz));
    red_add(P.dsigma, c.z0 * sz + c.y1 * sy + c.x0, gs * ((ux * c.wy) * uz));
    red_add(P.dsigma, c.z0 * sz + c.y1 * sy + c.x1, gs * ((c.wx * c.wy) * uz));
    red_add(P.dsigma, c.z1 * sz + c.y0 * sy + c.x0, gs * ((ux * uy) * c.wz));
    red_add(P.dsigma, c.z1 * sz + c.y0 * sy + c.x1, gs * ((c.wx * uy) * c.wz));
    red_add(P.dsigma, c.z1 * sz + c.y1 * sy + c.x0, gs * ((ux * c.wy) * c.wz));
    red_add(P.dsigma, c.z1 * sz + c.y1 * sy + c.x1, gs * ((c.wx * c.wy) * c.wz));
#endif
}

template <bool ACC = false>
UIVR_DEV void scatter_albedo(const Params& P, float px, float py, float pz, const float g[3]) {
    GridCell c;
    if (!grid_cell(P, px, py, pz, c)) return;
    const size_t sy = (size_t) P.res[0], sz = (size_t) P.res[0] * P.res[1];
    const float ux = 1.0f - c.wx, uy = 1.0f - c.wy, uz = 1.0f - c.wz;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int x = (k & 1) ? c.x1 : c.x0, y = (k & 2) ? c.y1 : c.y0, z = (k & 4) ? c.z1 : c.z0;
        const float w = (((k & 1) ? c.wx : ux) * ((k & 2) ? c.wy : uy)) * ((k & 4) ? c.wz : uz);
#if UIVR_DALBEDO_V4
        if (ACC || P.dalbedo4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(P.dalbedo4 + (z * sz + y * sy + x)),
                         "f"(g[0] * w), "f"(g[1] * w), "f"(g[2] * w), "f"(0.0f) : "memory");
            continue;
        }
        if (ACC) continue;
#endif
        float* dst = P.dalbedo + 3 * (z * sz + y * sy + x);
        atomicAdd(dst + 0, g[0] * w);
        atomicAdd(dst + 1, g[1] * w);
        atomicAdd(dst + 2, g[2] * w);
    }
}

// --------------------------------------------------------------------------------------
// segments (local-space origin / direction, t in world units)
// --------------------------------------------------------------------------------------
struct Seg {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, tmax;
};

UIVR_DEV void dir_to_local(const Params& P, float wx, float wy, float wz, float& dx, float& dy, float& dz) {
    const float* M = P.to_local;
    dx = fmaf(M[0], wx, fmaf(M[1], wy, M[2] * wz));
    dy = fmaf(M[4], wx, fmaf(M[5], wy, M[6] * wz));
    dz = fmaf(M[8], wx, fmaf(M[9], wy, M[10] * wz));
}

UIVR_DEV float exit_axis(float o, float d, float& inv) {
    if (d != 0.0f) {
        inv = 1.0f / d;
        return ((d > 0.0f ? 1.0f : 0.0f) - o) * inv;
    }
    inv = UIVR_INF;
    return UIVR_INF;
}

// distance to the box boundary from a point inside local [0,1]^3; fills inv_d
UIVR_DEV float exit_distance(Seg& s) {
    float t = exit_axis(s.ox, s.dx, s.ix);
    t = fminf(t, exit_axis(s.oy, s.dy, s.iy));
    t = fminf(t, exit_axis(s.oz, s.dz, s.iz));
    return t;
}

// new segment leaving local point p along world direction w; false on accidental escape
UIVR_DEV bool make_segment(const Params& P, float px, float py, float pz, float wx, float wy, float wz, Seg& s) {
    s.ox = px; s.oy = py; s.oz = pz;
    dir_to_local(P, wx, wy, wz, s.dx, s.dy, s.dz);
    s.tmax = exit_distance(s);
    return s.tmax > 0.0f && s.tmax < UIVR_INF;
}

// perspective sensor + reach_medium.  0 = missed (escaped), 1 = entered, 2 = dead.
// F = sensor frame: origin[3] left[3] up[3] dir[3] tan_x tan_y near_clip; (u, v) = film position in [0,1]^2
UIVR_DEV int camera_segment_frame(const Params& P, const float F[15], float u, float v, Seg& s);

// the scene's own sensor: frame F and film position (u, v) of pixel `pix` with sub-pixel offset (jx, jy)
UIVR_DEV void sensor_film_position(const Params& P, uint32_t pix, float jx, float jy, float F[15], float& u, float& v) {
    uint32_t px = pix % (uint32_t) P.width, py = pix / (uint32_t) P.width;
    u = ((float) px + jx) * P.inv_w;
    v = ((float) py + jy) * P.inv_h;
    F[0] = P.cam_origin[0]; F[1] = P.cam_origin[1]; F[2] = P.cam_origin[2];
    F[3] = P.cam_left[0]; F[4] = P.cam_left[1]; F[5] = P.cam_left[2];
    F[6] = P.cam_up[0]; F[7] = P.cam_up[1]; F[8] = P.cam_up[2];
    F[9] = P.cam_dir[0]; F[10] = P.cam_dir[1]; F[11] = P.cam_dir[2];
    F[12] = P.tan_x; F[13] = P.tan_y; F[14] = P.near_clip;
}

UIVR_DEV int camera_segment(const Params& P, uint32_t pix, float jx, float jy, Seg& s) {
    float F[15], u, v;
    sensor_film_position(P, pix, jx, jy, F, u, v);
    return camera_segment_frame(P, F, u, v, s);
}

UIVR_DEV int camera_segment_frame(const Params& P, const float F[15], float u, float v, Seg& s) {
    float cx = F[12] * fmaf(-2.0f, u, 1.0f);
    float cy = F[13] * fmaf(-2.0f, v, 1.0f);
    float d0 = fmaf(cx, F[3], fmaf(cy, F[6], F[9]));
    float d1 = fmaf(cx, F[4], fmaf(cy, F[7], F[10]));
    float d2 = fmaf(cx, F[5], fmaf(cy, F[8], F[11]));
    float len = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
    float inv_len = 1.0f / len;
    float near_t = F[14] * len;
    d0 *= inv_len; d1 *= inv_len; d2 *= inv_len;
    float o0 = fmaf(near_t, d0, F[0]);
    float o1 = fmaf(near_t, d1, F[1]);
    float o2 = fmaf(near_t, d2, F[2]);
    const float* M = P.to_local;
    float ol[3], dl[3];
    ol[0] = fmaf(M[0], o0, fmaf(M[1], o1, fmaf(M[2], o2, M[3])));
    ol[1] = fmaf(M[4], o0, fmaf(M[5], o1, fmaf(M[6], o2, M[7])));
    ol[2] = fmaf(M[8], o0, fmaf(M[9], o1, fmaf(M[10], o2, M[11])));
    dir_to_local(P, d0, d1, d2, dl[0], dl[1], dl[2]);
    s.dx = dl[0]; s.dy = dl[1]; s.dz = dl[2];  // also for rays that miss the box: the envmap lookup needs it
    float tn = -UIVR_INF, tf = UIVR_INF;
#pragma unroll
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
    if (!(tn > 0.0f)) return 2;
    float e0 = fmaf(tn, dl[0], ol[0]), e1 = fmaf(tn, dl[1], ol[1]), e2 = fmaf(tn, dl[2], ol[2]);
    const float lo = UIVR_ENTRY_EPS, hi = 1.0f - UIVR_ENTRY_EPS;
    s.ox = e0 < lo ? lo : (e0 > hi ? hi : e0);
    s.oy = e1 < lo ? lo : (e1 > hi ? hi : e1);
    s.oz = e2 < lo ? lo : (e2 > hi ? hi : e2);
    s.tmax = exit_distance(s);
    return (s.tmax > 0.0f && s.tmax < UIVR_INF) ? 1 : 2;
}

// sample_batch_pixels + sample_batch_rays (batched.py:397-467): ray of wavefront entry idx = b * spp + j
template <typename SeedFn>
UIVR_DEV void batch_film_position_t(const Params& P, uint32_t b, uint32_t idx, float F[15], float& u, float& v, SeedFn seed) {
    Rng q;
    seed(q, P.seed_pixels, b);
    const float u0 = q.f(), u1 = q.f(), u2 = q.f();
    uint32_t si = (uint32_t) ((float) P.n_sensors * u0);
    uint32_t px = (uint32_t) ((float) P.film_w * u1), py = (uint32_t) ((float) P.film_h * u2);
    si = min(si, (uint32_t) P.n_sensors - 1u);
    px = min(px, (uint32_t) P.film_w - 1u);
    py = min(py, (uint32_t) P.film_h - 1u);
    Rng o;
    seed(o, P.seed_offsets, idx);
    const float jx = o.f(), jy = o.f();
    u = ((float) px + jx) * (1.0f / (float) P.film_w);
    v = ((float) py + jy) * (1.0f / (float) P.film_h);
    const float4* f4 = reinterpret_cast<const float4*>(P.sensors + 16 * (size_t) si);
    const float4 a = __ldg(f4), bb = __ldg(f4 + 1), c = __ldg(f4 + 2), d = __ldg(f4 + 3);
    F[0] = a.x; F[1] = a.y; F[2] = a.z; F[3] = a.w; F[4] = bb.x; F[5] = bb.y; F[6] = bb.z; F[7] = bb.w;
    F[8] = c.x; F[9] = c.y; F[10] = c.z; F[11] = c.w; F[12] = d.x; F[13] = d.y; F[14] = d.z;
}

UIVR_DEV int batch_segment(const Params& P, uint32_t b, uint32_t idx, Seg& s) {
    float F[15], u, v;
    batch_film_position_t(P, b, idx, F, u, v, [](Rng& r, uint32_t sd, uint32_t i) { r.seed_sampler(sd, i); });
    return camera_segment_frame(P, F, u, v, s);
}

// --------------------------------------------------------------------------------------
// free-flight walk over the majorant supergrid (Medium::sample_interaction, App. B.5)
// --------------------------------------------------------------------------------------
constexpr uint32_t kWalkEmpty = 0x80000000u;
constexpr uint32_t kWalkBorder = 0x800001FFu;  // bit 8 tells the border from an empty in-grid cell (event counting)

struct Walk {
    float t, tmax;
    float tnx, tny, tnz;
    int cx, cy, cz;
    float sb;
    bool exit;      // the current cell is empty and so is everything ahead of it in the ray's octant
    unsigned obit;  // 1 << octant
};

// majorant of cell (cx, cy, cz) + "nothing but empty space ahead" flag, from the walk table
template <bool COUNT>
UIVR_DEV float majorant_at(const Params& P, int cx, int cy, int cz, unsigned obit, bool& exit, Counters<COUNT>& K) {
    K.add(C_MAJ, 1);
    const uint32_t e = __ldg(P.wtab + ((size_t) (cz + 1) * P.pm[1] + (cy + 1)) * P.pm[0] + (cx + 1));
    exit = (int) e < 0 && (e & obit) != 0u;
    return (int) e > 0 ? __uint_as_float(e) : 0.0f;
}

UIVR_DEV void walk_axis_init(float o, float d, float inv, float fm, float cs, int m, int& c, float& tn) {
    c = clampi((int) floorf(o * fm), 0, m - 1);
    if (d > 0.0f) tn = ((float) (c + 1) * cs - o) * inv;
    else if (d < 0.0f) tn = ((float) c * cs - o) * inv;
    else tn = UIVR_INF;
}

template <bool COUNT>
UIVR_DEV void walk_init(const Params& P, const Seg& s, Walk& w, Counters<COUNT>& K) {
    w.t = 0.0f;
    w.tmax = s.tmax;
    walk_axis_init(s.ox, s.dx, s.ix, P.fmres[0], P.mcs[0], P.mres[0], w.cx, w.tnx);
    walk_axis_init(s.oy, s.dy, s.iy, P.fmres[1], P.mcs[1], P.mres[1], w.cy, w.tny);
    walk_axis_init(s.oz, s.dz, s.iz, P.fmres[2], P.mcs[2], P.mres[2], w.cz, w.tnz);
    w.obit = 1u << ((s.dx < 0.0f ? 1 : 0) | (s.dy < 0.0f ? 2 : 0) | (s.dz < 0.0f ? 4 : 0));
    w.sb = majorant_at<COUNT>(P, w.cx, w.cy, w.cz, w.obit, w.exit, K);
}

// advance to the next tentative collision; false when the segment end is reached
template <bool COUNT>
UIVR_DEV bool walk_next(const Params& P, const Seg& s, Walk& w, float u, Counters<COUNT>& K) {
    float tau = neg_log1m(u);
    for (;;) {
        if (w.exit) return false;
        int ax = 0;
        float tn = w.tnx;
        if (w.tny < tn) { ax = 1; tn = w.tny; }
        if (w.tnz < tn) { ax = 2; tn = w.tnz; }
        const float t_end = tn < w.tmax ? tn : w.tmax;
        float len = t_end - w.t;
        if (len < 0.0f) len = 0.0f;
        if (w.sb > 0.0f) {
            const float dtau = w.sb * len;
            if (tau < dtau) {
                float t = w.t + tau / w.sb;
                if (t > t_end) t = t_end;
                w.t = t;
                return true;
            }
            tau -= dtau;
        }
        if (t_end > w.t) w.t = t_end;
        if (!(tn < w.tmax)) return false;
        if (ax == 0) {
            w.cx += s.dx > 0.0f ? 1 : -1;
            if (w.cx < 0 || w.cx >= P.mres[0]) return false;
            w.tnx += fabsf(P.mcs[0] * s.ix);
        } else if (ax == 1) {
            w.cy += s.dy > 0.0f ? 1 : -1;
            if (w.cy < 0 || w.cy >= P.mres[1]) return false;
            w.tny += fabsf(P.mcs[1] * s.iy);
        } else {
            w.cz += s.dz > 0.0f ? 1 : -1;
            if (w.cz < 0 || w.cz >= P.mres[2]) return false;
            w.tnz += fabsf(P.mcs[2] * s.iz);
        }
        w.sb = majorant_at<COUNT>(P, w.cx, w.cy, w.cz, w.obit, w.exit, K);
    }
}

}  // namespace uivr
