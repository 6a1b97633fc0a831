// uivr_kernels.cuh -- __global__ entry points (sm_100a).
#pragma once

#include "uivr_path.cuh"

namespace uivr {

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------------
// K4a: corner-octet tap layout.  Cell (ix,iy,iz) in [0,res]^3 holds the 8 voxels a
// trilinear lookup with floor(q)+1 == (ix,iy,iz) needs, border clamping applied, as two
// float4: {v(x0,y0,z0), v(x1,y0,z0), v(x0,y1,z0), v(x1,y1,z0)}, {.. z1 ..}.  One tap = one
// aligned 32-byte sector.  Costs 8x the memory of sigma_t (HBM is 180 GB; 512^3 -> 4.3 GB).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_build_octets(const float* __restrict__ sigma_t, float4* __restrict__ oct,
                                                         int rx, int ry, int rz) {
    const int ox = rx + 1, oy = ry + 1, oz = rz + 1;
    const size_t n = (size_t) ox * oy * oz;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int ix = (int) (i % ox), iy = (int) ((i / ox) % oy), iz = (int) (i / ((size_t) ox * oy));
        const int x0 = max(ix - 1, 0), x1 = min(ix, rx - 1);
        const int y0 = max(iy - 1, 0), y1 = min(iy, ry - 1);
        const int z0 = max(iz - 1, 0), z1 = min(iz, rz - 1);
        const size_t sy = (size_t) rx, sz = (size_t) rx * ry;
        float4 a, b;
        a.x = __ldg(sigma_t + z0 * sz + y0 * sy + x0);
        a.y = __ldg(sigma_t + z0 * sz + y0 * sy + x1);
        a.z = __ldg(sigma_t + z0 * sz + y1 * sy + x0);
        a.w = __ldg(sigma_t + z0 * sz + y1 * sy + x1);
        b.x = __ldg(sigma_t + z1 * sz + y0 * sy + x0);
        b.y = __ldg(sigma_t + z1 * sz + y0 * sy + x1);
        b.z = __ldg(sigma_t + z1 * sz + y1 * sy + x0);
        b.w = __ldg(sigma_t + z1 * sz + y1 * sy + x1);
        oct[2 * i] = a;
        oct[2 * i + 1] = b;
    }
}

// ---------------------------------------------------------------------------------------
// K4b: majorant supergrid: cell value = scale * max over the voxels whose
// trilinear footprint touches the cell.  One warp per cell.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int floordiv_pos(int a, int b) {
    int q = a / b;
    if ((a % b) != 0 && a < 0) --q;
    return q;
}

__global__ void __launch_bounds__(kBlock) k_build_majorant(const float* __restrict__ sigma_t, float* __restrict__ maj,
                                                           int rx, int ry, int rz, int mx, int my, int mz, float scale) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int ncell = mx * my * mz;
    for (int cell = warp; cell < ncell; cell += nwarps) {
        const int cx = cell % mx, cy = (cell / mx) % my, cz = cell / (mx * my);
        const int lx = clampi(floordiv_pos(2 * cx * rx - mx, 2 * mx), 0, rx - 1);
        const int hx = clampi(floordiv_pos(2 * (cx + 1) * rx - mx, 2 * mx) + 1, 0, rx - 1);
        const int ly = clampi(floordiv_pos(2 * cy * ry - my, 2 * my), 0, ry - 1);
        const int hy = clampi(floordiv_pos(2 * (cy + 1) * ry - my, 2 * my) + 1, 0, ry - 1);
        const int lz = clampi(floordiv_pos(2 * cz * rz - mz, 2 * mz), 0, rz - 1);
        const int hz = clampi(floordiv_pos(2 * (cz + 1) * rz - mz, 2 * mz) + 1, 0, rz - 1);
        const int nx = hx - lx + 1, ny = hy - ly + 1, nz = hz - lz + 1;
        const int total = nx * ny * nz;
        float m = 0.0f;
        for (int i = lane; i < total; i += 32) {
            const int x = lx + i % nx, y = ly + (i / nx) % ny, z = lz + i / (nx * ny);
            m = fmaxf(m, __ldg(sigma_t + ((size_t) z * ry + y) * rx + x));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) maj[cell] = scale * m;
    }
}

// ---------------------------------------------------------------------------------------
// K4c: walk table = supergrid + exit mask (uivr_device.cuh, Params::wtab).  Bit o of an empty cell's exit
// mask says that every cell of the box between the cell and the grid corner octant o points to is empty:
// an AND over that box, computed as three separable prefix / suffix sweeps (x, then y, then z), one thread
// per grid line.  After the sweep along axis a, bit index gains bit a (0: suffix = direction positive,
// 1: prefix = direction negative), so the final byte is indexed by the octant.
// ---------------------------------------------------------------------------------------
// sweep along the axis with stride `sa` and length `na`; `nb` bits in, 2 * nb bits out
__global__ void __launch_bounds__(kBlock) k_exit_sweep(const float* __restrict__ maj, const uint8_t* __restrict__ in,
                                                       uint8_t* __restrict__ out, int na, size_t sa, int n1, size_t s1,
                                                       int n2, size_t s2, int nb) {
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n1 * n2) return;
    const size_t base = (size_t) (line % n1) * s1 + (size_t) (line / n1) * s2;
    const unsigned full = (1u << nb) - 1u;
    unsigned run = full;
    for (int i = na - 1; i >= 0; --i) {  // suffix: cells i' >= i
        const size_t c = base + (size_t) i * sa;
        run &= in ? (unsigned) in[c] : (maj[c] > 0.0f ? 0u : 1u);
        out[c] = (uint8_t) run;
    }
    run = full;
    for (int i = 0; i < na; ++i) {       // prefix: cells i' <= i
        const size_t c = base + (size_t) i * sa;
        run &= in ? (unsigned) in[c] : (maj[c] > 0.0f ? 0u : 1u);
        out[c] = (uint8_t) (out[c] | (run << nb));
    }
}

// `wtab` points at cell 0 of the padded grid; `slack` words before and after it belong to the allocation too and
// are filled with border words: the walkers request the word of the cell AFTER the next one speculatively, which
// for a walk standing next to the border lies one stride (<= (mx+2)(my+2) words) outside the padded grid.
__global__ void __launch_bounds__(kBlock) k_build_walk_table(const float* __restrict__ maj, const uint8_t* __restrict__ mask,
                                                             uint32_t* __restrict__ wtab, int mx, int my, int mz, int slack) {
    const int px = mx + 2, py = my + 2, pz = mz + 2;
    const int n = px * py * pz;
    for (int i = -slack + (int) (blockIdx.x * blockDim.x + threadIdx.x); i < n + slack; i += gridDim.x * blockDim.x) {
        uint32_t w = kWalkBorder;
        if (i >= 0 && i < n) {
            const int x = i % px - 1, y = (i / px) % py - 1, z = i / (px * py) - 1;
            if (x >= 0 && x < mx && y >= 0 && y < my && z >= 0 && z < mz) {
                const size_t c = ((size_t) z * my + y) * mx + x;
                const float m = maj[c];
                w = m > 0.0f ? __float_as_uint(m) : (kWalkEmpty | (uint32_t) mask[c]);
            }
        }
        wtab[i] = w;
    }
}

// (UIVR_DSIGMA_TILED builds) fold the four tile copies into the caller's (Z,Y,X) gradient (overwritten).  One thread
// per 2 x 2 block of voxels (x = 2X, 2X+1; y = 2Y, 2Y+1): the block IS tile (X, Y) of copy (0, 0), it straddles two
// tiles of the copies whose origin is odd in one axis and four of the copy that is odd in both -- nine aligned
// 16-byte loads, neighbours in a warp read neighbouring tiles.  Copy p = px + 2 py holds the tiles with origin
// (2 Xt + px, 2 Yt + py); slot = (x - origin) + 2 (y - origin).  Sum order per voxel: copy 0, 1, 2, 3.
__global__ void __launch_bounds__(kBlock) k_tiles_to_dsigma(const float4* __restrict__ tiles, float* __restrict__ out, int rx,
                                                            int ry, int rz, int tx, int ty) {
    const size_t plane = (size_t) tx * ty, nblk = plane * rz;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nblk; i += (size_t) gridDim.x * blockDim.x) {
        const int X = (int) (i % tx), Y = (int) ((i / tx) % ty), z = (int) (i / plane);
        auto T = [&](int p, int Xt, int Yt) { return __ldg(tiles + ((size_t) p * rz + z) * plane + (size_t) Yt * tx + Xt); };
        const float4 c0 = T(0, X, Y);
        float a00 = c0.x, a10 = c0.y, a01 = c0.z, a11 = c0.w;   // a<x offset><y offset>
        {   // copy 1: origin odd in x
            const float4 r = T(1, X, Y);
            a10 += r.x; a11 += r.z;
            if (X > 0) { const float4 l = T(1, X - 1, Y); a00 += l.y; a01 += l.w; }
        }
        {   // copy 2: origin odd in y
            const float4 u = T(2, X, Y);
            a01 += u.x; a11 += u.y;
            if (Y > 0) { const float4 d = T(2, X, Y - 1); a00 += d.z; a10 += d.w; }
        }
        {   // copy 3: origin odd in both
            a11 += T(3, X, Y).x;
            if (X > 0) a01 += T(3, X - 1, Y).y;
            if (Y > 0) a10 += T(3, X, Y - 1).z;
            if (X > 0 && Y > 0) a00 += T(3, X - 1, Y - 1).w;
        }
        const int x0 = 2 * X, y0 = 2 * Y;
        float* o = out + ((size_t) z * ry + y0) * rx + x0;
        o[0] = a00;
        if (x0 + 1 < rx) o[1] = a10;
        if (y0 + 1 < ry) {
            o[rx] = a01;
            if (x0 + 1 < rx) o[rx + 1] = a11;
        }
    }
}

// same for 2 x 2 x 2 tiles (UIVR_DSIGMA_TILED == 2): eight copies, two float4 (z slot 0 / 1) per tile
__global__ void __launch_bounds__(kBlock) k_tiles3_to_dsigma(const float4* __restrict__ tiles, float* __restrict__ out, int rx,
                                                             int ry, int rz, int tx, int ty, int tz) {
    const size_t n = (size_t) rx * ry * rz;
    const float* t = reinterpret_cast<const float*>(tiles);
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int x = (int) (i % rx), y = (int) ((i / rx) % ry), z = (int) (i / ((size_t) rx * ry));
        float acc = 0.0f;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int px = p & 1, py = (p >> 1) & 1, pz = p >> 2;
            if (x < px || y < py || z < pz) continue;
            const int X = (x - px) >> 1, Y = (y - py) >> 1, Z = (z - pz) >> 1;
            const int slot = ((x - px) & 1) + 2 * ((y - py) & 1) + 4 * ((z - pz) & 1);
            acc += t[((((size_t) p * tz + Z) * ty + Y) * tx + X) * 8 + slot];
        }
        out[i] = acc;
    }
}

// (UIVR_DALBEDO_V4 builds) fold the RGBA-padded accumulation buffer into the caller's (Z,Y,X,3) gradient (overwritten)
__global__ void __launch_bounds__(kBlock) k_rgba_to_rgb(const float4* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float4 v = in[i];
        out[3 * i + 0] = v.x;
        out[3 * i + 1] = v.y;
        out[3 * i + 2] = v.z;
    }
}

__global__ void __launch_bounds__(kBlock) k_scale(float* __restrict__ x, size_t n, float s) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        x[i] = x[i] * s;
}

// ---------------------------------------------------------------------------------------
// K6: optimiser step.  mi.ad.Adam.step() (optimize.py:352) fused with
// enforce_valid_params (optimize.py:169-179): one streaming pass, 16 B read + 12 B written per
// element (param, grad, m, v -> param, m, v); HBM-bound.
// ---------------------------------------------------------------------------------------
UIVR_DEV void adam_element(float& p, float g, float& m, float& v, float step, float beta1, float omb1, float beta2,
                           float omb2, float eps, float lo, float hi) {
    m = fmaf(beta1, m, omb1 * g);
    v = fmaf(beta2, v, (omb2 * g) * g);
    float q = p - (step * m) / (sqrtf(v) + eps);
    q = q < lo ? lo : q;
    p = q > hi ? hi : q;
}

__global__ void __launch_bounds__(kBlock) k_adam_step(float* __restrict__ param, const float* __restrict__ grad,
                                                      float* __restrict__ m, float* __restrict__ v, size_t n, float step,
                                                      float beta1, float beta2, float eps, float lo, float hi) {
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const size_t n4 = n / 4, tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t) gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (size_t i = tid; i < n4; i += nth) {
        float4 p = p4[i], mm = m4[i], vv = v4[i];
        const float4 g = __ldg(g4 + i);
        adam_element(p.x, g.x, mm.x, vv.x, step, beta1, omb1, beta2, omb2, eps, lo, hi);
        adam_element(p.y, g.y, mm.y, vv.y, step, beta1, omb1, beta2, omb2, eps, lo, hi);
        adam_element(p.z, g.z, mm.z, vv.z, step, beta1, omb1, beta2, omb2, eps, lo, hi);
        adam_element(p.w, g.w, mm.w, vv.w, step, beta1, omb1, beta2, omb2, eps, lo, hi);
        p4[i] = p; m4[i] = mm; v4[i] = vv;
    }
    for (size_t i = 4 * n4 + tid; i < n; i += nth)
        adam_element(param[i], grad[i], m[i], v[i], step, beta1, omb1, beta2, omb2, eps, lo, hi);
}

// ---------------------------------------------------------------------------------------
// K7: multires upsampling x2 (upsample_grid, optimize.py:203-225: scipy zoom(order=1, mode='nearest',
// grid_mode=True)).  Output voxel o samples the input at o/2 - 1/4 per axis, i.e. weights
// (1/4, 3/4) on the two nearest input voxels, indices clamped at the border.  (Z,Y,X,C) layout.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_upsample2x(const float* __restrict__ in, float* __restrict__ out, int rx, int ry,
                                                       int rz, int ch) {
    const size_t ox = 2 * (size_t) rx, oy = 2 * (size_t) ry, oz = 2 * (size_t) rz;
    const size_t n = ox * oy * oz * ch;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const int c = (int) (i % ch);
        size_t r = i / ch;
        const int x = (int) (r % ox); r /= ox;
        const int y = (int) (r % oy);
        const int z = (int) (r / oy);
        // even output index 2k: inputs (k-1, k) with weights (1/4, 3/4); odd 2k+1: (k, k+1) with (3/4, 1/4)
        const int x0 = max((x - 1) >> 1, 0), x1 = min((x + 1) >> 1, rx - 1);
        const int y0 = max((y - 1) >> 1, 0), y1 = min((y + 1) >> 1, ry - 1);
        const int z0 = max((z - 1) >> 1, 0), z1 = min((z + 1) >> 1, rz - 1);
        const float wx = (x & 1) ? 0.25f : 0.75f, wy = (y & 1) ? 0.25f : 0.75f, wz = (z & 1) ? 0.25f : 0.75f;
#define UIVR_V(zz, yy, xx) __ldg(in + (((size_t) (zz) * ry + (yy)) * rx + (xx)) * ch + c)
        const float c00 = lerpf(UIVR_V(z0, y0, x0), UIVR_V(z0, y0, x1), wx), c10 = lerpf(UIVR_V(z0, y1, x0), UIVR_V(z0, y1, x1), wx);
        const float c01 = lerpf(UIVR_V(z1, y0, x0), UIVR_V(z1, y0, x1), wx), c11 = lerpf(UIVR_V(z1, y1, x0), UIVR_V(z1, y1, x1), wx);
#undef UIVR_V
        out[i] = lerpf(lerpf(c00, c10, wy), lerpf(c01, c11, wy), wz);
    }
}

// ---------------------------------------------------------------------------------------
// variant 1: one sample per lane, warps pull 32-sample chunks from a global counter
// ---------------------------------------------------------------------------------------
UIVR_DEV bool next_chunk(const Params& P, uint64_t total, uint32_t& item) {
    unsigned base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(P.work_counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((uint64_t) base >= total) return false;
    item = base + (threadIdx.x & 31);
    return true;
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlock) k_forward_v1(const Params P) {
    Counters<COUNT> K;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    uint32_t item;
    while (next_chunk(P, total, item)) {
        uint32_t pix = 0;
        const bool live = (uint64_t) item < total && slot_to_pixel(P, item / P.spp, pix);
        float L[3] = {0.0f, 0.0f, 0.0f};
        if (live) {
            const uint32_t idx = pix * P.spp + item % P.spp;
            sample_from_camera<false, COUNT>(P, idx, nullptr, L, K);
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
__global__ void __launch_bounds__(kBlock) k_backward_v1(const Params P) {
    Counters<COUNT> K;
    const uint64_t total = (uint64_t) P.n_slots * P.spp;
    uint32_t item;
    while (next_chunk(P, total, item)) {
        uint32_t pix = 0;
        const bool live = (uint64_t) item < total && slot_to_pixel(P, item / P.spp, pix);
        if (live) {
            const uint32_t idx = pix * P.spp + item % P.spp;
            float L[3] = {0.0f, 0.0f, 0.0f};
            // batched.py:255-264: detached primal pass at seed_grad -> state_in
            sample_from_camera<false, COUNT>(P, idx, nullptr, L, K);
            K.add(C_SAMPLES, 1);
            if (P.sample_L) {
                P.sample_L[3 * (size_t) idx + 0] = L[0];
                P.sample_L[3 * (size_t) idx + 1] = L[1];
                P.sample_L[3 * (size_t) idx + 2] = L[2];
            }
            // batched.py:272-306: box film => dL = grad_image[pixel] / spp
            float dL[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) dL[c] = __ldg(P.grad_image + 3 * (size_t) pix + c) * P.inv_spp;
            // batched.py:309-318: sample(Backward, state_in = L)
            sample_from_camera<true, COUNT>(P, idx, dL, L, K);
        }
        __syncwarp();
    }
    K.flush(P.counters);
}

// ---------------------------------------------------------------------------------------
// primitive test kernels (bit-exactness checks against the oracle)
// ---------------------------------------------------------------------------------------
__global__ void k_test_neg_log1m(const float* u, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = neg_log1m(u[i]);
}
__global__ void k_test_sincos2pi(const float* x, int n, float* s, float* c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sincos2pi(x[i], s[i], c[i]);
}
__global__ void k_test_sampler(uint32_t seed, uint32_t idx0, int nstreams, int ndraws, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nstreams) {
        Rng r;
        r.seed_sampler(seed, idx0 + (uint32_t) i);
        for (int k = 0; k < ndraws; ++k) out[(size_t) i * ndraws + k] = r.f();
    }
}
__global__ void k_test_sigma_lookup(const Params P, const float* p, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sigma_tap(P, p[3 * i], p[3 * i + 1], p[3 * i + 2]);
}

}  // namespace uivr
