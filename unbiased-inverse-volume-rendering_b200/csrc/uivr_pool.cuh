// uivr_pool.cuh -- persistent SLOT-POOL megakernels (sm_100a): the default pipeline.
//
// ONE CTA per SM owns a pool of NSLOT in-flight samples whose path state lives in SHARED MEMORY (SoA,
// field-major), and every state of the per-sample state machine has a CTA-wide queue of slot ids:
//
//   Q_FREE -> [fetch] -> Q_WALK -> (walk) -> Q_TAP -> [tap] -> Q_WALK ...            (null collision)
//                                                           -> Q_VERTEX(_ADJ) -> [vertex] -> Q_SPAWN -> [spawn] -> Q_WALK
//                                         -> Q_VERTEX / Q_NEE_END / Q_SPAWN (segment end) ... -> Q_PATH_END -> Q_FREE
//
// Warps are specialised.  WALKER warps run ONLY the supergrid DDA of the free-flight walk (Medium::
// sample_interaction and its ratio-tracking / DRT siblings): ~48 instructions per cell, no RNG, no taps, no
// divisions.  The walk is a CONTINUATION stored in the pool (next-boundary times, their increments, cell
// index, remaining optical depth tau, position t): a walker lane picks one up with a dozen shared-memory
// loads, steps cells until the walk ends or tau is used up inside a cell, and hands the slot on -- on a
// tentative collision it first saves the eight words that changed.  Everything else runs in HANDLER warps,
// which pop FULL batches of 32 slots of one queue, so that the expensive code (ray generation, the walk
// set-up with its divisions, the sigma_t tap + accept/reject decision, vertices, emitter / phase sampling,
// gradient scatter) always runs with all lanes: compaction of live rays across the whole CTA.
//
// Round 1 kept the tap and the walk set-up inside the walker warps, where they ran with 11-14 of 32 lanes
// (profiles/r01_split_*_regions.txt: 65 % of the walkers' issue slots went to idle lanes); see
// profiles/r02_history.md for the measurements that led here.
//
// KIND selects what is compiled in: the forward / primal kernel, or one of the two halves of the backward
// pipeline (adjoint replay; DRT pass), see KIND_* below.
//
// Per-sample arithmetic, RNG draw order and handler logic are exactly those of uivr_path.cuh
// (= volpathsimple.py), so results stay bit-identical per sample (tests/test_gpu_parity.py).
#pragma once

#include "uivr_kernels.cuh"
#include "uivr_env.cuh"

namespace uivr {

// tuning knobs (overridable at build time for sweeps: scripts/sweep_pool.sh)
#ifndef UIVR_POOL_BLOCK_ADJ
#define UIVR_POOL_BLOCK_ADJ 896
#endif
#ifndef UIVR_POOL_HANDLERS_ADJ
#define UIVR_POOL_HANDLERS_ADJ 12
#endif
#ifndef UIVR_POOL_SLOTS_ADJ
#define UIVR_POOL_SLOTS_ADJ 1088
#endif
#ifndef UIVR_POOL_BLOCK_DRT
#define UIVR_POOL_BLOCK_DRT 896
#endif
#ifndef UIVR_POOL_HANDLERS_DRT
#define UIVR_POOL_HANDLERS_DRT 12
#endif
#ifndef UIVR_POOL_SLOTS_DRT
#define UIVR_POOL_SLOTS_DRT 928
#endif
#ifndef UIVR_POOL_BLOCK_FWD
#define UIVR_POOL_BLOCK_FWD 896
#endif
#ifndef UIVR_POOL_HANDLERS_FWD
#define UIVR_POOL_HANDLERS_FWD 12
#endif
#ifndef UIVR_POOL_SLOTS_FWD
#define UIVR_POOL_SLOTS_FWD 1280
#endif
#ifndef UIVR_POOL_QUANTUM
#define UIVR_POOL_QUANTUM 16
#endif
#ifndef UIVR_POOL_MINBATCH
#define UIVR_POOL_MINBATCH 12
#endif
#ifndef UIVR_POOL_STARVE
#define UIVR_POOL_STARVE 32
#endif
#ifndef UIVR_POOL_SETMAXNREG
#define UIVR_POOL_SETMAXNREG 1
#endif
#ifndef UIVR_POOL_WALKER_REGS
#define UIVR_POOL_WALKER_REGS 56
#endif
#ifndef UIVR_POOL_INLINE_SPAWN
#define UIVR_POOL_INLINE_SPAWN 1   // emitter / phase sampling right after the vertex / NEE-end handler (no queue hop)
#endif
#ifndef UIVR_POOL_FOCUS
#define UIVR_POOL_FOCUS 1          // 1: the handler warps of a CTA prefer to serve the same queue (instruction cache)
#endif
// work items a CTA reserves from the global queue at a time (0: one global atomic per batch of 32).  Measured (profiles/
// r02_history.md, call V): the forward kernel gains 5 % at 2048, the backward kernels lose 2 % at any size -- their
// handler warps are the bottleneck, and a visit that finds the reservation empty while another warp refills it is a
// wasted visit
#ifndef UIVR_POOL_SCATTER_AT_END
#define UIVR_POOL_SCATTER_AT_END 1   // the path-end handler scatters the first described vertex itself (one queue hop less per path)
#endif
#ifndef UIVR_POOL_CHUNK_FWD
#define UIVR_POOL_CHUNK_FWD 2048
#endif
#ifndef UIVR_POOL_CHUNK_ADJ
#define UIVR_POOL_CHUNK_ADJ 0
#endif
#ifndef UIVR_POOL_CHUNK_DRT
#define UIVR_POOL_CHUNK_DRT 0
#endif
#ifndef UIVR_POOL_SMEMTAB
#define UIVR_POOL_SMEMTAB 0   // 1 (A/B build): the whole walk table lives in shared memory, loaded once per CTA with
#endif                        //    cp.async.bulk (TMA bulk copy) + mbarrier; needs small pools (the table is 166 KB at 256^3 / 8)
#ifndef UIVR_POOL_HANDLERS_LAST
#define UIVR_POOL_HANDLERS_LAST 0   // 1: the handler warps are the LAST warps of the CTA (scheduler priority A/B)
#endif
#ifndef UIVR_POOL_SUBSTEPS
#define UIVR_POOL_SUBSTEPS 1
#endif
constexpr int kWalkQuantum = UIVR_POOL_QUANTUM;   // a walker warp hands its finished lanes on once this many have finished
constexpr int kWalkSubSteps = UIVR_POOL_SUBSTEPS; // supergrid cells per lane between two warp votes
// walker lane states
enum : int { W_IDLE = 0, W_WALKING = 1, W_HIT = 2, W_END = 3 };
constexpr unsigned kPoolEmpty = 0xFFFFu;
constexpr int kPoolMinBatch = UIVR_POOL_MINBATCH;    // smallest handler batch while the walk queue runs dry
constexpr int kPoolStarveBelow = UIVR_POOL_STARVE;   // "runs dry": fewer walk jobs than this are queued
constexpr unsigned kPoolHandlerSleepMax = 512;   // ns; idle handler warps back off up to this
#ifndef UIVR_POOL_WALKER_SLEEP_MAX
#define UIVR_POOL_WALKER_SLEEP_MAX 512
#endif
constexpr unsigned kPoolWalkerSleepMax = UIVR_POOL_WALKER_SLEEP_MAX;  // ns; the same for idle walker warps
constexpr int kPoolSpinLimit = 1 << 22;          // watchdog: mailbox spins
constexpr int kPoolWalkLimit = 1 << 24;          // watchdog: iterations of one walk quantum
constexpr long long kPoolIdleLimit = 4000000000ll;  // watchdog: cycles without any progress of a warp

enum : int { Q_FREE = 0, Q_WALK, Q_TAP, Q_VERTEX, Q_VERTEX_ADJ, Q_SPAWN, Q_NEE_END, Q_PATH_END, Q_NUM };
// KIND_ADJ only: a finished path waits here while its deferred gradient scatter runs, one vertex per visit (the
// adjoint kernel has no plain Q_VERTEX traffic: all its vertices go through Q_VERTEX_ADJ)
constexpr int Q_SCATTER = Q_VERTEX;
enum : int { PM_DELTA = 0, PM_NEE, PM_NEE_ADJ, PM_DRT };
enum : int { PP_PRIMAL = 0, PP_ADJ, PP_DRTV, PP_REC };

// pool fields (one 32-bit word per slot each).  Shared memory spent on the pool is L1 taken from
// the supergrid / sigma_t taps (the carve-out is shared), so fields with disjoint lifetimes alias:
//   F_TS   sigma_t at the real collision (tap -> vertex) | transmittance T (spawn / tap -> NEE end; DRT walk)
//          | sum of the NEE adjoint (NEE end -> replay walk)
//   F_SEQ  PCG32 stream selector v1: inc = (v1 << 1) | 1   (the 64-bit increment is not stored)
//   pixel = idx / spp is recomputed instead of stored
//   DRT reservoir sums (live during the adjoint replay) | Li and albedo of the DRT vertex (after it)
//   sampler clone of the NEE replay (adjoint replay)     | T of the replay walk | DRT distance-sampling result
enum : int {
    F_IDX = 0, F_RNG_LO, F_RNG_HI, F_SEQ,
    F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ,
    F_B0, F_B1, F_B2, F_R0, F_R1, F_R2, F_TS, F_FLAGS, F_DEPTH,
    F_NUM_FWD,
    // state of both backward kernels: dL of the pixel, reservoir weight sums (adjoint) | Li of the DRT vertex (DRT),
    // logged collisions of the current NEE walk (adjoint) | sigma_t at the DRT vertex (DRT)
    F_DL0 = F_NUM_FWD, F_DL1, F_DL2,
    F_RSW0, F_RSW1, F_RSW2,
    F_DRT_ST,
    // alt sampler (adjoint) | albedo of the DRT vertex (DRT)
    F_ALT_LO, F_ALT_HI, F_ASEQ,
    F_NUM_ADJ,
    // DRT kernel only: the reservoir segment, DRT distance-sampling sums
    F_RSOX = F_NUM_ADJ, F_RSOY, F_RSOZ, F_RSDX, F_RSDY, F_RSDZ, F_DRT_D, F_DRT_T,
    F_NUM_DRT,
    // aliases
    F_ST = F_TS, F_T = F_TS, F_ASUM = F_TS,
    F_LI0 = F_RSW0, F_LI1 = F_RSW1, F_LI2 = F_RSW2, F_AL0 = F_ALT_LO, F_AL1 = F_ALT_HI, F_AL2 = F_ASEQ,
    F_NLOG = F_DRT_ST,  // adjoint kernel: tentative collisions of the current NEE walk
    F_T2 = F_DRT_ST     // adjoint kernel: T of the NEE replay walk (log overflow only; the log count is consumed by then)
};
// The adjoint kernel keeps what is touched once or twice per path OUT of shared memory (every word per slot is
// in-flight samples lost): the reservoir's candidate (weight, segment) is the vertex descriptor of the depth the
// reservoir picked + a fifth vector written when a candidate is accepted; the sampler clone of the NEE replay lives in
// the slot's collision log (entry kNeeLog).

// The CONTINUATION of the free-flight walk is a 12-word record per slot (array of structures, 48 bytes): a
// walker lane picks it up with three 128-bit loads (conflict-free for any 8 slots per phase: the stride of 12
// words maps 8 consecutive quarter-warps onto distinct bank quads) and saves a tentative collision with one
// 128-bit + one 64-bit store.
enum : int {
    C_TNX = 0, C_TNY, C_TNZ, C_TAU,    // boundary times of the three axes (as seen from the NEXT cell), remaining optical depth
    C_ADX, C_ADY, C_ADZ, C_TMAX,       // their increments per cell, segment end t_exit
    C_CI, C_WT, C_CIN, C_TCUR,         // current cell | octant << 28, position t, next cell | end queue << 28, time the
    C_WORDS                            // ray leaves the current cell
};

// F_FLAGS bits
enum : unsigned {
    FL_PASS_MASK = 3u, FL_MODE_SHIFT = 2, FL_MODE_MASK = 3u << 2,
    FL_DID_SCATTER = 1u << 4, FL_ESCAPED = 1u << 5, FL_HAS_SCATTERED = 1u << 6, FL_ACTIVE = 1u << 7,
    FL_NEE_VALID = 1u << 8, FL_RS_VALID = 1u << 9, FL_DRT_FOUND = 1u << 10, FL_SPAWN_PHASE = 1u << 11,
    FL_ENDQ_SHIFT = 12, FL_ENDQ_MASK = 7u << 12   // queue the slot goes to when its walk ends
};
// F_CI: padded linear supergrid cell index | octant << 28
constexpr unsigned kCiMask = 0x0FFFFFFFu;

struct PoolCtl {
    unsigned head[Q_NUM];
    unsigned tail[Q_NUM];
    int count[Q_NUM];
    int live;        // slots that may still carry work
    int exhausted;   // the global sample queue is empty
    int abort;       // watchdog tripped: every warp leaves
    int focus;       // (UIVR_POOL_FOCUS) queue the handler warps currently prefer
    unsigned long long chunk;  // (UIVR_POOL_CHUNK) the CTA's reservation from the global queue: next item << 24 | items left
    int refill;      // lock: one warp at a time reserves the next chunk
};
static_assert(UIVR_POOL_CHUNK_FWD < (1 << 24) && UIVR_POOL_CHUNK_ADJ < (1 << 24) && UIVR_POOL_CHUNK_DRT < (1 << 24),
              "24 bits count the items left of a chunk");

// One decision of the supergrid DDA: the axis whose boundary the ray crosses first (ties: x before y before z, as
// in the oracle's walk_next), the time of that crossing (= when the ray leaves the current cell) and the cell
// behind it; the boundary time of that axis moves on by one cell.
UIVR_DEV void walk_decide(float& tnx, float& tny, float& tnz, float adx, float ady, float adz, int sxl, int syl,
                          int szl, int ci, float& tcur, int& cin) {
    const bool yx = tny < tnx;
    float tn = yx ? tny : tnx;
    const bool zb = tnz < tn;
    tn = zb ? tnz : tn;
    tcur = tn;
    cin = ci + (zb ? szl : (yx ? syl : sxl));
    tnx = (!zb && !yx) ? tnx + adx : tnx;
    tny = (!zb && yx) ? tny + ady : tny;
    tnz = zb ? tnz + adz : tnz;
}

// sampler.seed(seed, wavefront) for one lane, out of line: TEA + the PCG32 seeding sequence are ~150 instructions
// and FETCH / ray-batch generation need them at several places (instruction-cache footprint of the handlers)
#ifndef UIVR_POOL_SEED_INLINE
#define UIVR_POOL_SEED_INLINE 1   // (the out-of-line call cost 1.5 %: call AC)
#endif
#if UIVR_POOL_SEED_INLINE
UIVR_DEV void pool_seed_sampler(Rng& r, uint32_t seed, uint32_t idx) { r.seed_sampler(seed, idx); }
#else
__device__ __noinline__ void pool_seed_sampler(Rng& r, uint32_t seed, uint32_t idx) { r.seed_sampler(seed, idx); }
#endif

// the slot-pool backward always scatters into the accumulation buffers (uivr_api.cu sets them up with the launch)
#ifndef UIVR_POOL_ACC
#define UIVR_POOL_ACC 1   // 0 (A/B build): keep the scalar / v2 fall-back paths of the scatter in the pool kernels
#endif
constexpr bool kPoolAcc = UIVR_POOL_ACC && UIVR_DSIGMA_TILED && UIVR_DALBEDO_V4 && !UIVR_SCATTER_MATCH;
UIVR_DEV void batch_film_position(const Params& P, uint32_t b, uint32_t idx, float F[15], float& u, float& v) {
    batch_film_position_t(P, b, idx, F, u, v, [](Rng& r, uint32_t sd, uint32_t i) { pool_seed_sampler(r, sd, i); });
}

// Watchdog record: the thread that trips a limit raises the CTA's abort flag and leaves its reason in the debug buffer;
// the queue state is dumped at the end of the kernel by the CTA that tripped first.  Out of line by measurement
// (call AG: 786 vs 775 Msamples/s inline), like the sampler seeding is inline by measurement -- at this code size
// either choice moves the register allocation and the layout of the handler bodies by a percent.
#ifndef UIVR_POOL_TRIP_INLINE
#define UIVR_POOL_TRIP_INLINE 0
#endif
#if UIVR_POOL_TRIP_INLINE
UIVR_DEV
#else
__device__ __noinline__
#endif
void pool_trip(PoolCtl* ctl, unsigned* debug, unsigned why) {
    if (atomicExch(&ctl->abort, 1) == 0 && debug) {
        if (atomicExch(&debug[0], why) == 0u) {
            debug[1] = blockIdx.x;
            debug[2] = threadIdx.x;
        }
    }
}
__device__ __noinline__ void pool_trip_dump(const PoolCtl* ctl, unsigned* debug) {
    for (int q = 0; q < Q_NUM; ++q) {
        debug[4 + 3 * q] = ctl->head[q];
        debug[5 + 3 * q] = ctl->tail[q];
        debug[6 + 3 * q] = (unsigned) ctl->count[q];
    }
    debug[4 + 3 * Q_NUM] = (unsigned) ctl->live;
    debug[5 + 3 * Q_NUM] = (unsigned) ctl->exhausted;
}

// kernel kinds: the forward / primal kernel and the two halves of the backward pipeline -- the adjoint replay
// (gathers the primal radiance itself, scatters the free-flight / transmittance / NEE gradients, hands one
// reservoir record per sample to HBM), then the DRT pass on those records.  Splitting trades ~1 GB of coalesced
// HBM traffic (the path is at < 15 % of the HBM roofline) for lean kernels with fewer live modes each.
enum : int { KIND_FWD = 0, KIND_ADJ = 2, KIND_DRT = 3 };
constexpr int kRecWords = 16;  // reservoir record: seg(7) dL'(3) alt state(2) alt seq(1) depth(1) pad(2)
// Vertex descriptor of the adjoint kernel (5 x uint4, global memory, one area of max_depth + 1 descriptors per
// slot): what the free-flight (:152-172) and transmittance (:181-189) gradients of one path segment need apart from
// the radiance that is still to come -- which is only known when the path has ended.
//   {alt state lo, hi, sigma_t at the collision (0: the segment escaped), interval} {o.xyz, d.x} {d.yz, c.xy}
//   {c.z, albedo.xyz}    c = contribution of the next-event estimation that follows this vertex (written by the
//                        NEE handler; 0 without one): what the reference subtracts from L after the vertex (:214)
//   {reservoir weight.xyz, t_exit of the segment}   (only written when the reservoir accepts the vertex, :745-753)
constexpr int kDescVec = 5;
// NEE adjoint (volpathsimple.py:393-401, :483-492): the reference walks every shadow segment a second time, from a
// cloned sampler, to scatter -sum(adjoint)/sigma_n at each tentative collision once the contribution is known.
// The adjoint kernel instead LOGS the tentative collisions of the first walk (t, sigma_n; kNeeLog per slot, global
// memory) and scatters from the log at the NEE end: same positions, same values, no second walk.  A shadow walk
// with more collisions than the log holds falls back to the replay walk.
constexpr int kNeeLog = 32;
constexpr int kNeeLogStride = kNeeLog + 1;  // + the sampler clone (:383) for the fall-back

// ENV (envmap emitter, uivr_env.cuh): three more fields per slot hold the NEE weight
// throughput * phase * mis * Le / pdf of the direction sampled at Q_SPAWN until Q_NEE_END
__host__ __device__ constexpr int pool_fields(int kind) { return kind == 0 ? F_NUM_FWD : kind == 2 ? F_NUM_ADJ : F_NUM_DRT; }
template <int KIND, int NSLOT, bool ENV = false>
constexpr size_t pool_smem_bytes() {
    return 128 + (size_t) Q_NUM * NSLOT * sizeof(uint16_t) +
           (size_t) (pool_fields(KIND) + (ENV ? 3 : 0) + C_WORDS) * NSLOT * sizeof(uint32_t);
}
// The envmap instances carry three more words per slot: as many slots as fit the footprint of the constant-emitter
// instance (the shared-memory carve-out moves in steps; one step more halves the L1 that is left)
template <int KIND, int NSLOT>
constexpr int pool_env_slots() {
    int n = NSLOT;
    while (n > 256 && 128 + (size_t) n * (Q_NUM * sizeof(uint16_t) + (pool_fields(KIND) + 3 + C_WORDS) * sizeof(uint32_t)) >
                          pool_smem_bytes<KIND, NSLOT, false>())
        n -= 32;
    return n;
}

// BLOCK threads per CTA (one CTA per SM), of which HANDLERS warps serve the transition queues and the
// rest walk
// BATCH: ray-batch mode (uivr_set_batch): its ray generation is compiled into its own instances, the sensor-mode
// kernels do not carry it (every instruction of the handler bodies counts: profiles/r02_history.md, call AB)
template <int KIND, bool COUNT, int NSLOT, int BLOCK, int HANDLERS, bool ENV = false, bool BATCH = false>
__global__ void __launch_bounds__(BLOCK, 1) k_pool(const Params P) {
    constexpr bool BWD = KIND != KIND_FWD;               // any gradient work
    constexpr int F_NW0 = pool_fields(KIND);             // ENV only: NEE weight (3 words)
    constexpr bool HAS_ADJ = KIND == KIND_ADJ;           // adjoint replay (reservoir, NEE adjoint)
    constexpr bool HAS_DRT = KIND == KIND_DRT;           // DRT walk, DRT vertex, recursive path
    // (the shared-memory-table A/B build keeps its mbarrier where the chunk lock lives)
    constexpr unsigned kPoolChunk = UIVR_POOL_SMEMTAB ? 0u : (unsigned) (KIND == KIND_FWD ? UIVR_POOL_CHUNK_FWD : KIND == KIND_ADJ ? UIVR_POOL_CHUNK_ADJ : UIVR_POOL_CHUNK_DRT);
    constexpr int kPoolBlock = BLOCK;
    constexpr int kPoolHandlerWarps = HANDLERS;
    static_assert(NSLOT > (Q_NUM - 1) * 31 && NSLOT % 32 == 0, "pool too small for the full-batch scheduling rule");
    static_assert(sizeof(PoolCtl) <= 128, "PoolCtl must fit its 128-byte header");
    static_assert(Q_NUM <= 8, "the handler scheduler packs the queue id into 3 bits");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PoolCtl* const ctl = reinterpret_cast<PoolCtl*>(smem_raw);
    // one 16-bit mailbox cell per ring position.  A position has exactly one producer and one consumer per lap
    // (both reserved it through the atomic tail / head counters), so the cells need no atomics: the producer
    // waits for kPoolEmpty and stores the id, the consumer waits for an id and stores kPoolEmpty back.
    volatile uint16_t* const ring = reinterpret_cast<volatile uint16_t*>(smem_raw + 128);
    uint32_t* const pool = reinterpret_cast<uint32_t*>(smem_raw + 128 + Q_NUM * NSLOT * sizeof(uint16_t));
    static_assert(NSLOT < 65535 && (Q_NUM * NSLOT * sizeof(uint16_t)) % 16 == 0, "16-bit slot ids; the pool stays 16-byte aligned");
    uint32_t* const cont = pool + (size_t) (pool_fields(KIND) + (ENV ? 3 : 0)) * NSLOT;  // [slot][C_WORDS]
    static_assert(NSLOT % 4 == 0, "the continuation records must stay 16-byte aligned");
#if UIVR_POOL_SMEMTAB
    // walk table in shared memory (+ its slack on either side): one elected thread issues TMA bulk copies, the bytes
    // land asynchronously and complete the transaction count of an mbarrier every thread then waits on
    uint32_t* const wtab_s = cont + (size_t) NSLOT * C_WORDS + P.wtab_slack;   // cell 0
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(smem_raw + 120);
#define WT(ci) wtab_s[(ci)]
#else
#define WT(ci) __ldg(P.wtab + (ci))
#endif

    Counters<COUNT> K;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint64_t total = KIND == KIND_DRT ? (uint64_t) *P.rec_count : (uint64_t) P.n_slots * P.spp;
    const bool use_rsv = BWD && P.use_drt && P.use_drt_subsampling;

#define PU(f, s) pool[(f) * NSLOT + (s)]
#define PF(f, s) __uint_as_float(pool[(f) * NSLOT + (s)])
#define PSET(f, s, v) pool[(f) * NSLOT + (s)] = __float_as_uint(v)
#define CU(k, s) cont[(s) * C_WORDS + (k)]
#define CF(k, s) __uint_as_float(cont[(s) * C_WORDS + (k)])
#define CSET(k, s, v) cont[(s) * C_WORDS + (k)] = __float_as_uint(v)

    // ---- pool / queue initialisation: every slot starts in Q_FREE ----
    for (int i = threadIdx.x; i < Q_NUM * NSLOT; i += kPoolBlock)
        ring[i] = (uint16_t) ((i < NSLOT) ? (unsigned) i : kPoolEmpty);   // ring 0 == Q_FREE
    for (int i = threadIdx.x; i < NSLOT; i += kPoolBlock) PU(F_FLAGS, i) = 0u;
    if (threadIdx.x < Q_NUM) {
        ctl->head[threadIdx.x] = 0u;
        ctl->tail[threadIdx.x] = threadIdx.x == Q_FREE ? (unsigned) NSLOT : 0u;
        ctl->count[threadIdx.x] = threadIdx.x == Q_FREE ? NSLOT : 0;
    }
    if (threadIdx.x == 0) {
        ctl->live = NSLOT; ctl->exhausted = 0; ctl->abort = 0; ctl->focus = Q_FREE;
        ctl->chunk = 0ull; ctl->refill = 0;
    }
#if UIVR_POOL_SMEMTAB
    {
        const uint32_t mb = (uint32_t) __cvta_generic_to_shared(mbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t) P.wtab_words * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            const char* src = reinterpret_cast<const char*>(P.wtab - P.wtab_slack);
            const uint32_t dst = (uint32_t) __cvta_generic_to_shared(wtab_s - P.wtab_slack);
            for (uint32_t off = 0; off < bytes; off += 16384u) {
                const uint32_t n = bytes - off < 16384u ? bytes - off : 16384u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst + off), "l"(src + off), "r"(n), "r"(mb) : "memory");
            }
        }
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(mb) : "memory");
    }
#endif
    __syncthreads();

    // ---- queue primitives (warp-collective; each is instantiated ONCE per role to keep the code small) ----
    // watchdog: a tripped limit records the reason + queue state and makes every warp leave
    auto trip = [&](unsigned why) { pool_trip(ctl, P.debug, why); };
    // pops up to `want` ids from queue q (all-or-nothing when `exact`); lane i < got receives one
    auto q_pop = [&](int q, int want, bool exact, unsigned& slot) -> int {
        int got = 0;
        unsigned base = 0;
        if (lane == 0) {
            const int old = atomicSub(&ctl->count[q], want);
            got = old >= want ? want : ((exact || old <= 0) ? 0 : old);
            if (got < want) atomicAdd(&ctl->count[q], want - got);
            if (got) base = atomicAdd(&ctl->head[q], (unsigned) got);
        }
        got = __shfl_sync(FULL, got, 0);
        base = __shfl_sync(FULL, base, 0);
        if ((int) lane < got) {
            unsigned pos = base % (unsigned) NSLOT + lane;  // (head runs free; the ring has NSLOT cells)
            pos -= pos >= (unsigned) NSLOT ? (unsigned) NSLOT : 0u;
            volatile uint16_t* cell = &ring[q * NSLOT + pos];
            unsigned v;
            int spins = 0;
            // the position is reserved by a producer; its store may still be in flight
            while ((v = *cell) == kPoolEmpty) {
                if (++spins > kPoolSpinLimit) { trip(0x200u + (unsigned) q); v = 0; break; }
            }
            *cell = (uint16_t) kPoolEmpty;
            slot = v;
        }
        __threadfence_block();
        return got;
    };

    // route: hand every finished slot to its next queue (lane-wise: slot `s` -> queue `next`, -1: none)
    auto route = [&](unsigned s, int next) {
        unsigned todo = __ballot_sync(FULL, next >= 0);
        __threadfence_block();  // pool fields before the ids become visible
        while (todo) {
            const int q = __shfl_sync(FULL, next, __ffs(todo) - 1);
            const unsigned m = __ballot_sync(FULL, next == q);
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if ((int) lane == leader) base = atomicAdd(&ctl->tail[q], (unsigned) __popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (next == q) {
                unsigned pos = base % (unsigned) NSLOT + __popc(m & lt_mask);
                pos -= pos >= (unsigned) NSLOT ? (unsigned) NSLOT : 0u;
                volatile uint16_t* cell = &ring[q * NSLOT + pos];
                int spins = 0;
                // the cell is free unless the consumer of the previous lap has not taken its id yet
                while (*cell != kPoolEmpty) {
                    if (++spins > kPoolSpinLimit) { trip(0x100u + (unsigned) q); break; }
                }
                *cell = (uint16_t) s;
            }
            __threadfence_block();  // the ids before the count that makes them poppable
            __syncwarp();
            if ((int) lane == leader) atomicAdd(&ctl->count[q], __popc(m));
            todo &= ~m;
        }
    };

    // Warp roles.  The two loops share no registers, so the walker loop (the hot code: ~100 instructions
    // that stay in the L0 instruction cache) is not charged for the handlers' working set.
    const int warp_id = threadIdx.x >> 5;
    const bool is_handler = UIVR_POOL_HANDLERS_LAST ? warp_id >= kPoolBlock / 32 - kPoolHandlerWarps
                                                    : warp_id < kPoolHandlerWarps;
    long long t_progress = clock64();

    // Register re-allocation between the roles (sm_90a+ `setmaxnreg`, per warpgroup of 4 warps): a walker
    // lane needs ~30 registers, a handler lane wants ~100; launched with 64 per thread (1024 threads), the
    // walkers give registers up and the handlers take them.  HANDLERS must be a multiple of 4.
#if UIVR_POOL_SETMAXNREG
    static_assert(HANDLERS % 4 == 0 && (BLOCK / 32 - HANDLERS) % 4 == 0 && BLOCK % 128 == 0,
                  "setmaxnreg works on aligned warpgroups of 4 warps");
    // the CTA's register pool is what it was launched with: (registers per thread under __launch_bounds__(BLOCK, 1),
    // a multiple of 8) x BLOCK; the walkers keep kRegWalker each, the handlers share the rest
    constexpr int kRegLaunch = (65536 / BLOCK) / 8 * 8 > 255 ? 248 : (65536 / BLOCK) / 8 * 8;
    constexpr int kRegWalker = UIVR_POOL_WALKER_REGS;
    constexpr int kRegHandlerRaw = ((kRegLaunch * (BLOCK / 32) - (BLOCK / 32 - HANDLERS) * kRegWalker) / HANDLERS) / 8 * 8;
    constexpr int kRegHandler = kRegHandlerRaw > 232 ? 232 : kRegHandlerRaw;
    static_assert(kRegWalker <= kRegLaunch && kRegHandler >= kRegLaunch, "walkers give registers up, handlers take them");
#endif
    if (!is_handler) {
#if UIVR_POOL_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegWalker));
#endif
        // ==================================================================================
        // WALKER WARPS: supergrid DDA of the free-flight walk (Medium::sample_interaction, App. B.5)
        // ==================================================================================
        int wslot = -1;            // pool slot this lane walks for, -1 = idle
        int wst = W_IDLE;
        int wendq = 0;             // queue of the slot when the walk ends without a collision
        float tnx = 0.0f, tny = 0.0f, tnz = 0.0f, adx = 0.0f, ady = 0.0f, adz = 0.0f;
        float tau = 0.0f, wt = 0.0f, tmax = 0.0f;
        float tcur = 0.0f;         // time the ray leaves the current cell
        int ci = 0, cin = 0, sxl = 0, syl = 0, szl = 0;
        unsigned e = 0u, en = 0u, obit = 0u, oct = 0u;   // walk-table words of the current / the next cell
        unsigned widle_ns = 64;
        int wphase = 0, wcur = 0;  // register-pair roles of the DDA loop (see dda_step)
        for (;;) {
            if (__shfl_sync(FULL, *((volatile int*) &ctl->abort), 0)) break;
            // ---- 1. idle lanes pick up walk continuations from Q_WALK ----
            const unsigned idle = __ballot_sync(FULL, wslot < 0);
            if (idle) {
                const int cnt = __shfl_sync(FULL, (lane == 0) ? *((volatile int*) &ctl->count[Q_WALK]) : 0, 0);
                if (cnt > 0) {
                    unsigned got_slot = 0;
                    const int got = q_pop(Q_WALK, __popc(idle), false, got_slot);
                    if (got) {
                        const int rank = __popc(idle & lt_mask);
                        const unsigned s = __shfl_sync(FULL, got_slot, rank & 31);
                        if (wslot < 0 && rank < got) {
                            wslot = (int) s;
                            const uint4* rec = reinterpret_cast<const uint4*>(cont + s * C_WORDS);
                            const uint4 ra = rec[0], rb = rec[1], rc = rec[2];
                            tnx = __uint_as_float(ra.x); tny = __uint_as_float(ra.y); tnz = __uint_as_float(ra.z);
                            tau = __uint_as_float(ra.w);
                            adx = __uint_as_float(rb.x); ady = __uint_as_float(rb.y); adz = __uint_as_float(rb.z);
                            tmax = __uint_as_float(rb.w);
                            ci = (int) (rc.x & kCiMask);
                            oct = rc.x >> 28;
                            wt = __uint_as_float(rc.y);
                            cin = (int) (rc.z & kCiMask);
                            wendq = (int) (rc.z >> 28);
                            tcur = __uint_as_float(rc.w);
                            e = WT(ci);
                            en = WT(cin);
                            obit = 1u << oct;
                            const int px = P.pm[0], pxy = P.pm[0] * P.pm[1];
                            sxl = (oct & 1u) ? -1 : 1;
                            syl = (oct & 2u) ? -px : px;
                            szl = (oct & 4u) ? -pxy : pxy;
                            wst = W_WALKING;
                        }
                    }
                }
            }
            const unsigned m_walk = __ballot_sync(FULL, wst == W_WALKING);
            if (!m_walk) {
                if (__shfl_sync(FULL, *((volatile int*) &ctl->live), 0) <= 0) break;
                if (__shfl_sync(FULL, clock64() - t_progress > kPoolIdleLimit ? 1 : 0, 0)) { trip(0x300u); break; }
                __nanosleep(widle_ns);
                widle_ns = widle_ns < kPoolWalkerSleepMax ? widle_ns * 2 : widle_ns;  // back off: polling costs issue slots
                continue;
            }
            widle_ns = 64;
            t_progress = clock64();
            // ---- 2. step cells until kWalkQuantum lanes have finished (or a short-handed warp can refill) ----
            const int n0 = __popc(m_walk);
            const int n_stop = n0 > kWalkQuantum ? n0 - kWalkQuantum : 0;
            const bool short_handed = n0 <= 32 - kWalkQuantum;
            // One DDA step of this lane (ec / cc: walk-table word and index of the current cell, en / cn: of the next
            // one, whose word was requested one step earlier).  The caller alternates the roles of the two register
            // pairs from step to step, so that entering the next cell moves no registers.
            auto dda_step = [&](unsigned& ec, unsigned& en_, int& cc, int& cn) {
                if (wst == W_WALKING) {
                    // empty cell with only empty cells ahead (or the border): no further collision
                    const bool gone = (int) ec < 0 && (ec & obit) != 0u;
                    const bool more = tcur < tmax;
                    const float t_end = more ? tcur : tmax;
                    const float len = fmaxf(t_end - wt, 0.0f);
                    const bool dense = (int) ec > 0;
                    const float dtau = __uint_as_float(ec) * len;
                    // tau is used up inside this cell: the tap handler takes over
                    const bool hit = dense && tau < dtau;
                    if (gone || hit) {
                        wst = gone ? W_END : W_HIT;
                        wcur = &ec == &e ? 0 : 1;
                    } else {
                        tau = dense ? tau - dtau : tau;
                        wt = fmaxf(wt, t_end);
                        if (more) {
                            // enter the next cell (its word is en_); decide the step after it and request that
                            // cell's word now: the load latency hides behind a whole step
                            if (COUNT && (en_ & 0x80000100u) != 0x80000100u) K.add(C_MAJ, 1);
                            walk_decide(tnx, tny, tnz, adx, ady, adz, sxl, syl, szl, cn, tcur, cc);
                            ec = WT(cc);
                        } else {
                            wst = W_END;  // t_exit reached
                        }
                    }
                }
            };
            for (int it = 1;; ++it) {
                // even steps: (e, ci) current, (en, cin) next; odd steps: the other way round
                dda_step(e, en, ci, cin);
                int n_walk = __popc(__ballot_sync(FULL, wst == W_WALKING));
                if (n_walk <= n_stop) { wphase = 1; break; }
                dda_step(en, e, cin, ci);
                n_walk = __popc(__ballot_sync(FULL, wst == W_WALKING));
                if (n_walk <= n_stop) { wphase = 0; break; }
                if ((it & 3) == 0) {
                    if (it > P.walk_limit) { trip(0x400u); wst = W_END; wphase = 0; break; }
                    // a warp that started short of lanes leaves as soon as the queue can fill them
                    if (short_handed &&
                        __shfl_sync(FULL, (lane == 0) ? *((volatile int*) &ctl->count[Q_WALK]) : 0, 0) >= kWalkQuantum) {
                        wphase = 0;
                        break;
                    }
                }
            }
            // Back to the canonical naming ((e, ci) current, (en, cin) next) for the lanes that walk on; a lane that
            // stopped (hit / end) recorded which pair held its current cell.
            {
                const bool swap = wst == W_WALKING ? wphase == 1 : wcur == 1;
                if (swap) {
                    const unsigned te = e; e = en; en = te;
                    const int tc = ci; ci = cin; cin = tc;
                }
            }
            // ---- 3. hand finished lanes on: a tentative collision saves the words that changed ----
            unsigned s = 0;
            int next = -1;
            if (wslot >= 0 && wst != W_WALKING) {
                s = (unsigned) wslot;
                if (wst == W_HIT) {
                    uint32_t* rec = cont + s * C_WORDS;
                    *reinterpret_cast<uint4*>(rec) = make_uint4(__float_as_uint(tnx), __float_as_uint(tny), __float_as_uint(tnz),
                                                                __float_as_uint(tau));
                    *reinterpret_cast<uint4*>(rec + C_CI) = make_uint4((unsigned) ci | (oct << 28), __float_as_uint(wt),
                                                                       (unsigned) cin | ((unsigned) wendq << 28),
                                                                       __float_as_uint(tcur));
                    next = Q_TAP;
                } else {
                    next = wendq;
                }
                wslot = -1;
                wst = W_IDLE;
            }
            route(s, next);
        }
    } else {
        // ==================================================================================
        // HANDLER WARPS: a FULL batch of a transition queue if there is one; partial batches only
        // once the global sample queue is exhausted or while the walkers run dry.  Before exhaustion no
        // slot retires, so "no full queue and nothing to walk anywhere" cannot happen: NSLOT > (Q_NUM - 1) * 31
        // slots cannot all sit in non-full queues.
        // ==================================================================================
#if UIVR_POOL_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegHandler));
#endif
        int last_work = Q_FREE;
        unsigned idle_ns = 32;
        for (;;) {
            // one shared-memory read per lane: the fill level of queue `lane` (lanes 0..7), the queue the handler warps
            // currently prefer (lane 8), the abort flag (lane 9)
            const int* cw = lane < Q_NUM ? &ctl->count[lane] : (lane == 8 ? &ctl->focus : &ctl->abort);
            const int cv = lane <= 9 ? *((volatile const int*) cw) : 0;
            if (__shfl_sync(FULL, cv, 9)) break;
            const int cnt_all = (lane < Q_NUM) ? cv : 0;
            const int cnt = (lane != Q_WALK) ? cnt_all : 0;
            int work = -1;
            bool exact = true;
#if UIVR_POOL_FOCUS
            // Fast path (two dependent shuffles): the preferred queue still holds a full batch.  All handler warps of
            // the CTA on the same handler body is what the instruction caches like (profiles/r02_history.md).
            const int focus = __shfl_sync(FULL, cv, 8) & 7;
            if (__shfl_sync(FULL, cnt, focus) >= 32) {
                work = focus;
            } else
#endif
            {
                // the walkers are running dry: accept smaller batches rather than let them idle
                const int min_batch = __shfl_sync(FULL, cnt_all, Q_WALK) < kPoolStarveBelow ? kPoolMinBatch : 32;
                // else the fullest queue (which becomes the preferred one), else the queue served last
                const int c_last = __shfl_sync(FULL, cnt, last_work & 31);
                int best = (cnt << 3) | (int) lane;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
                best = __shfl_sync(FULL, best, 0);
#if UIVR_POOL_FOCUS
                if ((best >> 3) >= 32) {
                    work = best & 7;
                    if (lane == 0) ctl->focus = work;
                } else
#endif
                if (c_last >= 32) {
                    work = last_work;
                } else if ((best >> 3) >= min_batch) {
                    work = best & 7;
                    exact = (best >> 3) >= 32;
                } else if ((best >> 3) > 0 && __shfl_sync(FULL, *((volatile int*) &ctl->exhausted), 0)) {
                    work = best & 7;
                    exact = false;
                }
            }
            if (work < 0) {
                if (__shfl_sync(FULL, *((volatile int*) &ctl->live), 0) <= 0) break;
                if (__shfl_sync(FULL, clock64() - t_progress > kPoolIdleLimit ? 1 : 0, 0)) { trip(0x300u); break; }
                __nanosleep(idle_ns);
                idle_ns = idle_ns < kPoolHandlerSleepMax ? idle_ns * 2 : idle_ns;  // back off: polling costs issue slots
                continue;
            }
            t_progress = clock64();
            idle_ns = 32;
            last_work = work;

            // per-lane result of the work item: slot `s` goes to queue `next` (-1: nothing to route)
            unsigned s = 0;
            int next = -1;
            bool resume = false;  // next == Q_WALK continues a walk after a null collision (no set-up)
            // gradient scatter request of the handlers (executed at one site below)
            bool sc_taps = false, sc_ff = false, sc_alb = true;  // sc_alb: the vertex request also scatters d albedo
            bool scatter_now = false;  // (adjoint kernel) the path has just ended: its first described vertex is due
            float sc_g = 0.0f, sc_int = 0.0f, sc_gs = 0.0f, sc_ga[3] = {0.0f, 0.0f, 0.0f};
            float sc_ox = 0.0f, sc_oy = 0.0f, sc_oz = 0.0f, sc_dx = 0.0f, sc_dy = 0.0f, sc_dz = 0.0f;
            float sc_vx = 0.0f, sc_vy = 0.0f, sc_vz = 0.0f;
            unsigned nee_n = 0u;  // NEE end: logged tentative collisions of the shadow walk to scatter, weight nee_a
            float nee_a = 0.0f;
            Rng alt;
            alt.state = alt.inc = 0;
            // ==========================================================================
            // transition handlers (one batch of up to 32 slots of queue `work`)
            // ==========================================================================
            const int got = q_pop(work, 32, exact, s);
            const bool act = (int) lane < got;
            if (work == Q_TAP) {
                // ---- tentative collision: its position inside the cell, the sigma_t tap and the per-mode
                //      decision (:347-367 delta tracking, :465-502 ratio tracking, App. B.6 DRT) ----
                if (act) {
                    unsigned fl = PU(F_FLAGS, s);
                    const int mode = (int) ((fl & FL_MODE_MASK) >> FL_MODE_SHIFT);
                    const float sb = __uint_as_float(WT(CU(C_CI, s) & kCiMask));
                    // end of the cell along the ray, as the walker saw it
                    const float tn = CF(C_TCUR, s), tmax = CF(C_TMAX, s);
                    const float t_end = tn < tmax ? tn : tmax;
                    const float t = CF(C_WT, s) + CF(C_TAU, s) / sb;
                    const float wt = t > t_end ? t_end : t;
                    CSET(C_WT, s, wt);
                    const float px = fmaf(wt, PF(F_DX, s), PF(F_OX, s)), py = fmaf(wt, PF(F_DY, s), PF(F_OY, s)),
                                pz = fmaf(wt, PF(F_DZ, s), PF(F_OZ, s));
                    Rng r;
                    r.state = (uint64_t) PU(F_RNG_LO, s) | ((uint64_t) PU(F_RNG_HI, s) << 32);
                    r.inc = ((uint64_t) PU(F_SEQ, s) << 1) | 1ull;
                    // one extra draw for delta tracking (accept test, :359) and for DRT (reservoir, drawn
                    // before the lookup); the lookup itself draws nothing, so the order is immaterial
                    float u = 0.0f;
                    if (mode == PM_DELTA || (HAS_DRT && mode == PM_DRT)) u = draw(r, K);
                    const float st = sigma_tap(P, px, py, pz);
                    K.add(C_SIGMA, 1);
                    // sigma_t / sigma_bar (delta tracking) or sigma_n / sigma_bar (ratio tracking, DRT): one
                    // division site for all modes
                    const float sn = sb - st;
                    const float q = (mode == PM_DELTA ? st : sn) / sb;
                    bool go_on = true;
                    if (mode == PM_DELTA) {
                        // :354-361 real vs null collision
                        if (!(u >= q)) {
                            PSET(F_ST, s, st);
                            fl |= FL_DID_SCATTER;
                            go_on = false;
                        }
                    } else {
                        const int ft = (HAS_ADJ && mode == PM_NEE_ADJ) ? F_T2 : F_T;
                        float T = PF(ft, s);
                        if (HAS_DRT && mode == PM_DRT) {
                            // sample_interaction_drt (App. B.6): candidate weight T/sigma_bar, size-1 reservoir
                            const float wi = T / sb;
                            const float D = PF(F_DRT_D, s) + wi;
                            PSET(F_DRT_D, s, D);
                            if (u <= wi / D) {
                                PSET(F_DRT_T, s, wt);
                                PSET(F_DRT_ST, s, st);
                                fl |= FL_DRT_FOUND;
                            }
                        } else if (HAS_ADJ && mode == PM_NEE_ADJ && q > 0.0f) {
                            // ratio tracking adjoint: -sum(adj)/sigma_n (:483-492)   [replay walk: log overflow only;
                            // executed at the scatter site below, which keeps this rare route out of the tap handler]
                            sc_vx = px; sc_vy = py; sc_vz = pz;
                            sc_gs = -PF(F_ASUM, s) / sn;
                            sc_ff = true;
                            sc_alb = false;
                        } else if (HAS_ADJ && mode == PM_NEE && q > 0.0f) {
                            // log the collision for the NEE adjoint (scattered at the NEE end, when its weight is known)
                            const unsigned nl = PU(F_NLOG, s);
                            if (nl < (unsigned) kNeeLog)
                                __stcg(P.neelog + ((size_t) blockIdx.x * NSLOT + s) * kNeeLogStride + nl, make_float2(wt, sn));
                            PU(F_NLOG, s) = nl + 1u;
                        }
                        // ratio tracking (:461-502) / running transmittance of the DRT walk
                        T *= q;
                        PSET(ft, s, T);
                        if ((HAS_DRT && mode == PM_DRT) ? !(T > 0.0f) : T == 0.0f) go_on = false;
                    }
                    if (go_on) {
                        CSET(C_TAU, s, neg_log1m(draw(r, K)));
                        next = Q_WALK;  // the walker resumes in the same cell
                        resume = true;
                    } else {
                        next = (int) ((fl & FL_ENDQ_MASK) >> FL_ENDQ_SHIFT);
                    }
                    PU(F_RNG_LO, s) = (uint32_t) r.state;
                    PU(F_RNG_HI, s) = (uint32_t) (r.state >> 32);
                    PU(F_FLAGS, s) = fl;
                }
            } else if ((!HAS_ADJ && work == Q_VERTEX) || work == Q_VERTEX_ADJ) {   // (adjoint kernel: Q_VERTEX is Q_SCATTER)
                // ---- end of a delta-tracking segment (:130-245) or of the DRT walk (:550-558) ----
                if (act) {
                    unsigned fl = PU(F_FLAGS, s);
                    const unsigned dw = PU(F_DEPTH, s);
                    int depth = (int) (dw & 0xFFFFu);
                    const int pass = (int) (fl & FL_PASS_MASK);
                    const bool is_drt = HAS_DRT && ((fl & FL_MODE_MASK) >> FL_MODE_SHIFT) == (unsigned) PM_DRT;
                    const bool ds = is_drt ? (fl & FL_DRT_FOUND) != 0u : (fl & FL_DID_SCATTER) != 0u;
                    // vertex position: on the stored reservoir segment for DRT, else on the current segment
                    const int fo = (HAS_DRT && is_drt) ? F_RSOX : F_OX, fd = (HAS_DRT && is_drt) ? F_RSDX : F_DX;
                    const float sox = PF(fo, s), soy = PF(fo + 1, s), soz = PF(fo + 2, s);
                    const float sdx = PF(fd, s), sdy = PF(fd + 1, s), sdz = PF(fd + 2, s);
                    const float swt = (HAS_DRT && is_drt) ? PF(F_DRT_T, s) : CF(C_WT, s);
                    const float vx = fmaf(swt, sdx, sox), vy = fmaf(swt, sdy, soy), vz = fmaf(swt, sdz, soz);
                    float albedo[3] = {1.0f, 1.0f, 1.0f};
                    if (ds) {
                        albedo_tap(P, vx, vy, vz, albedo);
                        K.add(C_ALBEDO, 1);
                        // the vertex becomes the origin of whatever leaves it (NEE segment, then the phase-sampled one)
                        PSET(F_OX, s, vx); PSET(F_OY, s, vy); PSET(F_OZ, s, vz);
                    }
                    if (HAS_DRT && is_drt) {
                        if (ds) {
                            PSET(F_AL0, s, albedo[0]); PSET(F_AL1, s, albedo[1]); PSET(F_AL2, s, albedo[2]);
                            PSET(F_LI0, s, 0.0f); PSET(F_LI1, s, 0.0f); PSET(F_LI2, s, 0.0f);
                            PSET(F_B0, s, 1.0f); PSET(F_B1, s, 1.0f); PSET(F_B2, s, 1.0f);
                            fl = (fl & ~FL_PASS_MASK) | (unsigned) PP_DRTV;
                            fl = P.use_nee ? (fl & ~FL_SPAWN_PHASE) : (fl | FL_SPAWN_PHASE);
                            next = Q_SPAWN;
                        } else {
                            next = Q_FREE;
                        }
                    } else {
                        const float stmax = CF(C_TMAX, s);
                        const float beta[3] = {PF(F_B0, s), PF(F_B1, s), PF(F_B2, s)};
                        if (ds) {
                            fl |= FL_HAS_SCATTERED;
                            K.add(C_REAL, 1);
                        }
                        if (HAS_ADJ && pass == PP_ADJ) {
                            alt.state = (uint64_t) PU(F_ALT_LO, s) | ((uint64_t) PU(F_ALT_HI, s) << 32);
                            alt.inc = ((uint64_t) PU(F_ASEQ, s) << 1) | 1ull;
                            const float st = PF(F_ST, s);
                            if (use_rsv) {
                                // DRTReservoir.update (:745-753), weight = throughput before this vertex
                                const float u = draw(alt, K);
                                float wsum[3] = {PF(F_RSW0, s), PF(F_RSW1, s), PF(F_RSW2, s)};
                                float ratio[3];
                                bool took = false;
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    wsum[c] += beta[c];
                                    ratio[c] = beta[c] / wsum[c];
                                }
                                PSET(F_RSW0, s, wsum[0]); PSET(F_RSW1, s, wsum[1]); PSET(F_RSW2, s, wsum[2]);
                                if (u <= mean3(ratio)) {
                                    // the candidate IS this vertex's descriptor (segment) + its fifth vector
                                    took = true;
                                    PU(F_DEPTH, s) = (dw & 0xFFFFu) | ((unsigned) depth << 16);
                                    fl |= FL_RS_VALID;
                                }
                                if (took)
                                    __stcg(P.desc + (((size_t) blockIdx.x * NSLOT + s) * P.desc_cap + depth) * kDescVec + 4,
                                           make_uint4(__float_as_uint(beta[0]), __float_as_uint(beta[1]), __float_as_uint(beta[2]),
                                                      __float_as_uint(stmax)));
                            }
                            // The free-flight (:152-172) and transmittance (:181-189) gradients of this segment need
                            // the radiance the path gathers from here on, Li = L - (what was gathered before).  The
                            // reference gets L from a separate primal pass (batched.py:255-264) and subtracts as it
                            // replays; here the replay itself gathers L, so the vertex is only DESCRIBED now and its
                            // gradients are scattered when the path has ended (Q_SCATTER): no primal replay launch.
                            {
                                uint4* dsc = P.desc + (((size_t) blockIdx.x * NSLOT + s) * P.desc_cap + depth) * kDescVec;
                                __stcg(dsc + 0, make_uint4((uint32_t) alt.state, (uint32_t) (alt.state >> 32),
                                                           ds ? __float_as_uint(st) : 0u, __float_as_uint(ds ? swt : stmax)));
                                __stcg(dsc + 1, make_uint4(__float_as_uint(sox), __float_as_uint(soy), __float_as_uint(soz),
                                                           __float_as_uint(sdx)));
                                __stcg(dsc + 2, make_uint4(__float_as_uint(sdy), __float_as_uint(sdz), 0u, 0u));
                                __stcg(dsc + 3, make_uint4(0u, __float_as_uint(albedo[0]), __float_as_uint(albedo[1]),
                                                           __float_as_uint(albedo[2])));
                                // the four tap positions are drawn from the alt stream when the gradients are scattered
                                alt.next(); alt.next(); alt.next(); alt.next();
                                PU(F_ALT_LO, s) = (uint32_t) alt.state;
                                PU(F_ALT_HI, s) = (uint32_t) (alt.state >> 32);
                            }
                        }
                        // :193-200
                        if (ds) {
                            PSET(F_B0, s, beta[0] * albedo[0]);
                            PSET(F_B1, s, beta[1] * albedo[1]);
                            PSET(F_B2, s, beta[2] * albedo[2]);
                            depth += 1;
                        }
                        const bool active = ds && (depth < P.max_depth);
                        fl = active ? (fl | FL_ACTIVE) : (fl & ~FL_ACTIVE);
                        if (!ds) {
                            fl |= FL_ESCAPED;  // :244-245
                            next = Q_PATH_END;
                        } else {
                            fl = (P.use_nee && active) ? (fl & ~FL_SPAWN_PHASE) : (fl | FL_SPAWN_PHASE);
                            next = Q_SPAWN;
                        }
                        PU(F_DEPTH, s) = (PU(F_DEPTH, s) & 0xFFFF0000u) | (unsigned) depth;
                    }
                    PU(F_FLAGS, s) = fl;
                }
            } else if (work == Q_NEE_END) {
                // ---- sample_emitter_for_nee (:380-403) ----
                if (act) {
                    unsigned fl = PU(F_FLAGS, s);
                    const int pass = (int) (fl & FL_PASS_MASK);
                    const float Tn = PF(F_T, s);
                    float contrib[3];
                    if (ENV) {
                        contrib[0] = PF(F_NW0 + 0, s) * Tn;
                        contrib[1] = PF(F_NW0 + 1, s) * Tn;
                        contrib[2] = PF(F_NW0 + 2, s) * Tn;
                    } else {
                        contrib[0] = (PF(F_B0, s) * P.half_le[0]) * Tn;
                        contrib[1] = (PF(F_B1, s) * P.half_le[1]) * Tn;
                        contrib[2] = (PF(F_B2, s) * P.half_le[2]) * Tn;
                    }
                    next = Q_SPAWN;
                    if (HAS_DRT && pass == PP_DRTV) {
                        PSET(F_LI0, s, contrib[0]); PSET(F_LI1, s, contrib[1]); PSET(F_LI2, s, contrib[2]);
                    } else if (HAS_ADJ && pass == PP_ADJ) {
                        PSET(F_R0, s, PF(F_R0, s) + contrib[0]);  // gathered like the primal pass; :214 subtracts it from L
                        PSET(F_R1, s, PF(F_R1, s) + contrib[1]);  // instead, which Q_SCATTER does in the same order
                        PSET(F_R2, s, PF(F_R2, s) + contrib[2]);
                        {
                            const unsigned k = (PU(F_DEPTH, s) & 0xFFFFu) - 1u;  // the vertex this NEE belongs to
                            float* c = reinterpret_cast<float*>(P.desc + (((size_t) blockIdx.x * NSLOT + s) * P.desc_cap + k) * kDescVec) + 10;
                            __stcg(c + 0, contrib[0]); __stcg(c + 1, contrib[1]); __stcg(c + 2, contrib[2]);
                        }
                        if (fl & FL_NEE_VALID) {
                            const float a = (PF(F_DL0, s) * contrib[0] + PF(F_DL1, s) * contrib[1]) + PF(F_DL2, s) * contrib[2];
                            nee_n = PU(F_NLOG, s);
                            if (nee_n > (unsigned) kNeeLog) {
                                // more collisions than the log holds: walk the segment again (:393-401)
                                nee_n = 0u;
                                PSET(F_ASUM, s, a);
                                // the replay consumes exactly the same draws again
                                const float2 cl = __ldcg(P.neelog + ((size_t) blockIdx.x * NSLOT + s) * kNeeLogStride + kNeeLog);
                                PU(F_RNG_LO, s) = __float_as_uint(cl.x);
                                PU(F_RNG_HI, s) = __float_as_uint(cl.y);
                                fl = (fl & ~FL_MODE_MASK) | ((unsigned) PM_NEE_ADJ << FL_MODE_SHIFT);
                                next = Q_WALK;
                            } else {
                                nee_a = a;  // scattered from the log below; the sampler already stands behind the walk
                            }
                        }
                    } else {
                        PSET(F_R0, s, PF(F_R0, s) + contrib[0]);
                        PSET(F_R1, s, PF(F_R1, s) + contrib[1]);
                        PSET(F_R2, s, PF(F_R2, s) + contrib[2]);
                    }
                    if (next == Q_SPAWN) fl |= FL_SPAWN_PHASE;
                    PU(F_FLAGS, s) = fl;
                }
            } else if (work == Q_PATH_END) {
                bool want_rec = false;
                float wdl[3] = {0.0f, 0.0f, 0.0f};  // reservoir weight * dL: the adjoint handed to the DRT pass
                uint4 rs1 = make_uint4(0u, 0u, 0u, 0u), rs2 = rs1, rs4 = rs1;  // descriptor of the reservoir's vertex
                if (act) {
                    unsigned fl = PU(F_FLAGS, s);
                    const int pass = (int) (fl & FL_PASS_MASK);
                    const unsigned dw = PU(F_DEPTH, s);
                    const int depth = (int) (dw & 0xFFFFu);
                    float R[3] = {PF(F_R0, s), PF(F_R1, s), PF(F_R2, s)};
                    if (pass == PP_PRIMAL || (HAS_DRT && pass == PP_REC) || (HAS_ADJ && pass == PP_ADJ)) {
                        // :263-285 envmap
                        if ((fl & FL_ESCAPED) && !(depth <= 0 && P.hide_emitters)) {
                            if (ENV) {
                                // emitter.eval(si) with hit_mis_weight (:270-285); the slot still holds the last segment
                                float le[3], pdf;
                                env_eval(P, PF(F_DX, s), PF(F_DY, s), PF(F_DZ, s), le, pdf);
                                const bool hs = (fl & FL_HAS_SCATTERED) != 0u;
                                float wmis = 1.0f;
                                if (P.use_nee) wmis = mis_power(hs ? UIVR_INV_4PI : 1.0f, hs ? pdf : 0.0f);
#pragma unroll
                                for (int c = 0; c < 3; ++c) R[c] = R[c] + (PF(F_B0 + c, s) * wmis) * le[c];
                            } else {
                                const float wmis = (P.use_nee && (fl & FL_HAS_SCATTERED)) ? 0.5f : 1.0f;
#pragma unroll
                                for (int c = 0; c < 3; ++c) R[c] = fmaf(PF(F_B0 + c, s) * wmis, P.radiance[c], R[c]);
                            }
                        }
                    }
                    next = Q_FREE;
                    if (pass == PP_PRIMAL) {
                        const uint32_t idx = PU(F_IDX, s), pix = idx / P.spp;
                        if (P.sample_L) {
                            P.sample_L[3 * (size_t) idx + 0] = R[0];
                            P.sample_L[3 * (size_t) idx + 1] = R[1];
                            P.sample_L[3 * (size_t) idx + 2] = R[2];
                        }
                        if (P.image) {
                            atomicAdd(P.image + 3 * (size_t) pix + 0, R[0]);
                            atomicAdd(P.image + 3 * (size_t) pix + 1, R[1]);
                            atomicAdd(P.image + 3 * (size_t) pix + 2, R[2]);
                        }
                    } else if (HAS_ADJ && pass == PP_ADJ) {
                        // R is now the primal radiance L of this sample (state_in of sample(Backward), batched.py:309-318)
                        if (P.sample_L) {
                            const uint32_t idx = PU(F_IDX, s);
                            P.sample_L[3 * (size_t) idx + 0] = R[0];
                            P.sample_L[3 * (size_t) idx + 1] = R[1];
                            P.sample_L[3 * (size_t) idx + 2] = R[2];
                        }
                        if (use_rsv && (fl & FL_RS_VALID)) {
                            // DRTReservoir.get (:756-760) and adjoint = weight * dL (:255)
                            const uint4* dsc = P.desc + (((size_t) blockIdx.x * NSLOT + s) * P.desc_cap + (dw >> 16)) * kDescVec;
                            rs1 = __ldcg(dsc + 1); rs2 = __ldcg(dsc + 2); rs4 = __ldcg(dsc + 4);
                            const float wcur[3] = {__uint_as_float(rs4.x), __uint_as_float(rs4.y), __uint_as_float(rs4.z)};
                            const float wsum[3] = {PF(F_RSW0, s), PF(F_RSW1, s), PF(F_RSW2, s)};
                            const float d = mean3(wcur), ws = mean3(wsum);
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float W = (d != 0.0f) ? (ws * wcur[c]) / d : 0.0f;
                                wdl[c] = W * PF(F_DL0 + c, s);
                            }
                            want_rec = true;  // the DRT pass is a separate launch: hand the sample over through HBM
                        }
                        // deferred gradient scatter: one descriptor per segment end (every real collision, + the escape)
                        const unsigned n_desc = (unsigned) depth + ((fl & FL_ESCAPED) ? 1u : 0u);
                        if (n_desc) {
                            PSET(F_R0, s, R[0]); PSET(F_R1, s, R[1]); PSET(F_R2, s, R[2]);
                            PU(F_TS, s) = n_desc << 16;  // next vertex to scatter (low half) of n_desc (high half)
#if UIVR_POOL_SCATTER_AT_END
                            scatter_now = true;          // the first one in this very visit (below)
#else
                            next = Q_SCATTER;
#endif
                        }
                    } else if (HAS_DRT) {  // PP_REC: Li complete -> DRT gradient (:571-581)
                        const float dst = PF(F_DRT_ST, s), dD = PF(F_DRT_D, s), dt = PF(F_DRT_T, s);
                        const float m = P.use_drt_mis ? 1.0f / (1.0f + dst * dst) : 1.0f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float Li = PF(F_LI0 + c, s) + R[c];
                            const float term = ((m * dD) * PF(F_DL0 + c, s)) * Li;
                            sc_gs = fmaf(term, PF(F_AL0 + c, s), sc_gs);
                            sc_ga[c] = term * dst;
                        }
                        sc_vx = fmaf(dt, PF(F_RSDX, s), PF(F_RSOX, s));
                        sc_vy = fmaf(dt, PF(F_RSDY, s), PF(F_RSOY, s));
                        sc_vz = fmaf(dt, PF(F_RSDZ, s), PF(F_RSOZ, s));
                        sc_ff = true;
                    }
                    PU(F_FLAGS, s) = fl;
                }
                if (KIND == KIND_ADJ) {
                    // reservoir record for the DRT launch (warp-aggregated append)
                    const unsigned m = __ballot_sync(FULL, want_rec);
                    if (m) {
                        const int leader = __ffs(m) - 1;
                        unsigned base = 0;
                        if ((int) lane == leader) base = atomicAdd(P.rec_count, (unsigned) __popc(m));
                        base = __shfl_sync(FULL, base, leader);
                        if (want_rec) {
                            uint32_t* rec = P.records + (size_t) (base + __popc(m & lt_mask)) * kRecWords;
                            uint4* r4 = reinterpret_cast<uint4*>(rec);
                            r4[0] = rs1;  // o.xyz, d.x
                            r4[1] = make_uint4(rs2.x, rs2.y, rs4.w, __float_as_uint(wdl[0]));  // d.yz, t_exit
                            r4[2] = make_uint4(__float_as_uint(wdl[1]), __float_as_uint(wdl[2]), PU(F_ALT_LO, s), PU(F_ALT_HI, s));
                            r4[3] = make_uint4(PU(F_ASEQ, s), PU(F_DEPTH, s) >> 16, 0u, 0u);
                        }
                    }
                }
            } else if (work == Q_FREE) {
                // ---- Q_FREE: next work item from the global queue ----
                bool none_left = true;
                const int exh = __shfl_sync(FULL, *((volatile int*) &ctl->exhausted), 0);
                const unsigned fresh = __ballot_sync(FULL, act);
                uint64_t item = 0;
                bool mine = false;
                if (kPoolChunk == 0u && exh == 0 && fresh) {
                    const int leader = __ffs(fresh) - 1;
                    unsigned base = 0;
                    if ((int) lane == leader) base = atomicAdd(P.work_counter, (unsigned) __popc(fresh));
                    base = __shfl_sync(FULL, base, leader);
                    item = (uint64_t) base + __popc(fresh & lt_mask);
                    mine = act && item < total;
                    if (mine) none_left = false;
                    if ((uint64_t) base + __popc(fresh) >= total && (int) lane == leader) ctl->exhausted = 1;
                }
                if constexpr (kPoolChunk != 0u) if (exh == 0 && fresh) {
                    // The CTA reserves kPoolChunk items at a time (one global round trip per chunk instead of per
                    // batch) and its warps take from the reservation with a shared-memory compare-and-swap.
                    const int leader = __ffs(fresh) - 1;
                    const unsigned n = (unsigned) __popc(fresh);
                    unsigned base = 0, cnt = 0;
                    if ((int) lane == leader) {
                        for (;;) {
                            unsigned long long cs = *((volatile unsigned long long*) &ctl->chunk);
                            const unsigned left = (unsigned) (cs & 0xFFFFFFull);
                            if (left) {
                                const unsigned take = left < n ? left : n;
                                const unsigned long long ns = (((cs >> 24) + take) << 24) | (unsigned long long) (left - take);
                                if (atomicCAS(&ctl->chunk, cs, ns) == cs) { base = (unsigned) (cs >> 24); cnt = take; break; }
                                continue;
                            }
                            // empty: one warp reserves the next chunk; the others come back later (their slots stay free)
                            if (atomicCAS(&ctl->refill, 0, 1) != 0) break;
                            cs = *((volatile unsigned long long*) &ctl->chunk);
                            if (cs & 0xFFFFFFull) { atomicExch(&ctl->refill, 0); continue; }
                            const unsigned chunk = (unsigned) P.fetch_chunk;  // <= kPoolChunk: smaller for small launches (load balance)
                            const unsigned gb = atomicAdd(P.work_counter, chunk);
                            const uint64_t avail = (uint64_t) gb < total ? (total - gb < chunk ? total - gb : (uint64_t) chunk) : 0ull;
                            if (avail == 0ull) {
                                ctl->exhausted = 1;  // the global queue is empty: from now on free slots retire
                            } else {
                                cnt = avail < n ? (unsigned) avail : n;
                                base = gb;
                                if (avail > cnt)
                                    *((volatile unsigned long long*) &ctl->chunk) =
                                        ((unsigned long long) (gb + cnt) << 24) | (unsigned long long) (avail - cnt);
                            }
                            __threadfence_block();
                            atomicExch(&ctl->refill, 0);
                            break;
                        }
                    }
                    base = __shfl_sync(FULL, base, leader);
                    cnt = __shfl_sync(FULL, cnt, leader);
                    const unsigned r = (unsigned) __popc(fresh & lt_mask);
                    mine = act && r < cnt;
                    item = (uint64_t) base + r;
                    if (mine) none_left = false;
                    else if (act) next = Q_FREE;  // nothing reserved right now: the slot stays free
                }
                // no more work items: the slot retires
                const unsigned retire = __ballot_sync(FULL, act && none_left && (kPoolChunk == 0u || exh != 0));
                if (retire && lane == 0) atomicSub(&ctl->live, __popc(retire));
                if (KIND == KIND_DRT) {
                    // DRT launch: the work items are the reservoir records of the adjoint launch
                    if (mine) {
                        const uint4* r4 = reinterpret_cast<const uint4*>(P.records + (size_t) item * kRecWords);
                        const uint4 a = r4[0], b = r4[1], c = r4[2], d = r4[3];
                        // DRT on the stored segment (:543-581): the walk draws from the alt stream
                        PU(F_OX, s) = a.x; PU(F_OY, s) = a.y; PU(F_OZ, s) = a.z; PU(F_DX, s) = a.w;
                        PU(F_DY, s) = b.x; PU(F_DZ, s) = b.y; CU(C_TMAX, s) = b.z;
                        PU(F_RSOX, s) = a.x; PU(F_RSOY, s) = a.y; PU(F_RSOZ, s) = a.z; PU(F_RSDX, s) = a.w;
                        PU(F_RSDY, s) = b.x; PU(F_RSDZ, s) = b.y;
                        PU(F_DL0, s) = b.w; PU(F_DL1, s) = c.x; PU(F_DL2, s) = c.y;
                        PU(F_RNG_LO, s) = c.z; PU(F_RNG_HI, s) = c.w; PU(F_SEQ, s) = d.x;
                        PU(F_DEPTH, s) = d.y;
                        PU(F_FLAGS, s) = (unsigned) PP_ADJ | ((unsigned) PM_DRT << FL_MODE_SHIFT);
                        next = Q_WALK;
                    }
                } else {
                    // ray generation + reach_medium (batched.py:426-467, volpathsimple.py:292-319)
                    uint32_t idx = 0, pix = 0;
                    bool have = false;
                    if (mine) {
                        const uint32_t it = (uint32_t) item;
                        if (slot_to_pixel(P, it / P.spp, pix)) {
                            idx = pix * P.spp + it % P.spp;
                            have = true;
                            K.add(C_SAMPLES, 1);
                        } else {
                            next = Q_FREE;  // padding slot of a shard: try again
                        }
                    }
                    if (have) {
                        const int pass = KIND == KIND_ADJ ? PP_ADJ : PP_PRIMAL;
                        Rng r;
                        r.state = r.inc = 0;
#pragma unroll 1
                        for (int k = HAS_ADJ ? 1 : 0; k >= 0; --k) {  // sampler.seed(seed, wavefront) [+ the alt sampler, :100-107]
                            pool_seed_sampler(r, k ? P.alt_seed : P.seed, idx);
                            if (k) {
                                PU(F_ALT_LO, s) = (uint32_t) r.state; PU(F_ALT_HI, s) = (uint32_t) (r.state >> 32);
                                PU(F_ASEQ, s) = (uint32_t) (r.inc >> 1);
                            }
                        }
                        if (HAS_ADJ) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                PSET(F_DL0 + c, s, ldg_tap(P.grad_image + 3 * (size_t) pix + c) * P.inv_spp);
                                PSET(F_RSW0 + c, s, 0.0f);
                                PSET(F_R0 + c, s, 0.0f);  // the replay gathers the primal radiance itself
                            }
                        } else {
                            PSET(F_R0, s, 0.0f); PSET(F_R1, s, 0.0f); PSET(F_R2, s, 0.0f);
                        }
                        Seg sg;
                        float F[15], fu, fv;
                        if (BATCH) {
                            // ray-batch mode: "pix" is the batch element; the path sampler draws no jitter (batched.py:390)
                            batch_film_position(P, pix, idx, F, fu, fv);
                        } else {
                            const float jx = draw(r, K), jy = draw(r, K);
                            sensor_film_position(P, pix, jx, jy, F, fu, fv);
                        }
                        const int status = camera_segment_frame(P, F, fu, fv, sg);
                        draw(r, K);  // :71
                        const bool active = status == 1;
                        const bool escaped = status == 0;
                        unsigned fl = (unsigned) pass | ((unsigned) PM_DELTA << FL_MODE_SHIFT) | (escaped ? FL_ESCAPED : 0u) |
                                      (active ? FL_ACTIVE : 0u);
                        if (active) {
                            draw(r, K);  // :99 alt_seed_rnd
                            K.add(C_HITS, 1);
                            draw(r, K);  // :120 Russian-roulette draw of the first loop iteration
                            PU(F_IDX, s) = idx;
                            PU(F_RNG_LO, s) = (uint32_t) r.state; PU(F_RNG_HI, s) = (uint32_t) (r.state >> 32);
                            PU(F_SEQ, s) = (uint32_t) (r.inc >> 1);
                            PSET(F_OX, s, sg.ox); PSET(F_OY, s, sg.oy); PSET(F_OZ, s, sg.oz);
                            PSET(F_DX, s, sg.dx); PSET(F_DY, s, sg.dy); PSET(F_DZ, s, sg.dz);
                            CSET(C_TMAX, s, sg.tmax);
                            PSET(F_B0, s, 1.0f); PSET(F_B1, s, 1.0f); PSET(F_B2, s, 1.0f);
                            PU(F_DEPTH, s) = 0u;
                            next = Q_WALK;
                        } else {
                            {
                                // the ray misses the medium: finish the sample right here (no gradient)
                                float R[3] = {0.0f, 0.0f, 0.0f};
                                if (escaped && !P.hide_emitters) {
                                    if (ENV) {
                                        float pdf;
                                        env_eval(P, sg.dx, sg.dy, sg.dz, R, pdf);  // no scattering: MIS weight 1
                                    } else {
#pragma unroll
                                        for (int c = 0; c < 3; ++c) R[c] = fmaf(1.0f, P.radiance[c], 0.0f);
                                    }
                                }
                                if (P.sample_L) {
                                    P.sample_L[3 * (size_t) idx + 0] = R[0];
                                    P.sample_L[3 * (size_t) idx + 1] = R[1];
                                    P.sample_L[3 * (size_t) idx + 2] = R[2];
                                }
                                if (P.image) {
                                    atomicAdd(P.image + 3 * (size_t) pix + 0, R[0]);
                                    atomicAdd(P.image + 3 * (size_t) pix + 1, R[1]);
                                    atomicAdd(P.image + 3 * (size_t) pix + 2, R[2]);
                                }
                            }
                            fl = 0u;
                            next = Q_FREE;
                        }
                        PU(F_FLAGS, s) = fl;
                    }
                }
            }

            // ---- deferred gradients of one described vertex of a finished path (see the vertex handler): for the
            //      batch of Q_SCATTER and, without a queue hop, for the first vertex of the paths that have just ended ----
            if (HAS_ADJ && __ballot_sync(FULL, act && (work == Q_SCATTER || scatter_now))) {
                if (act && (work == Q_SCATTER || scatter_now)) {
                    // vertices in path order: F_R holds L minus the NEE contributions of the vertices before this one,
                    // subtracted one by one exactly like the reference's replay does (:214)
                    const unsigned tw = PU(F_TS, s), k = tw & 0xFFFFu, n_desc = tw >> 16;
                    PU(F_TS, s) = tw + 1u;
                    const uint4* dsc = P.desc + (((size_t) blockIdx.x * NSLOT + s) * P.desc_cap + k) * kDescVec;
                    const uint4 d0 = __ldcg(dsc + 0), d1 = __ldcg(dsc + 1), d2 = __ldcg(dsc + 2), d3 = __ldcg(dsc + 3);
                    alt.state = (uint64_t) d0.x | ((uint64_t) d0.y << 32);
                    alt.inc = ((uint64_t) PU(F_ASEQ, s) << 1) | 1ull;
                    const float st = __uint_as_float(d0.z);
                    const bool ds = d0.z != 0u;
                    sc_int = __uint_as_float(d0.w);
                    sc_ox = __uint_as_float(d1.x); sc_oy = __uint_as_float(d1.y); sc_oz = __uint_as_float(d1.z);
                    sc_dx = __uint_as_float(d1.w); sc_dy = __uint_as_float(d2.x); sc_dz = __uint_as_float(d2.y);
                    const float dL[3] = {PF(F_DL0, s), PF(F_DL1, s), PF(F_DL2, s)};
                    const float R[3] = {PF(F_R0, s), PF(F_R1, s), PF(F_R2, s)};
                    PSET(F_R0, s, R[0] - __uint_as_float(d2.z));
                    PSET(F_R1, s, R[1] - __uint_as_float(d2.w));
                    PSET(F_R2, s, R[2] - __uint_as_float(d3.x));
                    const float albedo[3] = {__uint_as_float(d3.y), __uint_as_float(d3.z), __uint_as_float(d3.w)};
                    next = k + 1u < n_desc ? Q_SCATTER : Q_FREE;
                    // :152-172 free-flight scattering gradient
                    if ((!P.use_drt || P.use_drt_mis) && ds) {
                        float m = 1.0f;
                        if (P.use_drt && P.use_drt_mis) {
                            const float s2 = st * st;
                            m = s2 / (1.0f + s2);
                        }
                        const float inv_pdf = 1.0f / st;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float Li = R[c] / (albedo[c] > 1e-8f ? albedo[c] : 1e-8f);
                            const float term = ((m * dL[c]) * Li) * inv_pdf;
                            sc_gs = fmaf(term, albedo[c], sc_gs);
                            sc_ga[c] = term * st;
                        }
                        sc_ff = true;
                        sc_vx = fmaf(sc_int, sc_dx, sc_ox); sc_vy = fmaf(sc_int, sc_dy, sc_oy); sc_vz = fmaf(sc_int, sc_dz, sc_oz);
                    }
                    // :181-189, :584-607 transmittance gradient: 4 uniform taps on the segment
                    const float aw = fmaf(dL[2], R[2], fmaf(dL[1], R[1], dL[0] * R[0]));
                    sc_g = -(aw * (sc_int * 0.25f));
                    sc_taps = true;
                }
            }

            // ---- NEE adjoint from the collision log: d sigma_t(p_i) += -sum(adjoint) / sigma_n(p_i)  (:483-492) ----
            if (HAS_ADJ && work == Q_NEE_END) {
                unsigned n_max = nee_n;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(FULL, n_max, o));
                if (n_max) {
                    const float2* lg = P.neelog + ((size_t) blockIdx.x * NSLOT + s) * kNeeLogStride;
                    const float ox = nee_n ? PF(F_OX, s) : 0.0f, oy = nee_n ? PF(F_OY, s) : 0.0f, oz = nee_n ? PF(F_OZ, s) : 0.0f;
                    const float dx = nee_n ? PF(F_DX, s) : 0.0f, dy = nee_n ? PF(F_DY, s) : 0.0f, dz = nee_n ? PF(F_DZ, s) : 0.0f;
#pragma unroll 1
                    for (unsigned i = 0; i < n_max; ++i) {
                        if (i < nee_n) {
                            const float2 e = __ldcg(lg + i);
                            scatter_sigma<kPoolAcc>(P, fmaf(e.x, dx, ox), fmaf(e.x, dy, oy), fmaf(e.x, dz, oz), -nee_a / e.y);
                            K.add(C_SSCAT, 1);
                        }
                    }
                }
            }

            // ---- emitter sampling (:406-433) / phase sampling (:221-245, :626-652), one code site: for the batch of
            //      Q_SPAWN itself (end of a NEE-adjoint replay walk) and, without a queue hop in between, for the slots a
            //      vertex or NEE-end handler has just sent on to it ----
            if (__ballot_sync(FULL, act && (work == Q_SPAWN || (UIVR_POOL_INLINE_SPAWN && next == Q_SPAWN)))) {
                if (act && (work == Q_SPAWN || (UIVR_POOL_INLINE_SPAWN && next == Q_SPAWN))) {
                    unsigned fl = PU(F_FLAGS, s);
                    const bool phase = (fl & FL_SPAWN_PHASE) != 0u;
                    Rng r;
                    r.state = (uint64_t) PU(F_RNG_LO, s) | ((uint64_t) PU(F_RNG_HI, s) << 32);
                    r.inc = ((uint64_t) PU(F_SEQ, s) << 1) | 1ull;
                    if (phase) draw(r, K);
                    const float xi1 = draw(r, K), xi2 = draw(r, K);
                    float wx, wy, wz;
                    bool worked = true;
                    if (ENV && !phase) {
                        // scene.sample_emitter_direction on the envmap; the NEE weight waits in the slot (:385-391, :419-423)
                        float pdf, le[3];
                        env_sample(P, xi1, xi2, wx, wy, wz, pdf, le);
                        worked = pdf != 0.0f;
                        const float mis = mis_power(pdf, UIVR_INV_4PI);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            PSET(F_NW0 + c, s, worked ? ((PF(F_B0 + c, s) * UIVR_INV_4PI) * mis) * (le[c] / pdf) : 0.0f);
                    } else {
                        uniform_sphere(xi1, xi2, wx, wy, wz);
                    }
                    Seg sg;
                    const bool ok = make_segment(P, PF(F_OX, s), PF(F_OY, s), PF(F_OZ, s), wx, wy, wz, sg) && worked;
                    PSET(F_OX, s, sg.ox); PSET(F_OY, s, sg.oy); PSET(F_OZ, s, sg.oz);
                    PSET(F_DX, s, sg.dx); PSET(F_DY, s, sg.dy); PSET(F_DZ, s, sg.dz);
                    CSET(C_TMAX, s, sg.tmax);
                    bool active = (fl & FL_ACTIVE) != 0u;
                    bool rr = false;  // Russian-roulette draw + zero-throughput test of the next loop iteration
                    if (!phase) {
                        if (HAS_ADJ && (fl & FL_PASS_MASK) == (unsigned) PP_ADJ) {
                            // sampler.clone() position for the adjoint replay (:383)
                            __stcg(P.neelog + ((size_t) blockIdx.x * NSLOT + s) * kNeeLogStride + kNeeLog,
                                   make_float2(__uint_as_float((uint32_t) r.state), __uint_as_float((uint32_t) (r.state >> 32))));
                            PU(F_NLOG, s) = 0u;
                        }
                        fl = ok ? (fl | FL_NEE_VALID) : (fl & ~FL_NEE_VALID);
                        fl = (fl & ~FL_MODE_MASK) | ((unsigned) PM_NEE << FL_MODE_SHIFT);
                        PSET(F_T, s, ok ? 1.0f : 0.0f);
                        next = ok ? Q_WALK : Q_NEE_END;
                    } else {
                        fl = (fl & ~FL_MODE_MASK) | ((unsigned) PM_DELTA << FL_MODE_SHIFT);
                        if (HAS_DRT && (fl & FL_PASS_MASK) == (unsigned) PP_DRTV) {
                            const unsigned dw = PU(F_DEPTH, s);
                            const int depth = (int) (dw & 0xFFFFu) + 1;
                            PU(F_DEPTH, s) = (dw & 0xFFFF0000u) | (unsigned) depth;
                            active = ok && (depth < P.max_depth);
                            fl = (fl & ~FL_PASS_MASK) | (unsigned) PP_REC;
                            PSET(F_B0, s, 1.0f); PSET(F_B1, s, 1.0f); PSET(F_B2, s, 1.0f);
                            PSET(F_R0, s, 0.0f); PSET(F_R1, s, 0.0f); PSET(F_R2, s, 0.0f);
                            fl = (fl & ~FL_ESCAPED) | FL_HAS_SCATTERED;
                            if (active) draw(r, K);  // :99 of the recursive sample()
                        } else if (!ok) {
                            active = false;  // :240-241 accidental escape
                        }
                        rr = active;
                    }
                    if (rr) {
                        draw(r, K);  // :120
                        if (PF(F_B0, s) == 0.0f && PF(F_B1, s) == 0.0f && PF(F_B2, s) == 0.0f) active = false;  // :121
                    }
                    if (phase) {
                        fl = active ? (fl | FL_ACTIVE) : (fl & ~FL_ACTIVE);
                        next = active ? Q_WALK : Q_PATH_END;
                    }
                    PU(F_RNG_LO, s) = (uint32_t) r.state;
                    PU(F_RNG_HI, s) = (uint32_t) (r.state >> 32);
                    PU(F_FLAGS, s) = fl;
                }
            }

            // ---- gradient scatter of the batch (one code site): 4 transmittance taps, then the
            //      collision / DRT vertex itself ----
            if (BWD && __ballot_sync(FULL, sc_taps || sc_ff)) {
#pragma unroll 1
                for (int k = 0; k < 5; ++k) {
                    const bool on = k < 4 ? sc_taps : sc_ff;
                    if (on) {
                        float px = sc_vx, py = sc_vy, pz = sc_vz, g = sc_gs;
                        if (k < 4) {
                            const float tk = draw(alt, K) * sc_int;
                            px = fmaf(tk, sc_dx, sc_ox); py = fmaf(tk, sc_dy, sc_oy); pz = fmaf(tk, sc_dz, sc_oz);
                            g = sc_g;
                        }
                        scatter_sigma<kPoolAcc>(P, px, py, pz, g);
                        K.add(C_SSCAT, 1);
                        if (k == 4 && sc_alb) {
                            scatter_albedo<kPoolAcc>(P, px, py, pz, sc_ga);
                            K.add(C_ASCAT, 1);
                        }
                    }
                }
            }

            // ---- set-up of a new free-flight walk (one code site; Medium::sample_interaction set-up, App. B.5):
            //      first cell, next-boundary times and their increments, the first optical depth tau ----
            if (__ballot_sync(FULL, next == Q_WALK && !resume)) {
                if (next == Q_WALK && !resume) {
                    const float ox = PF(F_OX, s), oy = PF(F_OY, s), oz = PF(F_OZ, s);
                    const float dx = PF(F_DX, s), dy = PF(F_DY, s), dz = PF(F_DZ, s);
                    const float ix = dx != 0.0f ? 1.0f / dx : UIVR_INF;
                    const float iy = dy != 0.0f ? 1.0f / dy : UIVR_INF;
                    const float iz = dz != 0.0f ? 1.0f / dz : UIVR_INF;
                    int cx, cy, cz;
                    float tnx, tny, tnz;
                    walk_axis_init(ox, dx, ix, P.fmres[0], P.mcs[0], P.mres[0], cx, tnx);
                    walk_axis_init(oy, dy, iy, P.fmres[1], P.mcs[1], P.mres[1], cy, tny);
                    walk_axis_init(oz, dz, iz, P.fmres[2], P.mcs[2], P.mres[2], cz, tnz);
                    const unsigned oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
                    const unsigned ci0 = (unsigned) (((cz + 1) * P.pm[1] + (cy + 1)) * P.pm[0] + (cx + 1));
                    K.add(C_MAJ, 1);  // the first cell (read by the walker)
                    Rng r;
                    r.state = (uint64_t) PU(F_RNG_LO, s) | ((uint64_t) PU(F_RNG_HI, s) << 32);
                    r.inc = ((uint64_t) PU(F_SEQ, s) << 1) | 1ull;
                    const float tau0 = neg_log1m(draw(r, K));
                    PU(F_RNG_LO, s) = (uint32_t) r.state;
                    PU(F_RNG_HI, s) = (uint32_t) (r.state >> 32);
                    // per-mode walk state, and the queue the slot goes to when the walk ends
                    unsigned fl = PU(F_FLAGS, s);
                    const int mode = (int) ((fl & FL_MODE_MASK) >> FL_MODE_SHIFT);
                    int endq = Q_VERTEX;
                    if (mode == PM_DELTA) {
                        // vertices of the adjoint replay scatter gradients: they get their own queue so that
                        // the scatter loop of a batch runs with all lanes
                        if (HAS_ADJ && (fl & FL_PASS_MASK) == (unsigned) PP_ADJ) endq = Q_VERTEX_ADJ;
                    } else if (mode == PM_NEE) {
                        endq = Q_NEE_END;
                    } else if (HAS_ADJ && mode == PM_NEE_ADJ) {
                        endq = Q_SPAWN;
                        fl |= FL_SPAWN_PHASE;
                        PSET(F_T2, s, 1.0f);
                    } else if (HAS_DRT && mode == PM_DRT) {
                        PSET(F_T, s, 1.0f);
                        PSET(F_DRT_D, s, 0.0f);
                    }
                    fl = (fl & ~(FL_ENDQ_MASK | FL_DID_SCATTER | FL_DRT_FOUND)) | ((unsigned) endq << FL_ENDQ_SHIFT);
                    PU(F_FLAGS, s) = fl;
                    // the continuation record (C_TMAX was written with the segment): the first step is decided here
                    const int ppx = P.pm[0], ppxy = P.pm[0] * P.pm[1];
                    const float adx = fabsf(P.mcs[0] * ix), ady = fabsf(P.mcs[1] * iy), adz = fabsf(P.mcs[2] * iz);
                    float tcur;
                    int cin;
                    walk_decide(tnx, tny, tnz, adx, ady, adz, (oct & 1u) ? -1 : 1, (oct & 2u) ? -ppx : ppx,
                                (oct & 4u) ? -ppxy : ppxy, (int) ci0, tcur, cin);
                    uint32_t* rec = cont + s * C_WORDS;
                    *reinterpret_cast<uint4*>(rec) = make_uint4(__float_as_uint(tnx), __float_as_uint(tny), __float_as_uint(tnz),
                                                                __float_as_uint(tau0));
                    rec[C_ADX] = __float_as_uint(adx);
                    rec[C_ADY] = __float_as_uint(ady);
                    rec[C_ADZ] = __float_as_uint(adz);
                    *reinterpret_cast<uint4*>(rec + C_CI) = make_uint4(ci0 | (oct << 28), 0u /* t = 0 */,
                                                                       (unsigned) cin | ((unsigned) endq << 28),
                                                                       __float_as_uint(tcur));
                }
            }
            route(s, next);
        }
    }
#undef WT
#undef PU
#undef PF
#undef PSET
#undef CU
#undef CF
#undef CSET
    // (after an abort) the thread that recorded the reason dumps the queue state of its CTA
    if (*((volatile int*) &ctl->abort) && P.debug && P.debug[1] == blockIdx.x && P.debug[2] == threadIdx.x && P.debug[0] != 0u)
        pool_trip_dump(ctl, P.debug);
    K.flush(P.counters);
}

// NSLOT: in-flight samples per CTA (one CTA per SM).  Per slot: 19 (forward) / 29 (adjoint) / 37 (DRT) pool words + the
// 12-word continuation record + 8 16-bit ring cells.  The pools are sized to the 196 KB step of the shared-memory
// carve-out (199 680 bytes usable): one step more leaves 28 instead of 60 KB of L1 for the walk table and the taps.
// kind: KIND_FWD / KIND_ADJ / KIND_DRT
inline int launch_pool(int num_sms, int kind, bool counting, const Params& P, cudaStream_t st) {
    cudaError_t e;
    const bool env = P.env_data != nullptr;
    const bool batch = P.sensors != nullptr;
#define UIVR_POOL_LAUNCH(KD, C, N, T, H, E, B)                                                          \
    do {                                                                                                \
        const size_t smem = pool_smem_bytes<KD, N, E>() +                                               \
                            (UIVR_POOL_SMEMTAB ? (size_t) P.wtab_words * 4 : 0);                        \
        if (smem > 232448) return -4; /* the walk table does not fit next to the pool */                \
        e = cudaFuncSetAttribute(k_pool<KD, C, N, T, H, E, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        if (e != cudaSuccess) return -2;                                                                \
        if (UIVR_POOL_SETMAXNREG) {                                                                     \
            /* setmaxnreg.inc waits for registers of the CTA's pool: the pool must be what the budget assumes */ \
            cudaFuncAttributes fa;                                                                      \
            if (cudaFuncGetAttributes(&fa, k_pool<KD, C, N, T, H, E, B>) != cudaSuccess ||              \
                fa.numRegs != (65536 / (T)) / 8 * 8) return -3;                                         \
        }                                                                                               \
        k_pool<KD, C, N, T, H, E, B><<<num_sms, T, smem, st>>>(P);                                      \
    } while (0)
#define UIVR_POOL_LAUNCH3(KD, C, N, T, H, E)                                                            \
    do {                                                                                                \
        if (batch) UIVR_POOL_LAUNCH(KD, C, N, T, H, E, true); else UIVR_POOL_LAUNCH(KD, C, N, T, H, E, false); \
    } while (0)
#define UIVR_POOL_LAUNCH2(KD, N, T, H)                                                                  \
    do {                                                                                                \
        if (env) {                                                                                      \
            constexpr int NE = pool_env_slots<KD, N>();                                                 \
            if (counting) UIVR_POOL_LAUNCH3(KD, true, NE, T, H, true); else UIVR_POOL_LAUNCH3(KD, false, NE, T, H, true); \
        } else {                                                                                        \
            if (counting) UIVR_POOL_LAUNCH3(KD, true, N, T, H, false); else UIVR_POOL_LAUNCH3(KD, false, N, T, H, false); \
        }                                                                                               \
    } while (0)
    switch (kind) {
        case KIND_FWD: UIVR_POOL_LAUNCH2(KIND_FWD, UIVR_POOL_SLOTS_FWD, UIVR_POOL_BLOCK_FWD, UIVR_POOL_HANDLERS_FWD); break;
        case KIND_ADJ: UIVR_POOL_LAUNCH2(KIND_ADJ, UIVR_POOL_SLOTS_ADJ, UIVR_POOL_BLOCK_ADJ, UIVR_POOL_HANDLERS_ADJ); break;
        case KIND_DRT: UIVR_POOL_LAUNCH2(KIND_DRT, UIVR_POOL_SLOTS_DRT, UIVR_POOL_BLOCK_DRT, UIVR_POOL_HANDLERS_DRT); break;
        default: return -1;
    }
#undef UIVR_POOL_LAUNCH2
#undef UIVR_POOL_LAUNCH3
#undef UIVR_POOL_LAUNCH
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace uivr
