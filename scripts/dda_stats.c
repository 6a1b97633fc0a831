/* dda_stats.c -- analysis build of the CPU oracle: counts the supergrid DDA iterations of the free-flight
 * walks by kind, to size empty-space optimisations of the CUDA walker BEFORE spending GPU time:
 *   - iterations in empty cells (majorant 0),
 *   - iterations an "exit mask" would cut (everything ahead of the cell, in the ray's octant, is empty),
 *   - iterations octant-cube / Chebyshev-distance jumps would replace, and the number of jumps.
 * Build + run: python scripts/dda_stats.py.  Test/analysis infrastructure, not product code. */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct {
    unsigned long long walks, iters, empty, exit_cut, oct_free, oct_jumps, iso_free, iso_jumps, first_iter_exit;
    unsigned long long hist_len[65];
} dda_stats_t;

static dda_stats_t g_stats;
static const uint8_t* g_exit;   /* per cell: bit o set = the octant box from this cell to the grid border is empty */
static const uint8_t* g_oct;    /* per cell x 8 octants: side of the largest empty cube anchored here (capped) */
static const uint8_t* g_iso;    /* per cell: Chebyshev distance to the nearest non-empty cell (capped) */

typedef struct {
    int oct, cut, n;
    int oc[3], od, ic[3], id;   /* running octant / iso jump: origin cell and distance (0 = none) */
    dda_stats_t s;
} walk_stats_t;
static __thread walk_stats_t tw;

#define UIVR_ORACLE_WALK_HOOK(C, s, w) stats_walk_begin((s)->d)
#define UIVR_ORACLE_STEP_HOOK(C, w) stats_step((C)->mres, (w)->cell, (w)->sig_bar)

static void stats_flush(void) {
    unsigned long long* g = (unsigned long long*) &g_stats;
    unsigned long long* l = (unsigned long long*) &tw.s;
    for (size_t i = 0; i < sizeof(dda_stats_t) / sizeof(unsigned long long); ++i)
        if (l[i]) __atomic_fetch_add(&g[i], l[i], __ATOMIC_RELAXED);
    memset(&tw.s, 0, sizeof(tw.s));
}

static inline void stats_walk_begin(const float d[3]) {
    if (tw.n) tw.s.hist_len[tw.n > 64 ? 64 : tw.n]++;
    if (tw.s.walks >= 4096) stats_flush();
    tw.s.walks++;
    tw.oct = (d[0] < 0.0f ? 1 : 0) | (d[1] < 0.0f ? 2 : 0) | (d[2] < 0.0f ? 4 : 0);
    tw.cut = 0; tw.n = 0; tw.od = 0; tw.id = 0;
}

static inline int cheb(const int a[3], const int b[3]) {
    int m = 0;
    for (int k = 0; k < 3; ++k) { int v = abs(a[k] - b[k]); if (v > m) m = v; }
    return m;
}

static inline void stats_step(const int32_t mres[3], const int cell[3], float sb) {
    const size_t ci = ((size_t) cell[2] * mres[1] + cell[1]) * mres[0] + cell[0];
    tw.s.iters++;
    tw.n++;
    if (sb > 0.0f) { tw.od = tw.id = 0; return; }
    tw.s.empty++;
    if (tw.cut || ((g_exit[ci] >> tw.oct) & 1)) {
        if (!tw.cut && tw.n == 1) tw.s.first_iter_exit++;
        if (tw.cut) tw.s.exit_cut++;   /* the iteration that sees the flag is still executed */
        tw.cut = 1;
        return;
    }
    /* octant-cube jumps */
    if (tw.od && cheb(cell, tw.oc) < tw.od) tw.s.oct_free++;
    else {
        const int d = g_oct[ci * 8 + tw.oct];
        if (d >= 2) { tw.od = d; memcpy(tw.oc, cell, sizeof(tw.oc)); tw.s.oct_jumps++; } else tw.od = 0;
    }
    if (tw.id && cheb(cell, tw.ic) < tw.id) tw.s.iso_free++;
    else {
        const int d = g_iso[ci];
        if (d >= 2) { tw.id = d; memcpy(tw.ic, cell, sizeof(tw.ic)); tw.s.iso_jumps++; } else tw.id = 0;
    }
}

#include "../oracle/uivr_oracle.c"

/* tables from the majorant grid */
static uint8_t *t_exit, *t_oct, *t_iso;

void dda_stats_prepare(const float* sigma_t, const int32_t res[3], float scale, int32_t factor, int cap) {
    int32_t m[3];
    for (int a = 0; a < 3; ++a) { m[a] = factor > 1 ? res[a] / factor : 1; if (m[a] < 1) m[a] = 1; }
    const size_t n = (size_t) m[0] * m[1] * m[2];
    float* maj = (float*) malloc(n * sizeof(float));
    uivr_oracle_build_majorant(sigma_t, res, scale, factor, m, maj);
    free(t_exit); free(t_oct); free(t_iso);
    t_exit = (uint8_t*) calloc(n, 1); t_oct = (uint8_t*) calloc(n * 8, 1); t_iso = (uint8_t*) calloc(n, 1);
#define CI(x, y, z) (((size_t) (z) * m[1] + (y)) * m[0] + (x))
    for (int o = 0; o < 8; ++o) {
        const int sx = (o & 1) ? -1 : 1, sy = (o & 2) ? -1 : 1, sz = (o & 4) ? -1 : 1;
        /* sweep against the octant direction so that the cells ahead are done first */
        for (int kz = 0; kz < m[2]; ++kz) for (int ky = 0; ky < m[1]; ++ky) for (int kx = 0; kx < m[0]; ++kx) {
            const int x = sx > 0 ? m[0] - 1 - kx : kx, y = sy > 0 ? m[1] - 1 - ky : ky, z = sz > 0 ? m[2] - 1 - kz : kz;
            const int empty = !(maj[CI(x, y, z)] > 0.0f);
            int ex = empty, cube = empty ? cap : 0;
            for (int k = 1; k < 8 && empty; ++k) {
                const int nx = x + ((k & 1) ? sx : 0), ny = y + ((k & 2) ? sy : 0), nz = z + ((k & 4) ? sz : 0);
                if (nx < 0 || ny < 0 || nz < 0 || nx >= m[0] || ny >= m[1] || nz >= m[2]) continue;  /* outside = empty */
                ex &= (t_exit[CI(nx, ny, nz)] >> o) & 1;
                const int c = t_oct[CI(nx, ny, nz) * 8 + o] + 1;
                if (c < cube) cube = c;
            }
            if (ex) t_exit[CI(x, y, z)] |= (uint8_t) (1 << o);
            t_oct[CI(x, y, z) * 8 + o] = (uint8_t) cube;
        }
    }
    for (int z = 0; z < m[2]; ++z) for (int y = 0; y < m[1]; ++y) for (int x = 0; x < m[0]; ++x) {
        int d = 0;
        if (!(maj[CI(x, y, z)] > 0.0f)) {
            for (d = 1; d < cap; ++d) {
                int hit = 0;
                for (int dz = -d; dz <= d && !hit; ++dz) for (int dy = -d; dy <= d && !hit; ++dy) for (int dx = -d; dx <= d && !hit; ++dx) {
                    if (abs(dx) != d && abs(dy) != d && abs(dz) != d) continue;
                    const int nx = x + dx, ny = y + dy, nz = z + dz;
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= m[0] || ny >= m[1] || nz >= m[2]) continue;
                    if (maj[CI(nx, ny, nz)] > 0.0f) hit = 1;
                }
                if (hit) break;
            }
        }
        t_iso[CI(x, y, z)] = (uint8_t) d;
    }
    g_exit = t_exit; g_oct = t_oct; g_iso = t_iso;
    free(maj);
    memset(&g_stats, 0, sizeof(g_stats));
}

void dda_stats_get(unsigned long long* out) {
    /* worker threads have exited; their thread-local tails were flushed by dda_stats_thread_done */
    memcpy(out, &g_stats, sizeof(g_stats));
}

/* called through the public entry points: flush the calling thread's tail */
void dda_stats_flush_thread(void) { stats_flush(); }
