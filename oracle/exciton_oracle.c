/*
 * exciton_oracle.c — CPU statement of the exciton-trajectory model ("Path X").
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (only tests/ and bench.py's CPU-baseline legs load it).
 *
 * PARITY UNPINNED.  The reference contains no exciton dynamics: only the event codes
 * (event.py:43-51: ISC = 3, Forster = 4, decay = 5), empty event lists (reseau.py:233-235), the
 * `...` stubs of the dispatcher (reseau.py:526-550), TADF.intersystem_crossing (molecule.py:299-307),
 * the 25 % / 75 % spin statistics (molecule.py:93), the per-site S1 / T1 energies
 * (molecule.py:182-183) and the intended routing comment (reseau.py:533).  Everything else below —
 * the rate laws, the parameters, rejection-free (BKL) selection — is this repository's own model,
 * declared in DESIGN.md §9; this file is the specification the CUDA kernel is tested against, not a
 * restatement of reference code.
 *
 * Model.  One exciton per trajectory on a lattice realisation, described by two per-site arrays
 * of TAGGED Boltzmann weights: ws1 = exp(-S1/kT), wt1 = exp(-T1/kT) whose two least-significant
 * mantissa bits are replaced by the site's species code (0 host, 1 TADF, 2 fluorescent).  The tag
 * perturbs a weight by at most 3 ulp (an energy shift of ~1e-17 eV) and lets one 8-byte load
 * deliver both the acceptor's weight and its species; precomputing the weights keeps every
 * transcendental off the selection path.  Channels of an exciton on site i (species a, spin s):
 *   m = 0..M-1   transfer to neighbour j = i + offset[m] (periodic x, y; open z; j invalid -> 0)
 *                singlet: Forster  k = (kF[a][b] * g6[m]) * min(1, ws1[j] / ws1[i])
 *                triplet: Dexter   k = (kD[a][b] * gD[m]) * min(1, wt1[j] / wt1[i])
 *   M            spin flip: singlet ISC k_isc[a]; triplet RISC k_risc[a] * min(1, ws1[i] / wt1[i])
 *   M+1          radiative decay      k_rad[a][s]
 *   M+2          non-radiative decay  k_nonrad[a][s]
 * Rejection-free selection: inclusive prefix sums P_c in the canonical order below, total R,
 * (canonical order below); waiting time dt = -log(1 - u_time) / R.  (u_sel, u_time) are the two halves of Philox
 * block `draws / 2` of stream 4 (XTRAJ) of the trajectory.  A decayed exciton is replaced at once
 * by a new one: site = interior site floor(u_a * n_interior), spin singlet iff u_b < 0.25.
 *
 * Canonical summation order (what makes the selection reproducible bit for bit on a warp):
 * channel c belongs to lane l = c % 32, slot k = c / 32.  S_l = rate(0,l) + rate(1,l) + ... summed
 * in slot order from 0.0; L = inclusive Kogge-Stone scan of S over the 32 lanes (offsets 1, 2, 4,
 * 8, 16, every stage reading the previous one); R = L_31; target = u_sel * R.  The chosen lane is
 * the first with L_l > target (if rounding leaves none: the last lane with S_l > 0); inside it the
 * prefix restarts from L_{l-1} (0 for lane 0) and adds the lane's rates in slot order, the chosen
 * slot being the first whose running sum exceeds target (if none: the lane's last non-zero rate).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HFX_API __attribute__((visibility("default")))

enum { X_SINGLET = 1, X_TRIPLET = 3 };
enum { EV_MOVE = 1, EV_ISC = 3, EV_FORSTER = 4, EV_DECAY = 5 };
enum { STREAM_XTRAJ = 4 };

/* tallies (u64), same indices as HF_XT_* in include/hfkmc.h */
enum {
    XT_GENERATED = 0, XT_EVENTS = 1, XT_FORSTER = 2, XT_DEXTER = 3, XT_ISC = 4, XT_RISC = 5,
    XT_RAD = 6,        /* +species: photons by emitting species            */
    XT_NONRAD_S = 9,   /* +species: non-radiative singlet losses           */
    XT_NONRAD_T = 12,  /* +species: non-radiative triplet losses           */
    XT_N = 16
};

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* both uniforms of block b of (stream, ident): the draw contract of oracle/philox.py */
static void block_uniforms(uint64_t seed, uint32_t stream, uint32_t ident, uint64_t b, double *u0, double *u1) {
    uint32_t c[4] = {(uint32_t)b, (uint32_t)(b >> 32), ident, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *u0 = (double)((((uint64_t)c[1] << 32) | c[0]) >> 11) * 0x1p-53;
    *u1 = (double)((((uint64_t)c[3] << 32) | c[2]) >> 11) * 0x1p-53;
}

typedef struct {
    int X, Y, Z;
    const uint8_t *species;
    const double *ws1, *wt1;
    int M;                       /* neighbour offsets                                     */
    const int8_t *offsets;       /* [M][3] dx, dy, dz                                     */
    const double *g6, *gd;       /* [M] r^-6 and exp(-2 r / L)                            */
    const double *kf, *kd;       /* [3][3] donor x acceptor prefactors                    */
    const double *k_isc, *k_risc;/* [3]                                                   */
    const double *k_rad, *k_nonrad; /* [3][2] species x (singlet, triplet)               */
    double singlet_fraction;
} hfx_model;

typedef struct {
    int x, y, z, spin;
    double time;                 /* since this exciton was created                        */
    uint64_t draws;
    uint64_t tallies[XT_N];
    double lifetime_sum;
} hfx_state;

static void spawn(const hfx_model *m, hfx_state *s, uint64_t seed, uint32_t ident) {
    double ua, ub;
    block_uniforms(seed, STREAM_XTRAJ, ident, s->draws >> 1, &ua, &ub);
    s->draws += 2;
    const int64_t n_interior = (int64_t)m->X * m->Y * (m->Z - 2);
    int64_t k = (int64_t)floor(ua * (double)n_interior);
    if (k >= n_interior) k = n_interior - 1;
    s->x = (int)(k % m->X); s->y = (int)((k / m->X) % m->Y); s->z = 1 + (int)(k / ((int64_t)m->X * m->Y));
    s->spin = ub < m->singlet_fraction ? X_SINGLET : X_TRIPLET;      /* molecule.py:93 */
    s->time = 0.0;
    s->tallies[XT_GENERATED] += 1;
}

static double min1(double v) { return v < 1.0 ? v : 1.0; }
static int tag_of(double w) { uint64_t b; memcpy(&b, &w, 8); return (int)(b & 3); }

/* One BKL event.  rates is a scratch array of (M + 3 rounded up to 32) doubles. */
static void step(const hfx_model *m, hfx_state *s, uint64_t seed, uint32_t ident, double *rates, double *unused,
                 int *ev_kind, int64_t *ev_initial, int64_t *ev_final, double *ev_dt, int *ev_channel) {
    (void)unused;
    const int64_t plane = (int64_t)m->X * m->Y;
    const int64_t i = ((int64_t)s->z * m->Y + s->y) * m->X + s->x;
    const int sidx = s->spin == X_SINGLET ? 0 : 1;
    const double *w = sidx == 0 ? m->ws1 : m->wt1;
    const double *g = sidx == 0 ? m->g6 : m->gd;
    const int a = tag_of(w[i]);
    const double *pref = (sidx == 0 ? m->kf : m->kd) + 3 * a;
    const double inv_wi = 1.0 / w[i];
    const int C = m->M + 3, padded = (C + 31) & ~31, K = padded / 32;
    for (int c = 0; c < padded; ++c) rates[c] = 0.0;
    for (int c = 0; c < m->M; ++c) {
        const int zz = s->z + m->offsets[3 * c + 2];
        if (zz < 0 || zz >= m->Z) continue;
        int xx = s->x + m->offsets[3 * c + 0], yy = s->y + m->offsets[3 * c + 1];
        if (xx < 0) xx += m->X; else if (xx >= m->X) xx -= m->X;
        if (yy < 0) yy += m->Y; else if (yy >= m->Y) yy -= m->Y;
        const int64_t j = (int64_t)zz * plane + (int64_t)yy * m->X + xx;
        rates[c] = (pref[tag_of(w[j])] * g[c]) * min1(w[j] * inv_wi);
    }
    rates[m->M] = sidx == 0 ? m->k_isc[a] : m->k_risc[a] * min1(m->ws1[i] * (1.0 / m->wt1[i]));
    rates[m->M + 1] = m->k_rad[2 * a + sidx];
    rates[m->M + 2] = m->k_nonrad[2 * a + sidx];
    /* canonical order: per-lane sums in slot order, Kogge-Stone scan over the lanes */
    double S[32], L[32], nxt[32];
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = acc + rates[k * 32 + l];
        S[l] = acc;
    }
    memcpy(L, S, sizeof(L));
    for (int off = 1; off < 32; off <<= 1) {
        for (int l = 0; l < 32; ++l) nxt[l] = l >= off ? L[l] + L[l - off] : L[l];
        memcpy(L, nxt, sizeof(L));
    }
    const double R = L[31];
    double u_sel, u_time;
    block_uniforms(seed, STREAM_XTRAJ, ident, s->draws >> 1, &u_sel, &u_time);
    s->draws += 2;
    const double target = u_sel * R;
    int lane = -1;
    for (int l = 0; l < 32; ++l) if (L[l] > target) { lane = l; break; }
    if (lane < 0) for (int l = 31; l >= 0; --l) if (S[l] > 0.0) { lane = l; break; }
    int slot = -1, last_nonzero = -1;
    double acc = lane > 0 ? L[lane - 1] : 0.0;
    for (int k = 0; k < K; ++k) {
        const double r = rates[k * 32 + lane];
        acc = acc + r;
        if (r > 0.0) last_nonzero = k;
        if (slot < 0 && acc > target) slot = k;
    }
    if (slot < 0) slot = last_nonzero;
    const int chosen = slot * 32 + lane;
    (void)C;
    const double dt = -log(1.0 - u_time) / R;
    s->time += dt;
    s->tallies[XT_EVENTS] += 1;
    *ev_initial = i; *ev_final = i; *ev_dt = dt; *ev_channel = chosen;
    if (chosen < m->M) {
        int xx = s->x + m->offsets[3 * chosen + 0], yy = s->y + m->offsets[3 * chosen + 1];
        if (xx < 0) xx += m->X; else if (xx >= m->X) xx -= m->X;
        if (yy < 0) yy += m->Y; else if (yy >= m->Y) yy -= m->Y;
        s->x = xx; s->y = yy; s->z += m->offsets[3 * chosen + 2];
        *ev_final = ((int64_t)s->z * m->Y + s->y) * m->X + s->x;
        if (sidx == 0) { *ev_kind = EV_FORSTER; s->tallies[XT_FORSTER] += 1; }
        else { *ev_kind = EV_MOVE; s->tallies[XT_DEXTER] += 1; }
    } else if (chosen == m->M) {
        *ev_kind = EV_ISC;                                             /* molecule.py:299-307 */
        if (sidx == 0) { s->spin = X_TRIPLET; s->tallies[XT_ISC] += 1; }
        else { s->spin = X_SINGLET; s->tallies[XT_RISC] += 1; }
    } else {
        *ev_kind = EV_DECAY;
        if (chosen == m->M + 1) s->tallies[XT_RAD + a] += 1;
        else s->tallies[(sidx == 0 ? XT_NONRAD_S : XT_NONRAD_T) + a] += 1;
        s->lifetime_sum += s->time;
        spawn(m, s, seed, ident);
    }
}

/* Run one trajectory for n_events BKL events from a fresh spawn.  Optional trace arrays of
 * n_events entries.  tallies_out[XT_N]; state_out = {site, spin, draws}. */
HFX_API void hfx_run(int X, int Y, int Z, const uint8_t *species, const double *ws1, const double *wt1,
                     int M, const int8_t *offsets, const double *g6, const double *gd, const double *kf,
                     const double *kd, const double *k_isc, const double *k_risc, const double *k_rad,
                     const double *k_nonrad, double singlet_fraction, uint64_t seed, uint32_t trajectory,
                     int64_t n_events, uint8_t *tr_kind, int32_t *tr_initial, int32_t *tr_final, double *tr_dt,
                     int32_t *tr_channel, uint8_t *tr_spin, uint64_t *tallies_out, int64_t *state_out,
                     double *lifetime_sum_out) {
    hfx_model m = {X, Y, Z, species, ws1, wt1, M, offsets, g6, gd, kf, kd, k_isc, k_risc, k_rad, k_nonrad, singlet_fraction};
    hfx_state s;
    memset(&s, 0, sizeof(s));
    const int padded = (M + 3 + 31) & ~31;
    double *rates = (double *)malloc(sizeof(double) * (size_t)padded * 2), *prefix = rates + padded;
    spawn(&m, &s, seed, trajectory);
    for (int64_t e = 0; e < n_events; ++e) {
        int kind = 0, channel = 0;
        int64_t ini = 0, fin = 0;
        double dt = 0.0;
        step(&m, &s, seed, trajectory, rates, prefix, &kind, &ini, &fin, &dt, &channel);
        if (tr_kind) {
            tr_kind[e] = (uint8_t)kind; tr_initial[e] = (int32_t)ini; tr_final[e] = (int32_t)fin;
            tr_dt[e] = dt; tr_channel[e] = channel; tr_spin[e] = (uint8_t)s.spin;
        }
    }
    memcpy(tallies_out, s.tallies, sizeof(s.tallies));
    state_out[0] = ((int64_t)s.z * Y + s.y) * X + s.x; state_out[1] = s.spin; state_out[2] = (int64_t)s.draws;
    *lifetime_sum_out = s.lifetime_sum;
    free(rates);
}

/* CPU baseline: trajectories [first, first + n) for n_events each on n_threads POSIX threads. */
#include <pthread.h>
typedef struct {
    hfx_model m; uint64_t seed; int64_t first, n, n_events; int64_t *next; uint64_t tallies[XT_N];
} hfx_job;

static void *hfx_worker(void *arg) {
    hfx_job *J = (hfx_job *)arg;
    const int padded = (J->m.M + 3 + 31) & ~31;
    double *rates = (double *)malloc(sizeof(double) * (size_t)padded * 2), *prefix = rates + padded;
    for (;;) {
        int64_t t = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED);
        if (t >= J->n) break;
        hfx_state s;
        memset(&s, 0, sizeof(s));
        spawn(&J->m, &s, J->seed, (uint32_t)(J->first + t));
        for (int64_t e = 0; e < J->n_events; ++e) {
            int kind, channel; int64_t ini, fin; double dt;
            step(&J->m, &s, J->seed, (uint32_t)(J->first + t), rates, prefix, &kind, &ini, &fin, &dt, &channel);
        }
        for (int k = 0; k < XT_N; ++k) J->tallies[k] += s.tallies[k];
    }
    free(rates);
    return NULL;
}

HFX_API void hfx_run_batch(int X, int Y, int Z, const uint8_t *species, const double *ws1, const double *wt1,
                           int M, const int8_t *offsets, const double *g6, const double *gd, const double *kf,
                           const double *kd, const double *k_isc, const double *k_risc, const double *k_rad,
                           const double *k_nonrad, double singlet_fraction, uint64_t seed, int64_t first,
                           int64_t n, int64_t n_events, int n_threads, uint64_t *tallies_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    int64_t next = 0;
    hfx_job *jobs = (hfx_job *)calloc((size_t)n_threads, sizeof(hfx_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        hfx_model m = {X, Y, Z, species, ws1, wt1, M, offsets, g6, gd, kf, kd, k_isc, k_risc, k_rad, k_nonrad, singlet_fraction};
        jobs[i].m = m; jobs[i].seed = seed; jobs[i].first = first; jobs[i].n = n; jobs[i].n_events = n_events; jobs[i].next = &next;
        if (i > 0) pthread_create(&th[i], NULL, hfx_worker, &jobs[i]);
    }
    hfx_worker(&jobs[0]);
    memset(tallies_out, 0, sizeof(uint64_t) * XT_N);
    for (int i = 0; i < n_threads; ++i) {
        if (i > 0) pthread_join(th[i], NULL);
        for (int k = 0; k < XT_N; ++k) tallies_out[k] += jobs[i].tallies[k];
    }
    free(jobs); free(th);
}
