/*
 * exciton_dense_oracle.c — CPU statement of the HIGH-DENSITY exciton model ("Path XD", SURVEY.md §8 f4,
 * BASELINE configs[3]: occupancy blocking and exciton-exciton annihilation).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (only tests/ and bench.py's CPU-baseline legs load it).
 *
 * PARITY UNPINNED.  The reference has no exciton dynamics at all (see exciton_oracle.c); its only
 * occupancy rule is same-type charge blocking (reseau.py:385, 394) and it has no annihilation.
 * This file is the specification the CUDA kernel hfxd_run_kernel is tested against, bit for bit.
 *
 * Model.  A DEVICE holds n excitons on one lattice realisation; devices are independent.  Every
 * site carries a 2-bit occupancy code (0 empty, 1 singlet, 3 triplet: the EXCITON codes of
 * molecule.py:16-21).  The channels of an exciton are those of exciton_oracle.c; a transfer
 * channel towards an OCCUPIED site j is an annihilation channel:
 *     rate = ((pref[b] * g) * min(1, w_j / w_i)) * ann[s][s_j]
 * (ann = 0: pure blocking).  When it fires the donor is lost; for triplet on triplet the acceptor
 * becomes a singlet with probability p_tta (one more uniform).  A lost exciton (decay or
 * annihilation) is replaced at once by a new one on a random EMPTY interior site, so the density
 * is stationary.
 *
 * One KMC event of a device (rejection-free over all channels of all its excitons):
 *   T_k    total rate of exciton k (canonical lane / slot / Kogge-Stone sum of exciton_oracle.c);
 *          T_k = 0 for the padding slots n <= k < n_pad, n_pad = n rounded up to 32
 *   A_l    = T_l + T_{32+l} + ... in slot order;  G = Kogge-Stone scan of A over the lanes;  R = G_31
 *   (u_sel, u_time) = Philox block draws/2 of stream 5 (XDENSE) of the device;  target = u_sel * R
 *   lane l* = first with G_l > target (else the last with A_l > 0); inside it acc starts from
 *   G_{l*-1} (0 for lane 0) and adds T in slot order: k* = first slot whose running sum exceeds
 *   target (else the lane's last non-zero), base = the running sum before it
 *   the channel of exciton k* is chosen by the rule of exciton_oracle.c with target2 = target - base
 *   dt = -log(1 - u_time) / R
 * Replacement draws: block after block of the same stream, site = floor(u_a * n_interior) until it
 * is empty, spin singlet iff u_b < singlet_fraction.  The TTA uniform is the first half of its
 * own block.  Initial state: slots 0..n-1 filled by the replacement rule, in order.
 *
 * Cached totals: T_k only changes when exciton k itself or a site within its cutoff sphere changes.
 * hfxd_run keeps the cache exactly like the device does and, with verify != 0, checks it against a
 * from-scratch evaluation of every exciton after every event.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HFX_API __attribute__((visibility("default")))

enum { X_SINGLET = 1, X_TRIPLET = 3 };
enum { EV_MOVE = 1, EV_ISC = 3, EV_FORSTER = 4, EV_DECAY = 5, EV_ANNIHILATION = 6 };   /* 6: the `unbound` slot of event.py:49 */
enum { STREAM_XDENSE = 5 };
enum {
    XT_GENERATED = 0, XT_EVENTS = 1, XT_FORSTER = 2, XT_DEXTER = 3, XT_ISC = 4, XT_RISC = 5,
    XT_RAD = 6, XT_NONRAD_S = 9, XT_NONRAD_T = 12,
    XT_ANN = 16,       /* + 2 * donor spin index + acceptor spin index */
    XT_TTA_UP = 20,
    XT_N = 24
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

static void block_uniforms(uint64_t seed, uint32_t stream, uint32_t ident, uint64_t b, double *u0, double *u1) {
    uint32_t c[4] = {(uint32_t)b, (uint32_t)(b >> 32), ident, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *u0 = (double)((((uint64_t)c[1] << 32) | c[0]) >> 11) * 0x1p-53;
    *u1 = (double)((((uint64_t)c[3] << 32) | c[2]) >> 11) * 0x1p-53;
}

typedef struct {
    int X, Y, Z;
    const double *ws1, *wt1;
    int M, R;
    const int8_t *offsets;
    const double *g6, *gd, *kf, *kd, *k_isc, *k_risc, *k_rad, *k_nonrad;
    double singlet_fraction;
    const double *ann;           /* [2][2] donor spin index x acceptor spin index */
    double p_tta;
} xd_model;

typedef struct {
    int n, n_pad;
    int *x, *y, *z, *spin;
    double *birth, *T;
    uint8_t *occ;                /* [N] occupancy codes */
    double time, lifetime_sum;
    uint64_t draws;
    uint64_t tallies[XT_N];
} xd_state;

static double min1(double v) { return v < 1.0 ? v : 1.0; }
static int tag_of(double w) { uint64_t b; memcpy(&b, &w, 8); return (int)(b & 3); }
static int64_t site_of(const xd_model *m, int x, int y, int z) { return ((int64_t)z * m->Y + y) * m->X + x; }

/* rates[padded] of the exciton in slot k */
static void rates_of(const xd_model *m, const xd_state *s, int k, double *rates) {
    const int sidx = s->spin[k] == X_SINGLET ? 0 : 1;
    const double *w = sidx == 0 ? m->ws1 : m->wt1, *g = sidx == 0 ? m->g6 : m->gd;
    const int64_t i = site_of(m, s->x[k], s->y[k], s->z[k]);
    const int a = tag_of(w[i]);
    const double *pref = (sidx == 0 ? m->kf : m->kd) + 3 * a;
    const double inv_wi = 1.0 / w[i];
    const int padded = (m->M + 3 + 31) & ~31;
    for (int c = 0; c < padded; ++c) rates[c] = 0.0;
    for (int c = 0; c < m->M; ++c) {
        const int zz = s->z[k] + m->offsets[3 * c + 2];
        if (zz < 0 || zz >= m->Z) continue;
        int xx = s->x[k] + m->offsets[3 * c + 0], yy = s->y[k] + m->offsets[3 * c + 1];
        if (xx < 0) xx += m->X; else if (xx >= m->X) xx -= m->X;
        if (yy < 0) yy += m->Y; else if (yy >= m->Y) yy -= m->Y;
        const int64_t j = site_of(m, xx, yy, zz);
        double r = (pref[tag_of(w[j])] * g[c]) * min1(w[j] * inv_wi);
        if (s->occ[j]) r = r * m->ann[2 * sidx + (s->occ[j] == X_SINGLET ? 0 : 1)];
        rates[c] = r;
    }
    rates[m->M] = sidx == 0 ? m->k_isc[a] : m->k_risc[a] * min1(m->ws1[i] * (1.0 / m->wt1[i]));
    rates[m->M + 1] = m->k_rad[2 * a + sidx];
    rates[m->M + 2] = m->k_nonrad[2 * a + sidx];
}

/* canonical lane sums and their Kogge-Stone scan over n_slots x 32 values */
static void lane_scan(const double *v, int n_slots, double *S, double *L) {
    double nxt[32];
    for (int l = 0; l < 32; ++l) {
        double acc = 0.0;
        for (int q = 0; q < n_slots; ++q) acc = acc + v[q * 32 + l];
        S[l] = acc;
    }
    memcpy(L, S, sizeof(double) * 32);
    for (int off = 1; off < 32; off <<= 1) {
        for (int l = 0; l < 32; ++l) nxt[l] = l >= off ? L[l] + L[l - off] : L[l];
        memcpy(L, nxt, sizeof(nxt));
    }
}

/* first entry whose canonical running sum exceeds target; *base = the running sum before it */
static int pick(const double *v, int n_slots, const double *S, const double *L, double target, double *base) {
    int lane = -1;
    for (int l = 0; l < 32; ++l) if (L[l] > target) { lane = l; break; }
    if (lane < 0) for (int l = 31; l >= 0; --l) if (S[l] > 0.0) { lane = l; break; }
    if (lane < 0) lane = 0;
    int slot = -1, last = -1;
    double acc = lane > 0 ? L[lane - 1] : 0.0, b_slot = acc, b_last = acc;
    for (int q = 0; q < n_slots; ++q) {
        const double r = v[q * 32 + lane], before = acc;
        acc = acc + r;
        if (r > 0.0) { last = q; b_last = before; }
        if (slot < 0 && acc > target) { slot = q; b_slot = before; }
    }
    if (slot < 0) { slot = last < 0 ? n_slots - 1 : last; b_slot = b_last; }
    *base = b_slot;
    return slot * 32 + lane;
}

static double total_of(const xd_model *m, const xd_state *s, int k, double *rates) {
    double S[32], L[32];
    rates_of(m, s, k, rates);
    lane_scan(rates, ((m->M + 3 + 31) & ~31) / 32, S, L);
    return L[31];
}

/* put a new exciton into slot k: first empty interior site along the device's stream */
static int64_t replace(const xd_model *m, xd_state *s, uint64_t seed, uint32_t ident, int k) {
    const int64_t n_interior = (int64_t)m->X * m->Y * (m->Z - 2);
    for (;;) {
        double ua, ub;
        block_uniforms(seed, STREAM_XDENSE, ident, s->draws >> 1, &ua, &ub);
        s->draws += 2;
        int64_t q = (int64_t)floor(ua * (double)n_interior);
        if (q >= n_interior) q = n_interior - 1;
        const int x = (int)(q % m->X), y = (int)((q / m->X) % m->Y), z = 1 + (int)(q / ((int64_t)m->X * m->Y));
        const int64_t site = site_of(m, x, y, z);
        if (s->occ[site]) continue;
        s->x[k] = x; s->y[k] = y; s->z[k] = z;
        s->spin[k] = ub < m->singlet_fraction ? X_SINGLET : X_TRIPLET;
        s->occ[site] = (uint8_t)s->spin[k];
        s->birth[k] = s->time;
        s->tallies[XT_GENERATED] += 1;
        return site;
    }
}

/* does the exciton in slot k see site (cx, cy, cz) inside its cutoff sphere (or sit on it)? */
static int sees(const xd_model *m, const xd_state *s, int k, int cx, int cy, int cz, double rc2) {
    int dx = cx - s->x[k], dy = cy - s->y[k];
    const int dz = cz - s->z[k];
    if (dx > m->X / 2) dx -= m->X; else if (dx < -(m->X / 2)) dx += m->X;     /* minimal image; X, Y >= 2R + 1 */
    if (dy > m->Y / 2) dy -= m->Y; else if (dy < -(m->Y / 2)) dy += m->Y;
    if (m->X % 2 == 0 && (dx == m->X / 2 || dx == -(m->X / 2))) return 0;     /* farther than R on an even ring */
    if (m->Y % 2 == 0 && (dy == m->Y / 2 || dy == -(m->Y / 2))) return 0;
    return (double)(dx * dx + dy * dy + dz * dz) <= rc2;
}

typedef struct { uint8_t kind, spin_after, annihilation; int32_t exciton, initial, final_, channel; double dt; } xd_event;

static void step(const xd_model *m, xd_state *s, uint64_t seed, uint32_t ident, double cutoff, double *rates, xd_event *ev) {
    const int n_slots = s->n_pad / 32, K = ((m->M + 3 + 31) & ~31) / 32;
    double A[32], G[32], S[32], L[32], base = 0.0, base2 = 0.0;
    lane_scan(s->T, n_slots, A, G);
    const double R = G[31];
    double u_sel, u_time;
    block_uniforms(seed, STREAM_XDENSE, ident, s->draws >> 1, &u_sel, &u_time);
    s->draws += 2;
    const double target = u_sel * R;
    const int k = pick(s->T, n_slots, A, G, target, &base);
    const double target2 = target - base;
    rates_of(m, s, k, rates);
    lane_scan(rates, K, S, L);
    const int chosen = pick(rates, K, S, L, target2, &base2);
    const double dt = -log(1.0 - u_time) / R;
    s->time += dt;
    s->tallies[XT_EVENTS] += 1;

    const int sidx = s->spin[k] == X_SINGLET ? 0 : 1;
    const int ix = s->x[k], iy = s->y[k], iz = s->z[k];
    const int64_t i = site_of(m, ix, iy, iz);
    const int a = tag_of((sidx == 0 ? m->ws1 : m->wt1)[i]);
    int changed[3][3], n_changed = 0;
    changed[n_changed][0] = ix; changed[n_changed][1] = iy; changed[n_changed][2] = iz; ++n_changed;
    int lost = 0;
    ev->exciton = k; ev->initial = (int32_t)i; ev->final_ = (int32_t)i; ev->channel = chosen; ev->dt = dt; ev->annihilation = 0;
    if (chosen < m->M) {
        int xx = ix + m->offsets[3 * chosen + 0], yy = iy + m->offsets[3 * chosen + 1];
        const int zz = iz + m->offsets[3 * chosen + 2];
        if (xx < 0) xx += m->X; else if (xx >= m->X) xx -= m->X;
        if (yy < 0) yy += m->Y; else if (yy >= m->Y) yy -= m->Y;
        const int64_t j = site_of(m, xx, yy, zz);
        ev->final_ = (int32_t)j;
        changed[n_changed][0] = xx; changed[n_changed][1] = yy; changed[n_changed][2] = zz; ++n_changed;
        if (s->occ[j] == 0) {
            s->occ[i] = 0; s->occ[j] = (uint8_t)s->spin[k];
            s->x[k] = xx; s->y[k] = yy; s->z[k] = zz;
            if (sidx == 0) { ev->kind = EV_FORSTER; s->tallies[XT_FORSTER] += 1; }
            else { ev->kind = EV_MOVE; s->tallies[XT_DEXTER] += 1; }
        } else {
            const int oidx = s->occ[j] == X_SINGLET ? 0 : 1;
            ev->kind = EV_ANNIHILATION; ev->annihilation = 1;
            s->tallies[XT_ANN + 2 * sidx + oidx] += 1;
            if (sidx == 1 && oidx == 1) {                       /* triplet-triplet: the survivor may be promoted */
                double u, unused;
                block_uniforms(seed, STREAM_XDENSE, ident, s->draws >> 1, &u, &unused);
                s->draws += 2;
                if (u < m->p_tta) {
                    for (int q = 0; q < s->n; ++q)
                        if (q != k && s->x[q] == xx && s->y[q] == yy && s->z[q] == zz) s->spin[q] = X_SINGLET;
                    s->occ[j] = X_SINGLET;
                    s->tallies[XT_TTA_UP] += 1;
                }
            }
            lost = 1;
        }
    } else if (chosen == m->M) {
        ev->kind = EV_ISC;
        if (sidx == 0) { s->spin[k] = X_TRIPLET; s->tallies[XT_ISC] += 1; }
        else { s->spin[k] = X_SINGLET; s->tallies[XT_RISC] += 1; }
        s->occ[i] = (uint8_t)s->spin[k];
    } else {
        ev->kind = EV_DECAY;
        if (chosen == m->M + 1) s->tallies[XT_RAD + a] += 1;
        else s->tallies[(sidx == 0 ? XT_NONRAD_S : XT_NONRAD_T) + a] += 1;
        lost = 1;
    }
    if (lost) {
        s->occ[i] = 0;
        s->lifetime_sum += s->time - s->birth[k];
        replace(m, s, seed, ident, k);
        changed[n_changed][0] = s->x[k]; changed[n_changed][1] = s->y[k]; changed[n_changed][2] = s->z[k]; ++n_changed;
    }
    ev->spin_after = (uint8_t)s->spin[k];
    /* refresh the cached totals of every exciton that sees a changed site */
    const double rc2 = cutoff * cutoff;
    for (int q = 0; q < s->n; ++q) {
        int dirty = q == k;
        for (int c = 0; c < n_changed && !dirty; ++c) dirty = sees(m, s, q, changed[c][0], changed[c][1], changed[c][2], rc2);
        if (dirty) s->T[q] = total_of(m, s, q, rates);
    }
}

/* One device: n excitons, n_events events from a fresh fill.  Optional trace arrays (n_events).
 * pos_out[n] / spin_out[n] / totals_out[n]: final slots; state_out = {draws}; returns the number of
 * cache mismatches found (0 unless verify and a bug). */
HFX_API int64_t hfxd_run(int X, int Y, int Z, const double *ws1, const double *wt1, int M, const int8_t *offsets,
                         const double *g6, const double *gd, const double *kf, const double *kd, const double *k_isc,
                         const double *k_risc, const double *k_rad, const double *k_nonrad, double singlet_fraction,
                         double cutoff, const double *ann, double p_tta, uint64_t seed, uint32_t device, int n,
                         int64_t n_events, int verify, uint8_t *tr_kind, int32_t *tr_exciton, int32_t *tr_initial,
                         int32_t *tr_final, int32_t *tr_channel, double *tr_dt, uint8_t *tr_spin, uint64_t *tallies_out,
                         int32_t *pos_out, int32_t *spin_out, double *totals_out, double *scalars_out) {
    int maxr = 0;
    for (int c = 0; c < 3 * M; ++c) { const int v = offsets[c] < 0 ? -offsets[c] : offsets[c]; if (v > maxr) maxr = v; }
    xd_model m = {X, Y, Z, ws1, wt1, M, maxr, offsets, g6, gd, kf, kd, k_isc, k_risc, k_rad, k_nonrad, singlet_fraction, ann, p_tta};
    xd_state s;
    memset(&s, 0, sizeof(s));
    s.n = n; s.n_pad = (n + 31) & ~31;
    const int64_t N = (int64_t)X * Y * Z;
    s.x = (int *)calloc((size_t)s.n_pad * 4, sizeof(int)); s.y = s.x + s.n_pad; s.z = s.y + s.n_pad; s.spin = s.z + s.n_pad;
    s.birth = (double *)calloc((size_t)s.n_pad * 2, sizeof(double)); s.T = s.birth + s.n_pad;
    s.occ = (uint8_t *)calloc((size_t)N, 1);
    const int padded = (M + 3 + 31) & ~31;
    double *rates = (double *)malloc(sizeof(double) * (size_t)padded);
    int64_t mismatches = 0;
    for (int k = 0; k < n; ++k) replace(&m, &s, seed, device, k);
    for (int k = 0; k < n; ++k) s.T[k] = total_of(&m, &s, k, rates);
    for (int64_t e = 0; e < n_events; ++e) {
        xd_event ev;
        step(&m, &s, seed, device, cutoff, rates, &ev);
        if (tr_kind) {
            tr_kind[e] = ev.kind; tr_exciton[e] = ev.exciton; tr_initial[e] = ev.initial; tr_final[e] = ev.final_;
            tr_channel[e] = ev.channel; tr_dt[e] = ev.dt; tr_spin[e] = ev.spin_after;
        }
        if (verify)
            for (int k = 0; k < n; ++k) {
                const double t = total_of(&m, &s, k, rates);
                if (memcmp(&t, &s.T[k], 8) != 0) ++mismatches;
            }
    }
    memcpy(tallies_out, s.tallies, sizeof(s.tallies));
    for (int k = 0; k < n; ++k) {
        if (pos_out) pos_out[k] = (int32_t)site_of(&m, s.x[k], s.y[k], s.z[k]);
        if (spin_out) spin_out[k] = s.spin[k];
        if (totals_out) totals_out[k] = s.T[k];
    }
    scalars_out[0] = s.time; scalars_out[1] = s.lifetime_sum; scalars_out[2] = (double)s.draws;
    free(s.x); free(s.birth); free(s.occ); free(rates);
    return mismatches;
}

/* CPU baseline: devices [first, first + n_devices) on n_threads POSIX threads; summed tallies. */
#include <pthread.h>
typedef struct {
    int X, Y, Z, M, n; const double *ws1, *wt1; const int8_t *offsets;
    const double *g6, *gd, *kf, *kd, *k_isc, *k_risc, *k_rad, *k_nonrad, *ann;
    double singlet_fraction, cutoff, p_tta; uint64_t seed; int64_t first, n_devices, n_events; int64_t *next;
    uint64_t tallies[XT_N];
} xd_job;

static void *xd_worker(void *arg) {
    xd_job *J = (xd_job *)arg;
    for (;;) {
        const int64_t d = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED);
        if (d >= J->n_devices) break;
        uint64_t t[XT_N];
        double scal[3];
        hfxd_run(J->X, J->Y, J->Z, J->ws1, J->wt1, J->M, J->offsets, J->g6, J->gd, J->kf, J->kd, J->k_isc, J->k_risc, J->k_rad,
                 J->k_nonrad, J->singlet_fraction, J->cutoff, J->ann, J->p_tta, J->seed, (uint32_t)(J->first + d), J->n,
                 J->n_events, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, t, NULL, NULL, NULL, scal);
        for (int k = 0; k < XT_N; ++k) J->tallies[k] += t[k];
    }
    return NULL;
}

HFX_API void hfxd_run_batch(int X, int Y, int Z, const double *ws1, const double *wt1, int M, const int8_t *offsets,
                            const double *g6, const double *gd, const double *kf, const double *kd, const double *k_isc,
                            const double *k_risc, const double *k_rad, const double *k_nonrad, double singlet_fraction,
                            double cutoff, const double *ann, double p_tta, uint64_t seed, int64_t first, int64_t n_devices,
                            int n, int64_t n_events, int n_threads, uint64_t *tallies_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    int64_t next = 0;
    xd_job *jobs = (xd_job *)calloc((size_t)n_threads, sizeof(xd_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        xd_job j = {X, Y, Z, M, n, ws1, wt1, offsets, g6, gd, kf, kd, k_isc, k_risc, k_rad, k_nonrad, ann,
                    singlet_fraction, cutoff, p_tta, seed, first, n_devices, n_events, &next, {0}};
        jobs[i] = j;
        if (i > 0) pthread_create(&th[i], NULL, xd_worker, &jobs[i]);
    }
    xd_worker(&jobs[0]);
    memset(tallies_out, 0, sizeof(uint64_t) * XT_N);
    for (int i = 0; i < n_threads; ++i) {
        if (i > 0) pthread_join(th[i], NULL);
        for (int k = 0; k < XT_N; ++k) tallies_out[k] += jobs[i].tallies[k];
    }
    free(jobs); free(th);
}
