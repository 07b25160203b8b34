// First-Reaction-Method kernels for sm_100a: one warp owns one trajectory (one reference
// `Lattice` instance's dynamic state); the 32 lanes are the <= 26 candidate hops of the charge
// whose event is being regenerated (reseau.py:380-396), or cooperate on the ordered lists.
//
// State of a trajectory lives in shared memory for the whole launch (loaded once, stored once):
//   pos[t]      ordered position lists  _electrons_locations / _holes_locations (reseau.py:208-209)
//   flag[t]     the SET of sites whose Molecule.electron / .hole flag is True.  The reference's
//               flags are toggles (molecule.py:78-86) that can disagree with the lists (two
//               charges on one site after a stale move or a re-injection), so they are state of
//               their own, not derived from pos[].
//   ev_*[t]     ordered move-event lists _move_electron_events / _move_hole_events
//   special     the single pending bound / decay / capture event (tau = 0, fires next step)
//   cache       ring of the last 10 fired events (Lattice._cache, reseau.py:124)
// t = 0 electrons, t = 1 holes; particule code = t + 1 (event.py:53-57).
//
// fp64 arithmetic that decides parity uses explicit round-to-nearest intrinsics so that no FMA
// contraction can change a bit relative to CPython's float arithmetic; only exp() and log()
// (CUDA libdevice vs glibc, both <= 1 ulp) can differ, which is why tau is compared at 1e-12.
#pragma once
#include <math.h>
#include <stdint.h>

#include "hf_philox.cuh"

#define HF_FULL 0xffffffffu
#define HF_CACHE 10

// trajectory status (mirrors hf_traj_status in include/hfkmc.h)
#define HF_ST_OK 0
#define HF_ST_NO_EVENTS 1
#define HF_ST_ZERO_DIVISION 2
#define HF_ST_VALUE_ERROR 3
#define HF_ST_REPLAY_EXHAUSTED 4
#define HF_ST_CAPACITY 5

// event / particle / exciton / species codes (event.py:43-57, molecule.py:16-21)
#define HF_K_MOVE 1
#define HF_K_BOUND 2
#define HF_K_DECAY 5
#define HF_K_CAPTURE 7
#define HF_PART_EXCITON 3
#define HF_XS_SINGLET 1
#define HF_XS_TRIPLET 3
#define HF_SPECIES_HOST 0

// constants.py:3-6 and reseau.py:148-151, as CPython evaluates them
#define HF_KT 0x1.a78f25677e0f1p-6            /* BOLTZMANN * 300.            */
#define HF_ELECTROSTATIC 0x1.01b112ab38751p-21 /* 1/(4 pi eps0 eps_r)        */
#define HF_RATE0 1e13                          /* 10.**13                     */

struct HfScalars {                 // one per trajectory, global memory
    int32_t n_pos[2], n_flag[2], n_ev[2];
    int32_t special, special_site, xstate, status;
    int32_t cache_head, cache_n;
    uint64_t draws, spins, emission, recombination, injection, steps, fired;
};

struct HfTraceRec {                // == hf_event in include/hfkmc.h
    int32_t initial, final_;
    double tau;
    uint8_t kind, particule, pad[2];
    uint32_t spins;
    uint64_t draws;
};

// ---- integrated exciton dynamics (own model; functions in hf_frm_x.cuh) ----------------------------------------
#define HF_STREAM_PX 5u             /* the trajectory's exciton draws: never the TRAJ stream, whose count is pinned */
#define HF_K_ISC 3
#define HF_K_FORSTER 4
#define HFI_SPIN(m) ((m) & 3)
#define HFI_ONSITE 4
#define HFI_RAD 8
#define HFI_KIND(m) (((m) >> 4) & 15)

// Tables of the integrated mode 2, in GLOBAL memory (written once by hf_p_excitons): the event generator is a
// separate, non-inlined function that must not take the address of a kernel parameter (that would copy the whole
// parameter block to the stack) nor share the step kernels' register allocation.
struct HfiTables {
    int32_t M, X, Y, Z;
    uint32_t k0, k1;
    const int32_t *offsets;           // [M] packed (dx+128) | (dy+128) << 8 | (dz+128) << 16   (hf_x_configure)
    const double *g6, *gd;            // [M]
    const double *ws1, *wt1;          // [n_real][w_stride] tagged weights, zero halo of w_halo sites below and above
    int64_t w_stride, w_halo;
    double k[36];                     // kf[9] kd[9] k_isc[3] k_risc[3] k_rad[6] k_nonrad[6]
};

struct HfIntegrated {
    int32_t mode;
    const HfiTables *tab;             // mode 2 only
    // state [B][cap_c] / [B]
    int32_t *x_site, *x_meta, *x_final, *x_n;
    double *x_tau, *x_clock;
    uint64_t *x_draws;
};

struct HfParams {
    int32_t X, Y, Z;
    int64_t N;
    double field;
    int32_t C, cap_c, cap_f;
    int32_t dist;                  // charge_tranfer_distance (reseau.py:110): 1 = the 3x3x3 table, > 1 = hf_gen_move_event_far
    int64_t B, traj_per_real;
    uint32_t k0, k1, first_traj, first_real;
    uint32_t rk[20];               // the ten Philox round keys of (k0, k1): hf_philox_round_keys
    int64_t setsize;               // random.sample's pool/set switch for k = C (random.py:421-424)
    // lattice [n_real][N]
    const uint8_t *species;
    const double *homo, *lumo;
    // state
    int32_t *pos[2], *flag[2], *ev_init[2], *ev_final[2];
    double *ev_tau[2];
    HfScalars *sc;
    int32_t *cache_init, *cache_final, *cache_kp;   // [10][B] (slot-major)
    double *cache_tau;
    // replay / trace
    const double *replay;
    int64_t n_replay;
    HfTraceRec *trace;
    const uint64_t *trace_base;    // [trace_n] HfScalars.fired of each traced trajectory when hf_set_trace armed the trace
    int64_t trace_first, trace_n, trace_cap;
    unsigned long long *totals;    // channel counters, HF_N_TALLIES
    int64_t n_steps;
    HfIntegrated xi;               // xi.mode == 0: the reference as it is
    // 1 / sqrt(d2) for every squared distance of the lattice (IEEE division and square root on the host: the same
    // bits as __ddiv_rn(1.0, __dsqrt_rn(d2))), or null.  The warp kernel's Coulomb sums take 2 (2C - 1) of them per
    // verified candidate (reseau.py:299-312): one L2-resident load instead of ~60 instructions each.
    const double *inv_dist;
};

struct HfWarp {
    // shared-memory lists; the two particle types are adjacent: list(t) = base + t * capacity
    double *ev_tau_b, *inv_old_b, *term_b, *cache_tau;
    int32_t *pos_b, *flag_b, *ev_init_b, *ev_final_b, *cache_init, *cache_final, *cache_kp;
    int32_t cap_c, cap_f;
    int64_t cache_stride;          // 1: the ring lives in shared memory (warp kernel); B: it is written straight to the
                                   // slot-major global arrays (thread-per-trajectory kernel: coalesced across the warp)
    // warp-uniform scalars (kept as named fields so that nothing is dynamically indexed)
    int32_t n_pos0, n_pos1, n_flag0, n_flag1, n_ev0, n_ev1;
    int32_t special, special_site, xstate, status, cache_head, cache_n;
    uint64_t draws, spins, emission, recombination, injection, steps, fired;
    uint32_t ident;                // global trajectory id (Philox ident)
    int64_t local;                 // local trajectory index
    const uint8_t *species;        // this trajectory's realisation
    const double *lumo, *homo;
    float *approx;                 // thread-per-trajectory kernel only: this thread's screening column (27 floats)
    // integrated excitons (P.xi.mode != 0): slots in creation order, see hf_frm_x.cuh
    int32_t *x_site, *x_meta, *x_final;
    double *x_tau, *x_scratch;     // x_scratch: 32 lane sums (warp kernel) / 8 prefix checkpoints (thread kernel)
    int32_t n_x;
    uint64_t xdraws;
    double clock;
    unsigned c_x_forster, c_x_dexter, c_x_isc, c_x_risc;
    unsigned c_move_e, c_move_h, c_bound, c_decay, c_capture_e, c_capture_h;
    unsigned c_recomb_host, c_recomb_tadf, c_recomb_fluo, c_emit_tadf, c_emit_fluo;

    __device__ __forceinline__ int32_t *pos(int t) const { return pos_b + t * cap_c; }
    __device__ __forceinline__ int32_t *flag(int t) const { return flag_b + t * cap_f; }
    __device__ __forceinline__ int32_t *ev_init(int t) const { return ev_init_b + t * cap_c; }
    __device__ __forceinline__ int32_t *ev_final(int t) const { return ev_final_b + t * cap_c; }
    __device__ __forceinline__ double *ev_tau(int t) const { return ev_tau_b + t * cap_c; }
    __device__ __forceinline__ double *inv_old(int s) const { return inv_old_b + s * cap_c; }
    __device__ __forceinline__ double *term(int s) const { return term_b + s * cap_c; }
    __device__ __forceinline__ const double *energy(int t) const { return t ? homo : lumo; }
    __device__ __forceinline__ int n_pos(int t) const { return t ? n_pos1 : n_pos0; }
    __device__ __forceinline__ int n_flag(int t) const { return t ? n_flag1 : n_flag0; }
    __device__ __forceinline__ int n_ev(int t) const { return t ? n_ev1 : n_ev0; }
    __device__ __forceinline__ void set_n_pos(int t, int v) { if (t) n_pos1 = v; else n_pos0 = v; }
    __device__ __forceinline__ void set_n_flag(int t, int v) { if (t) n_flag1 = v; else n_flag0 = v; }
    __device__ __forceinline__ void set_n_ev(int t, int v) { if (t) n_ev1 = v; else n_ev0 = v; }
};

__host__ __device__ inline int hf_xcap(const HfParams &P) { return P.xi.mode ? P.cap_c : 0; }
// xcap: exciton slots (cap_c in the integrated modes, else 0)
__host__ __device__ inline size_t hf_warp_smem_bytes(int cap_c, int cap_f, int xcap = 0) {
    size_t b = 8u * (6u * cap_c + HF_CACHE) + 4u * (6u * cap_c + 2u * cap_f + 3u * HF_CACHE);
    if (xcap) b += 8u * ((size_t)xcap + 32u) + 4u * 3u * (((size_t)xcap + 1u) & ~(size_t)1u);
    return (b + 15u) & ~(size_t)15u;
}

__device__ __forceinline__ void hf_carve(HfWarp &w, unsigned char *base, int cap_c, int cap_f, int xcap = 0) {
    w.cap_c = cap_c; w.cap_f = cap_f; w.cache_stride = 1;
    double *d = reinterpret_cast<double *>(base);
    w.ev_tau_b = d; d += 2 * cap_c;
    w.inv_old_b = d; d += 2 * cap_c;
    w.term_b = d; d += 2 * cap_c;
    w.cache_tau = d; d += HF_CACHE;
    w.x_tau = w.x_scratch = nullptr; w.x_site = w.x_meta = w.x_final = nullptr;
    if (xcap) { w.x_tau = d; d += xcap; w.x_scratch = d; d += 32; }
    int32_t *i = reinterpret_cast<int32_t *>(d);
    if (xcap) { const int xc = (xcap + 1) & ~1; w.x_site = i; i += xc; w.x_meta = i; i += xc; w.x_final = i; i += xc; }
    w.pos_b = i; i += 2 * cap_c;
    w.ev_init_b = i; i += 2 * cap_c;
    w.ev_final_b = i; i += 2 * cap_c;
    w.flag_b = i; i += 2 * cap_f;
    w.cache_init = i; i += HF_CACHE;
    w.cache_final = i; i += HF_CACHE;
    w.cache_kp = i;
}

// ---- packed coordinates ---------------------------------------------------------------------
__device__ __forceinline__ int hf_pack(int x, int y, int z) { return x | (y << 10) | (z << 20); }
__device__ __forceinline__ int hf_px(int p) { return p & 1023; }
__device__ __forceinline__ int hf_py(int p) { return (p >> 10) & 1023; }
__device__ __forceinline__ int hf_pz(int p) { return p >> 20; }
__device__ __forceinline__ int64_t hf_site(const HfParams &P, int p) {
    return ((int64_t)hf_pz(p) * P.Y + hf_py(p)) * P.X + hf_px(p);
}
__device__ __forceinline__ int64_t hf_site(int X, int Y, int p) {
    return ((int64_t)hf_pz(p) * Y + hf_py(p)) * X + hf_px(p);
}
__device__ __forceinline__ int hf_pack_site(const HfParams &P, int64_t s) {
    const int x = (int)(s % P.X), y = (int)((s / P.X) % P.Y), z = (int)(s / ((int64_t)P.X * P.Y));
    return hf_pack(x, y, z);
}
// |a - b|^2 as the reference computes it: raw index differences, no minimum image (quirk Q7)
__device__ __forceinline__ int hf_dist2(int a, int b) {
    const int dx = hf_px(a) - hf_px(b), dy = hf_py(a) - hf_py(b), dz = hf_pz(a) - hf_pz(b);
    return dx * dx + dy * dy + dz * dz;
}

// ---- uniforms -------------------------------------------------------------------------------
// d-th uniform of the trajectory's TRAJ stream; sets *exhausted when a replay buffer is short.
__device__ __forceinline__ double hf_traj_uniform(const HfParams &P, const HfWarp &w, uint64_t d, bool *exhausted) {
    if (P.replay) {
        if (d >= (uint64_t)P.n_replay) { *exhausted = true; return 0.5; }
        return P.replay[w.local * P.n_replay + (int64_t)d];
    }
    return hf_block_uniform(hf_philox_block_rk(P.rk, HF_STREAM_TRAJ, w.ident, d >> 1), d);
}

// random.py:264-269 _randbelow_without_getrandbits (warp-uniform; every lane computes the same)
__device__ __forceinline__ int64_t hf_randbelow(const HfParams &P, HfWarp &w, int64_t n, bool *exhausted) {
    const int64_t maxsize = (int64_t)1 << 53;
    const int64_t rem = maxsize % n;
    const double limit = (double)(maxsize - rem) * 0x1p-53;          // exact
    double r = hf_traj_uniform(P, w, w.draws++, exhausted);
    while (r >= limit && !*exhausted) r = hf_traj_uniform(P, w, w.draws++, exhausted);
    return (int64_t)floor(r * 0x1p53) % n;
}

// ---- lanes per trajectory ---------------------------------------------------------------------
// The bookkeeping below is written once for L lanes cooperating on one trajectory: L = 32 is the
// warp-per-trajectory kernel (hf_run_kernel), L = 1 the thread-per-trajectory kernel for small `charges`
// (hf_frm_small.cuh), where every collective degenerates to the identity and every strided loop to a
// serial one.  Same statements, same order, same results.
template <int L> __device__ __forceinline__ unsigned hf_ballot(bool p) {
    if constexpr (L == 32) return __ballot_sync(HF_FULL, p); else return p ? 1u : 0u;
}
template <int L, typename T> __device__ __forceinline__ T hf_shfl(T v, int src) {
    if constexpr (L == 32) return __shfl_sync(HF_FULL, v, src); else return v;
}
template <int L, typename T> __device__ __forceinline__ T hf_shfl_xor(T v, int off) {
    if constexpr (L == 32) return __shfl_xor_sync(HF_FULL, v, off); else return v;
}
template <int L> __device__ __forceinline__ void hf_sync() {
    if constexpr (L == 32) __syncwarp();
}
// thread-per-trajectory candidate loop (hf_frm_small.cuh)
__device__ __forceinline__ int hf_gen_move_event_serial(const HfParams &P, HfWarp &w, int t, int pos, bool use_flags);

// ---- ordered-list primitives (warp-cooperative, n is warp-uniform) ---------------------------
template <int L = 32>
__device__ __forceinline__ int hf_find_first(const int32_t *a, int n, int v, int lane) {
#pragma unroll 1
    for (int base = 0; base < n; base += L) {
        const int i = base + lane;
        const unsigned m = hf_ballot<L>(i < n && a[i] == v);
        if (m) return base + __ffs(m) - 1;
    }
    return -1;
}

template <int L = 32>
__device__ __forceinline__ bool hf_contains(const int32_t *a, int n, int v, int lane) {
    return hf_find_first<L>(a, n, v, lane) >= 0;
}

// list.remove(a[idx]) keeping the order of the rest
template <int L = 32>
__device__ __forceinline__ void hf_remove_at(int32_t *a, int n, int idx, int lane) {
#pragma unroll 1
    for (int base = idx; base < n - 1; base += L) {
        const int i = base + lane;
        int v = 0;
        if (i < n - 1) v = a[i + 1];
        hf_sync<L>();
        if (i < n - 1) a[i] = v;
        hf_sync<L>();
    }
}

// Molecule.switch_electron / switch_hole (molecule.py:78-86) on the flag SET
template <int L = 32>
__device__ __forceinline__ bool hf_toggle_flag(const HfParams &P, HfWarp &w, int t, int site, int lane) {
    const int idx = hf_find_first<L>(w.flag(t), w.n_flag(t), site, lane);
    if (idx >= 0) {
        if (lane == 0) w.flag(t)[idx] = w.flag(t)[w.n_flag(t) - 1];
        w.set_n_flag(t, w.n_flag(t) - 1);
    } else {
        if (w.n_flag(t) >= P.cap_f) return false;
        if (lane == 0) w.flag(t)[w.n_flag(t)] = site;
        w.set_n_flag(t, w.n_flag(t) + 1);
    }
    hf_sync<L>();
    return true;
}

template <int L = 32>
__device__ __forceinline__ void hf_clear_flag(HfWarp &w, int t, int site, int lane) {
    const int idx = hf_find_first<L>(w.flag(t), w.n_flag(t), site, lane);
    if (idx >= 0) {
        if (lane == 0) w.flag(t)[idx] = w.flag(t)[w.n_flag(t) - 1];
        w.set_n_flag(t, w.n_flag(t) - 1);
    }
    hf_sync<L>();
}

// _remove_move_*_events (reseau.py:352-364) under Event.__eq__ (event.py:141-147): drop every
// event of type t whose initial == a OR final == b (quirk Q4), keeping the order of the rest.
template <int L = 32>
__device__ __forceinline__ void hf_remove_events(HfWarp &w, int t, int a, int b, int lane) {
    const int n = w.n_ev(t);
    int out = 0;
#pragma unroll 1
    for (int base = 0; base < n; base += L) {
        const int i = base + lane;
        int ini = 0, fin = 0;
        double tau = 0.0;
        bool keep = false;
        if (i < n) {
            ini = w.ev_init(t)[i]; fin = w.ev_final(t)[i]; tau = w.ev_tau(t)[i];
            keep = !(ini == a || fin == b);
        }
        const unsigned m = hf_ballot<L>(keep);
        const int dst = out + __popc(m & ((1u << lane) - 1u));
        hf_sync<L>();
        if (keep) { w.ev_init(t)[dst] = ini; w.ev_final(t)[dst] = fin; w.ev_tau(t)[dst] = tau; }
        hf_sync<L>();
        out += __popc(m);
    }
    w.set_n_ev(t, out);
}

// _update_events (reseau.py:564-578): tau -= time on every pending move event
template <int L = 32, bool kX = false>
__device__ __forceinline__ void hf_update_events(HfWarp &w, double time, int lane) {
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll 1
        for (int i = lane; i < w.n_ev(t); i += L) w.ev_tau(t)[i] = __dsub_rn(w.ev_tau(t)[i], time);
    if constexpr (kX) {                            // _move_exciton_events, _isc_events, _decay_events (reseau.py:569-576)
        w.clock = __dadd_rn(w.clock, time);
#pragma unroll 1
        for (int i = lane; i < w.n_x; i += L) w.x_tau[i] = __dsub_rn(w.x_tau[i], time);
    }
    hf_sync<L>();
}

// ---- neighbour table: reseau.py:184-205 for charge_tranfer_distance == 1 (quirk Q3) ----------
// i-th entry of _born_von_karman(p, 1, axe) for a periodic axis (always 3 entries)
__device__ __forceinline__ int hf_bvk_xy(int p, int size, int i) {
    if (p - 1 > -1 && p + 1 < size) return p - 1 + i;
    if (p - 1 < 0) return i < 2 ? i : size - 1;              // [0, 1, size-1]
    return i < 2 ? p - 1 + i : 0;                            // [size-2, size-1, 0]
}
// open axis z: 3 entries in the bulk, 2 on a face
__device__ __forceinline__ int hf_bvk_z(int p, int size, int i) {
    if (p - 1 > -1 && p + 1 < size) return p - 1 + i;
    if (p - 1 < 0) return i;                                 // [0, 1]
    return p - 1 + i;                                        // [size-2, size-1]
}

// ---- neighbour table for any charge_tranfer_distance: _born_von_karman verbatim (reseau.py:190-205) ----------
// The three branches as written, including what they do away from distance 1 (quirk Q3): a position within
// `distance` of the LOWER edge gets [0 .. d] + [size-d .. size-1] whatever the position is; one within `distance`
// of the UPPER edge gets [p-d .. size-1] + [0 .. d-1] (up to 3d entries); z is never wrapped.  Duplicated and
// asymmetric neighbours are part of the table: each entry is a candidate with its own uniform.
__device__ __forceinline__ int hf_bvk_count(int p, int size, int d, bool is_z) {
    if (p - d > -1 && p + d < size) return 2 * d + 1;
    if (p - d < 0) return is_z ? d + 1 : 2 * d + 1;
    return (size - (p - d)) + (is_z ? 0 : d);
}
__device__ __forceinline__ int hf_bvk_entry(int p, int size, int d, int i) {
    if (p - d > -1 && p + d < size) return p - d + i;
    if (p - d < 0) return i <= d ? i : size - d + (i - d - 1);
    const int first = size - (p - d);
    return i < first ? p - d + i : i - first;
}

// 1 / sqrt(d2), exp and -log(rng) / k as the reference computes them.  kCompact routes them through single
// out-of-line copies: the thread-per-trajectory kernel is bound by instruction fetch (ncu: stall_no_inst 60 %
// with every fp64 division, square root, exp and log expanded inline at each use), the warp kernel is not.
__device__ __noinline__ double hf_inv_dist_call(int d2) { return __ddiv_rn(1.0, __dsqrt_rn((double)d2)); }
__device__ __noinline__ double hf_boltzmann_call(double dE) { return exp(__ddiv_rn(-dE, HF_KT)); }
__device__ __noinline__ double hf_wait_call(double rng, double k) { return __ddiv_rn(-log(rng), k); }
template <bool kCompact> __device__ __forceinline__ double hf_inv_dist(int d2) {
    if constexpr (kCompact) return hf_inv_dist_call(d2); else return __ddiv_rn(1.0, __dsqrt_rn((double)d2));
}
// the same value from the table when there is one (warp kernel)
__device__ __forceinline__ double hf_inv_dist_tab(const HfParams &P, int d2) {
    return P.inv_dist ? __ldg(P.inv_dist + d2) : __ddiv_rn(1.0, __dsqrt_rn((double)d2));
}

// One candidate hop of a charge of type t from `initial` to `cand` given its uniform u:
// _time_move_electron / _time_move_hole (reseau.py:283-292, 315-324) with the Coulomb term of
// reseau.py:297-313 / 329-345.  inv_old[0][j] / inv_old[1][j] hold 1/|loc_j - initial| for the
// same-type / opposite-type lists (0 where the distance is 0).  Returns tau; *rate gets k;
// *zero_div is set where the reference would raise ZeroDivisionError.
template <bool kCompact = false>
__device__ __forceinline__ double hf_hop_time(const HfParams &P, const HfWarp &w, int t, int initial, int cand,
                                              double u, double *rate, bool *zero_div) {
    const double rng = __dsub_rn(1.0, u);                                        // 284 / 316
    const double *E = w.energy(t);
    double dE = __dsub_rn(__ldg(E + hf_site(P, cand)), __ldg(E + hf_site(P, initial)));   // 286 / 318
    const double fz = t == 0 ? -P.field : P.field;                               // (-1.*E) / (1.*E)
    dE = __dadd_rn(dE, __dmul_rn(fz, (double)(hf_pz(cand) - hf_pz(initial))));   // 287 / 319
    double out = 0.0;
    const int ns = w.n_pos(t), no = w.n_pos(1 - t);
    for (int j = 0; j < ns; ++j) {                                               // 299-304 / 331-336
        const int loc = w.pos(t)[j];
        if (loc == initial) continue;
        const int d2 = hf_dist2(loc, cand);
        if (d2 == 0) { *zero_div = true; return 0.0; }
        const double inv_new = hf_inv_dist<kCompact>(d2);
        out = __dadd_rn(out, __dmul_rn(HF_ELECTROSTATIC, __dsub_rn(inv_new, w.inv_old(0)[j])));
    }
    for (int j = 0; j < no; ++j) {                                               // 305-312 / 337-344
        const int loc = w.pos(1 - t)[j];
        if (loc == cand) { out = -INFINITY; break; }
        const double io = w.inv_old(1)[j];
        if (io == 0.0) { *zero_div = true; return 0.0; }
        const double inv_new = hf_inv_dist<kCompact>(hf_dist2(loc, cand));
        out = __dsub_rn(out, __dmul_rn(HF_ELECTROSTATIC, __dsub_rn(inv_new, io)));
    }
    dE = __dadd_rn(dE, out);                                                     // 288 / 320
    double k = HF_RATE0;                                                         // 289 / 321
    if (dE >= 0) k = __dmul_rn(k, kCompact ? hf_boltzmann_call(dE) : exp(__ddiv_rn(-dE, HF_KT)));   // 290-291 / 322-323
    *rate = k;
    if (k == 0.0) { *zero_div = true; return 0.0; }                              // float division by zero
    return kCompact ? hf_wait_call(rng, k) : __ddiv_rn(-log(rng), k);            // 292 / 324
}

// 1/|loc_j - initial| for both lists, computed once per regenerated event instead of once per
// candidate (the reference recomputes it 26 times: SURVEY.md §8a2); same bits.
template <int L = 32>
__device__ __forceinline__ void hf_fill_inv_old(const HfParams &P, HfWarp &w, int t, int initial, int lane) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int32_t *list = s == 0 ? w.pos(t) : w.pos(1 - t);
        const int n = s == 0 ? w.n_pos(t) : w.n_pos(1 - t);
        for (int j = lane; j < n; j += L) {
            const int d2 = hf_dist2(list[j], initial);
            if constexpr (L == 1) w.inv_old(s)[j] = d2 == 0 ? 0.0 : hf_inv_dist<true>(d2);
            else w.inv_old(s)[j] = d2 == 0 ? 0.0 : hf_inv_dist_tab(P, d2);
        }
    }
    hf_sync<L>();
}

// Bit mask of the candidates (ix, iy, iz), numbered (ix * 3 + iy) * nz + iz like the reference's neighbour list,
// whose coordinates equal the packed site q.  xs / ys / zs: the coordinates of the three neighbour columns,
// rows and planes, 10 bits each (1023 = no such plane).  More than one bit only on lattices so small that a
// periodic image repeats.
__device__ __noinline__ unsigned hf_cand_bits_hit(unsigned bx, unsigned by, unsigned bz, int nz) {
    unsigned m = 0;
#pragma unroll 1
    for (int ix = 0; ix < 3; ++ix) {
        if (!(bx >> ix & 1u)) continue;
#pragma unroll 1
        for (int iy = 0; iy < 3; ++iy) {
            if (!(by >> iy & 1u)) continue;
#pragma unroll 1
            for (int iz = 0; iz < nz; ++iz)
                if (bz >> iz & 1u) m |= 1u << ((ix * 3 + iy) * nz + iz);
        }
    }
    return m;
}
__device__ __forceinline__ unsigned hf_cand_bits(int xs, int ys, int zs, int nz, int q) {
    const int qx = hf_px(q), qy = hf_py(q), qz = hf_pz(q);
    const unsigned bx = (unsigned)((xs & 1023) == qx) | (unsigned)(((xs >> 10) & 1023) == qx) << 1 | (unsigned)(((xs >> 20) & 1023) == qx) << 2;
    if (!bx) return 0;                                                           // not in a neighbouring column: the usual case
    const unsigned by = (unsigned)((ys & 1023) == qy) | (unsigned)(((ys >> 10) & 1023) == qy) << 1 | (unsigned)(((ys >> 20) & 1023) == qy) << 2;
    const unsigned bz = (unsigned)((zs & 1023) == qz) | (unsigned)(((zs >> 10) & 1023) == qz) << 1 | (unsigned)(((zs >> 20) & 1023) == qz) << 2;
    if (!by || !bz) return 0;
    return hf_cand_bits_hit(bx, by, bz, nz);
}

// hf_hop_time for ONE candidate with the L lanes sharing the Coulomb sum: the terms
// C_e (1/|loc - cand| - 1/|loc - initial|) — a division and a square root each, 96 % of the reference's run
// time — are computed lane-parallel into shared memory, then added in list order exactly like reseau.py:299-312
// (same operands, same order, same bits).  Only valid when no ZeroDivisionError is geometrically possible
// (no same-type charge on the candidate, no opposite charge on `initial`): the caller checks.
template <int L>
__device__ __forceinline__ double hf_hop_time_coop(const HfParams &P, const HfWarp &w, int t, int initial, int cand,
                                                   double u, bool onto_opposite, int lane, bool *zero_div) {
    const double rng = __dsub_rn(1.0, u);                                        // 284 / 316
    const double *E = w.energy(t);
    double dE = __dsub_rn(__ldg(E + hf_site(P, cand)), __ldg(E + hf_site(P, initial)));   // 286 / 318
    const double fz = t == 0 ? -P.field : P.field;
    dE = __dadd_rn(dE, __dmul_rn(fz, (double)(hf_pz(cand) - hf_pz(initial))));   // 287 / 319
    const int ns = w.n_pos(t), no = w.n_pos(1 - t);
    for (int j = lane; j < ns; j += L) {
        const int loc = w.pos(t)[j];
        w.term(0)[j] = loc == initial ? 0.0 : __dmul_rn(HF_ELECTROSTATIC, __dsub_rn(hf_inv_dist_tab(P, hf_dist2(loc, cand)), w.inv_old(0)[j]));
    }
    for (int j = lane; j < no; j += L) {
        const int loc = w.pos(1 - t)[j];
        w.term(1)[j] = loc == cand ? 0.0 : __dmul_rn(HF_ELECTROSTATIC, __dsub_rn(hf_inv_dist_tab(P, hf_dist2(loc, cand)), w.inv_old(1)[j]));
    }
    hf_sync<L>();
    // the sum itself, in list order.  A skipped entry (the moving charge itself, reseau.py:300 / 332) holds +0.0,
    // and x + 0.0 == x bit for bit (out is never -0.0: it starts at +0.0 and round-to-nearest never produces -0.0
    // from a cancellation), so the loops carry no test; a candidate ON an opposite charge ends at -inf whatever
    // came before (reseau.py:305-308).
    double out = 0.0;
    if (onto_opposite) {
        out = -INFINITY;
    } else {
        // two terms per 16-byte load where the lists allow it (even capacity: both term arrays 16-byte aligned); the
        // additions stay one after the other, in list order
        const bool wide = (w.cap_c & 1) == 0;
        int j = 0;
        if (wide) {
            const double2 *t0 = reinterpret_cast<const double2 *>(w.term(0));
#pragma unroll 4
            for (; j + 1 < ns; j += 2) { const double2 v = t0[j >> 1]; out = __dadd_rn(__dadd_rn(out, v.x), v.y); }
        }
#pragma unroll 4
        for (; j < ns; ++j) out = __dadd_rn(out, w.term(0)[j]);                  // 299-304 / 331-336
        j = 0;
        if (wide) {
            const double2 *t1 = reinterpret_cast<const double2 *>(w.term(1));
#pragma unroll 4
            for (; j + 1 < no; j += 2) { const double2 v = t1[j >> 1]; out = __dsub_rn(__dsub_rn(out, v.x), v.y); }
        }
#pragma unroll 4
        for (; j < no; ++j) out = __dsub_rn(out, w.term(1)[j]);                  // 309-312 / 341-344
    }
    hf_sync<L>();
    dE = __dadd_rn(dE, out);                                                     // 288 / 320
    double k = HF_RATE0;
    if (dE >= 0) k = __dmul_rn(k, exp(__ddiv_rn(-dE, HF_KT)));                   // 290-291 / 322-323
    if (k == 0.0) { *zero_div = true; return 0.0; }                              // float division by zero
    return __ddiv_rn(-log(rng), k);                                              // 292 / 324
}

// hf_gen_move_event for charge_tranfer_distance > 1: the neighbour list has up to 3d x 3d x (2d+1) entries
// (reseau.py:184-205), taken 32 at a time in list order.  Every unblocked candidate is evaluated exactly (no
// screening: this table is off every benchmark); uniforms are numbered in list order across
// the chunks, the first minimum wins, the first failing candidate raises.
template <int L>
__device__ __forceinline__ int hf_gen_move_event_far(const HfParams &P, HfWarp &w, int t, int pos, bool use_flags, int lane) {
    const int px = hf_px(pos), py = hf_py(pos), pz = hf_pz(pos), d = P.dist;
    const int nx = hf_bvk_count(px, P.X, d, false), ny = hf_bvk_count(py, P.Y, d, false), nz = hf_bvk_count(pz, P.Z, d, true);
    const int total = nx * ny * nz;
    const int32_t *blk = use_flags ? w.flag(t) : w.pos(t);
    const int nb = use_flags ? w.n_flag(t) : w.n_pos(t);
    hf_fill_inv_old<L>(P, w, t, pos, lane);
    double btau = INFINITY;
    int bfin = -1;
    uint64_t drawn = 0;
    for (int base = 0; base < total; base += L) {
        const int i = base + lane;
        int cand = pos;
        bool active = false;
        if (i < total) {
            cand = hf_pack(hf_bvk_entry(px, P.X, d, i / (ny * nz)), hf_bvk_entry(py, P.Y, d, (i / nz) % ny), hf_bvk_entry(pz, P.Z, d, i % nz));
            active = cand != pos;                                                // reseau.py:188
            for (int j = 0; j < nb && active; ++j) active = blk[j] != cand;      // 385 / 394 ; 249 / 268
        }
        const unsigned mask = hf_ballot<L>(active);
        if (mask == 0) continue;
        bool exhausted = false, zero_div = false;
        double tau = INFINITY, rate = 0.0;
        if (active) {
            const double u = hf_traj_uniform(P, w, w.draws + drawn + (uint64_t)__popc(mask & ((1u << lane) - 1u)), &exhausted);
            if (!exhausted) tau = hf_hop_time(P, w, t, pos, cand, u, &rate, &zero_div);
        }
        if (hf_ballot<L>(exhausted)) return HF_ST_REPLAY_EXHAUSTED;
        const unsigned zmask = hf_ballot<L>(active && zero_div);
        if (zmask) {                               // draws up to and including the failing candidate
            const int f = __ffs(zmask) - 1;
            w.draws += drawn + (uint64_t)__popc(mask & (f == 31 ? HF_FULL : ((2u << f) - 1u)));
            return HF_ST_ZERO_DIVISION;
        }
        int best = active ? lane : 99;
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) {
            const double otau = hf_shfl_xor<L>(tau, off);
            const int obest = hf_shfl_xor<L>(best, off);
            if (otau < tau || (otau == tau && obest < best)) { tau = otau; best = obest; }
        }
        const int fin = hf_shfl<L>(cand, best & 31);
        if (bfin < 0 || tau < btau) { btau = tau; bfin = fin; }                  // first minimum across the chunks
        drawn += (uint64_t)__popc(mask);
    }
    if (bfin < 0) return HF_ST_VALUE_ERROR;                                      // min() of an empty sequence
    w.draws += drawn;
    if (w.n_ev(t) >= P.cap_c) return HF_ST_CAPACITY;
    if (lane == 0) { w.ev_init(t)[w.n_ev(t)] = pos; w.ev_final(t)[w.n_ev(t)] = bfin; w.ev_tau(t)[w.n_ev(t)] = btau; }
    w.set_n_ev(t, w.n_ev(t) + 1);
    hf_sync<L>();
    return HF_ST_OK;
}

// _new_move_electron_events / _new_move_hole_events (reseau.py:380-396; use_flags = true) and the
// per-charge body of _init_move_*_events (reseau.py:240-276; use_flags = false: membership in
// the position list).  Lanes = candidates in the reference's x-major neighbour order; one TRAJ
// uniform per unblocked candidate in that order; first minimum wins.  Returns a status.
template <int L = 32>
__device__ __forceinline__ int hf_gen_move_event(const HfParams &P, HfWarp &w, int t, int pos, bool use_flags, int lane) {
    if constexpr (L == 1) return hf_gen_move_event_serial(P, w, t, pos, use_flags);
    if (P.dist != 1) return hf_gen_move_event_far<L>(P, w, t, pos, use_flags, lane);
    const int px = hf_px(pos), py = hf_py(pos), pz = hf_pz(pos);
    const int nz = (pz - 1 > -1 && pz + 1 < P.Z) ? 3 : 2;
    const int total = 9 * nz;
    const int ix = lane / (3 * nz), iy = (lane / nz) % 3, iz = lane % nz;
    const int cx = hf_bvk_xy(px, P.X, ix < 3 ? ix : 0), cy = hf_bvk_xy(py, P.Y, iy), cz = hf_bvk_z(pz, P.Z, iz);
    const int cand = hf_pack(cx, cy, cz);
    const int32_t *blk = use_flags ? w.flag(t) : w.pos(t);
    const int nb = use_flags ? w.n_flag(t) : w.n_pos(t);
    const int ns = w.n_pos(t), no = w.n_pos(1 - t);
    // which candidates are blocked (385 / 394 ; 249 / 268), lead onto an opposite charge, or onto a same-type
    // charge (flag / list desynchronisation: ZeroDivisionError in the reference): from the lists, lane-strided
    unsigned blocked_m = 0, opp_m = 0, same_m = 0;
    bool opp_on_pos = false;
    {
        int xs = 0, ys = 0, zs = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xs |= hf_bvk_xy(px, P.X, i) << (10 * i); ys |= hf_bvk_xy(py, P.Y, i) << (10 * i);
            zs |= (i < nz ? hf_bvk_z(pz, P.Z, i) : 1023) << (10 * i);
        }
        for (int j = lane; j < nb; j += L) blocked_m |= hf_cand_bits(xs, ys, zs, nz, blk[j]);
        for (int j = lane; j < no; j += L) { const int q = w.pos(1 - t)[j]; opp_m |= hf_cand_bits(xs, ys, zs, nz, q); opp_on_pos |= q == pos; }
        for (int j = lane; j < ns; j += L) { const int q = w.pos(t)[j]; if (q != pos) same_m |= hf_cand_bits(xs, ys, zs, nz, q); }
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) {
            blocked_m |= hf_shfl_xor<L>(blocked_m, off); opp_m |= hf_shfl_xor<L>(opp_m, off); same_m |= hf_shfl_xor<L>(same_m, off);
        }
    }
    const bool active = lane < total && cand != pos && !(blocked_m >> lane & 1u);
    const unsigned mask = hf_ballot<L>(active);
    if (mask == 0) return HF_ST_VALUE_ERROR;                                     // min() of an empty sequence
    hf_fill_inv_old<L>(P, w, t, pos, lane);
    bool exhausted = false;
    const int rank = __popc(mask & ((1u << lane) - 1u));
    double u = 0.5;
    if (active) u = hf_traj_uniform(P, w, w.draws + (uint64_t)rank, &exhausted);
    if (hf_ballot<L>(exhausted)) return HF_ST_REPLAY_EXHAUSTED;
    double tau = INFINITY;
    int best = 99;
    if (hf_ballot<L>(opp_on_pos) != 0 || (same_m & mask) != 0) {
        // a ZeroDivisionError is geometrically possible: every candidate exactly, the first failing one decides
        double rate = 0.0;
        bool zero_div = false;
        if (active) tau = hf_hop_time(P, w, t, pos, cand, u, &rate, &zero_div);
        const unsigned zmask = hf_ballot<L>(active && zero_div);
        if (zmask) {                               // draws up to and including the failing candidate
            const int f = __ffs(zmask) - 1;
            w.draws += (uint64_t)__popc(mask & (f == 31 ? HF_FULL : ((2u << f) - 1u)));
            return HF_ST_ZERO_DIVISION;
        }
        // min(events): first minimum in candidate order (387 / 396)
        best = active ? lane : 99;
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) {
            const double otau = hf_shfl_xor<L>(tau, off);
            const int obest = hf_shfl_xor<L>(best, off);
            if (otau < tau || (otau == tau && obest < best)) { tau = otau; best = obest; }
        }
    } else {
        // Screen-then-verify (cf. hf_c1_gen): an fp32 estimate of 1e13 * tau without the Coulomb sum for every
        // candidate (lanes), exact evaluation — the lanes then share the O(charges) Coulomb sum — only of the
        // candidates whose estimate lies within 1 + W of the smallest.  The estimate's relative error is
        // <= eps = 4e-5 (hf_c1_gen) + C_e (ns + no) / kT (every neglected term is |1/r' - 1/r| <= 1); a candidate
        // outside W = 4 eps (>= 2^-9) has tau >= a (1 - eps) > a_min (1 + W)(1 - eps) > a_min (1 + eps) >= tau of
        // the best estimate: it can neither win nor tie.
        float a = INFINITY;
        bool forced = false;
        if (active) {
            const double *E = w.energy(t);
            const double fz = t == 0 ? -P.field : P.field;
            const double dE = (__ldg(E + hf_site(P, cand)) - __ldg(E + hf_site(P, pos))) + fz * (double)(cz - pz);
            const float x = (opp_m >> lane & 1u) ? -INFINITY : (float)dE * (float)(1.0 / HF_KT);
            float t0;
            if (u < 0x1p-6) {
                const float uf = (float)u;
                t0 = uf * (1.0f + uf * (0.5f + uf * (1.0f / 3.0f)));
            } else {
                t0 = -__logf((float)(1.0 - u));
            }
            a = t0;
            if (x > 0.0f) a = t0 * __expf(x);
            if (x > 700.0f) forced = true;                                       // the rate may underflow: ZeroDivisionError
            if (t0 == 0.0f) a = 0.0f;                                            // 0 * inf
        }
        float amin = a;
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) amin = fminf(amin, hf_shfl_xor<L>(amin, off));
        const float eps = 4e-5f + (float)(HF_ELECTROSTATIC / HF_KT) * (float)(ns + no);
        const float window = fmaxf(0x1p-9f, 4.0f * eps);
        const float thr = amin * (1.0f + window);
        const unsigned verify = hf_ballot<L>(active && (forced || !(a > thr)));
        for (unsigned m = verify; m; m &= m - 1) {
            const int i = __ffs(m) - 1;
            const int cand_i = hf_shfl<L>(cand, i);
            const double u_i = hf_shfl<L>(u, i);
            bool zero_div = false;
            const double tau_i = hf_hop_time_coop<L>(P, w, t, pos, cand_i, u_i, (opp_m >> i & 1u) != 0, lane, &zero_div);
            if (zero_div) {                                                      // draws up to and including the failing candidate
                w.draws += (uint64_t)__popc(mask & (i == 31 ? HF_FULL : ((2u << i) - 1u)));
                return HF_ST_ZERO_DIVISION;
            }
            if (best == 99 || tau_i < tau) { tau = tau_i; best = i; }            // first minimum in candidate order
        }
    }
    w.draws += (uint64_t)__popc(mask);
    const int fin = hf_shfl<L>(cand, best & 31);
    if (w.n_ev(t) >= P.cap_c) return HF_ST_CAPACITY;
    if (lane == 0) { w.ev_init(t)[w.n_ev(t)] = pos; w.ev_final(t)[w.n_ev(t)] = fin; w.ev_tau(t)[w.n_ev(t)] = tau; }
    w.set_n_ev(t, w.n_ev(t) + 1);
    hf_sync<L>();
    return HF_ST_OK;
}

// _electron_reinjection / _hole_reinjection (reseau.py:448-470): choice() over the free sites of
// the injection face in x-major order, append, toggle the flag, _injection += 1.
template <int L = 32>
__device__ __forceinline__ int hf_reinject(const HfParams &P, HfWarp &w, int t, int lane) {
    const int zf = t == 0 ? P.Z - 1 : 0;
    const int n = w.n_pos(t);
    const int32_t *list = w.pos(t);
    // distinct occupied face sites
    int n_occ = 0;
    for (int base = 0; base < n; base += L) {
        const int j = base + lane;
        bool first = false;
        if (j < n && hf_pz(list[j]) == zf) {
            first = true;
            for (int q = 0; q < j; ++q) first &= (list[q] != list[j]);
        }
        n_occ += __popc(hf_ballot<L>(first));
    }
    const int64_t n_free = (int64_t)P.X * P.Y - n_occ;
    if (n_free <= 0) return HF_ST_VALUE_ERROR;                                   // choice() of an empty sequence
    bool exhausted = false;
    const int64_t k = hf_randbelow(P, w, n_free, &exhausted);                    // random.py:345-350
    if (exhausted) return HF_ST_REPLAY_EXHAUSTED;
    // k-th free face index f = x*Y + y: fixed point of f = k + #{distinct occupied <= f}
    int64_t f = k;
    for (;;) {
        int c = 0;
        for (int base = 0; base < n; base += L) {
            const int j = base + lane;
            bool hit = false;
            if (j < n && hf_pz(list[j]) == zf && (int64_t)hf_px(list[j]) * P.Y + hf_py(list[j]) <= f) {
                hit = true;
                for (int q = 0; q < j; ++q) hit &= (list[q] != list[j]);
            }
            c += __popc(hf_ballot<L>(hit));
        }
        const int64_t f2 = k + c;
        if (f2 == f) break;
        f = f2;
    }
    const int site = hf_pack((int)(f / P.Y), (int)(f % P.Y), zf);
    if (n >= P.cap_c) return HF_ST_CAPACITY;
    if (lane == 0) w.pos(t)[n] = site;
    w.set_n_pos(t, n + 1);
    hf_sync<L>();
    if (!hf_toggle_flag<L>(P, w, t, site, lane)) return HF_ST_CAPACITY;
    w.injection += 1;
    return HF_ST_OK;
}

// ---- state load / store ------------------------------------------------------------------------
// kX: the integrated exciton modes are compiled in (P.xi.mode != 0); the kernels of the reference mode carry none of it
template <int L = 32, bool kX = false>
__device__ __forceinline__ void hf_load_state(const HfParams &P, HfWarp &w, int64_t local, int lane) {
    const HfScalars sc = P.sc[local];
    w.local = local;
    w.ident = P.first_traj + (uint32_t)local;
#pragma unroll
    for (int t = 0; t < 2; ++t) { w.set_n_pos(t, sc.n_pos[t]); w.set_n_flag(t, sc.n_flag[t]); w.set_n_ev(t, sc.n_ev[t]); }
    w.special = sc.special; w.special_site = sc.special_site; w.xstate = sc.xstate; w.status = sc.status;
    w.cache_head = sc.cache_head; w.cache_n = sc.cache_n;
    w.draws = sc.draws; w.spins = sc.spins; w.emission = sc.emission; w.recombination = sc.recombination;
    w.injection = sc.injection; w.steps = sc.steps; w.fired = sc.fired;
    const int64_t r = local / P.traj_per_real;
    w.species = P.species + r * P.N;
    w.lumo = P.lumo + r * P.N;
    w.homo = P.homo + r * P.N;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll 1
        for (int i = lane; i < w.n_pos(t); i += L) w.pos(t)[i] = P.pos[t][local * P.cap_c + i];
#pragma unroll 1
        for (int i = lane; i < w.n_flag(t); i += L) w.flag(t)[i] = P.flag[t][local * P.cap_f + i];
#pragma unroll 1
        for (int i = lane; i < w.n_ev(t); i += L) {
            w.ev_init(t)[i] = P.ev_init[t][local * P.cap_c + i];
            w.ev_final(t)[i] = P.ev_final[t][local * P.cap_c + i];
            w.ev_tau(t)[i] = P.ev_tau[t][local * P.cap_c + i];
        }
    }
    if constexpr (L == 1) {                        // the ring is used in place: [slot][trajectory] in global memory
        w.cache_stride = P.B;
        w.cache_init = P.cache_init + local; w.cache_final = P.cache_final + local;
        w.cache_kp = P.cache_kp + local; w.cache_tau = P.cache_tau + local;
    } else
    for (int c = lane; c < HF_CACHE; c += L) {   // ring stored slot-major: [slot][trajectory]
        w.cache_init[c] = P.cache_init[c * P.B + local];
        w.cache_final[c] = P.cache_final[c * P.B + local];
        w.cache_kp[c] = P.cache_kp[c * P.B + local];
        w.cache_tau[c] = P.cache_tau[c * P.B + local];
    }
    w.c_move_e = w.c_move_h = w.c_bound = w.c_decay = w.c_capture_e = w.c_capture_h = 0;
    w.c_recomb_host = w.c_recomb_tadf = w.c_recomb_fluo = w.c_emit_tadf = w.c_emit_fluo = 0;
    w.c_x_forster = w.c_x_dexter = w.c_x_isc = w.c_x_risc = 0;
    w.n_x = 0; w.xdraws = 0; w.clock = 0.0;
    if constexpr (kX) {
        w.n_x = P.xi.x_n[local]; w.xdraws = P.xi.x_draws[local]; w.clock = P.xi.x_clock[local];
#pragma unroll 1
        for (int i = lane; i < w.n_x; i += L) {
            w.x_site[i] = P.xi.x_site[local * P.cap_c + i]; w.x_meta[i] = P.xi.x_meta[local * P.cap_c + i];
            w.x_final[i] = P.xi.x_final[local * P.cap_c + i]; w.x_tau[i] = P.xi.x_tau[local * P.cap_c + i];
        }
    }
    hf_sync<L>();
}

template <int L = 32, bool kX = false>
__device__ __forceinline__ void hf_store_state(const HfParams &P, HfWarp &w, int lane) {
    const int64_t local = w.local;
    hf_sync<L>();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll 1
        for (int i = lane; i < w.n_pos(t); i += L) P.pos[t][local * P.cap_c + i] = w.pos(t)[i];
#pragma unroll 1
        for (int i = lane; i < w.n_flag(t); i += L) P.flag[t][local * P.cap_f + i] = w.flag(t)[i];
#pragma unroll 1
        for (int i = lane; i < w.n_ev(t); i += L) {
            P.ev_init[t][local * P.cap_c + i] = w.ev_init(t)[i];
            P.ev_final[t][local * P.cap_c + i] = w.ev_final(t)[i];
            P.ev_tau[t][local * P.cap_c + i] = w.ev_tau(t)[i];
        }
    }
    if constexpr (L == 32)
    for (int c = lane; c < HF_CACHE; c += L) {
        P.cache_init[c * P.B + local] = w.cache_init[c];
        P.cache_final[c * P.B + local] = w.cache_final[c];
        P.cache_kp[c * P.B + local] = w.cache_kp[c];
        P.cache_tau[c * P.B + local] = w.cache_tau[c];
    }
    if constexpr (kX) {
#pragma unroll 1
        for (int i = lane; i < w.n_x; i += L) {
            P.xi.x_site[local * P.cap_c + i] = w.x_site[i]; P.xi.x_meta[local * P.cap_c + i] = w.x_meta[i];
            P.xi.x_final[local * P.cap_c + i] = w.x_final[i]; P.xi.x_tau[local * P.cap_c + i] = w.x_tau[i];
        }
        if (lane == 0) {
            P.xi.x_n[local] = w.n_x; P.xi.x_draws[local] = w.xdraws; P.xi.x_clock[local] = w.clock;
            const unsigned cx[4] = {w.c_x_forster, w.c_x_dexter, w.c_x_isc, w.c_x_risc};
#pragma unroll
            for (int i = 0; i < 4; ++i) if (cx[i]) atomicAdd(P.totals + 17 + i, (unsigned long long)cx[i]);
        }
    }
    if (lane == 0) {
        HfScalars sc;
#pragma unroll
        for (int t = 0; t < 2; ++t) { sc.n_pos[t] = w.n_pos(t); sc.n_flag[t] = w.n_flag(t); sc.n_ev[t] = w.n_ev(t); }
        sc.special = w.special; sc.special_site = w.special_site; sc.xstate = w.xstate; sc.status = w.status;
        sc.cache_head = w.cache_head; sc.cache_n = w.cache_n;
        sc.draws = w.draws; sc.spins = w.spins; sc.emission = w.emission; sc.recombination = w.recombination;
        sc.injection = w.injection; sc.steps = w.steps; sc.fired = w.fired;
        P.sc[local] = sc;
        // channel counters: one atomic per non-zero counter per trajectory per launch
        const unsigned c[11] = {w.c_move_e, w.c_move_h, w.c_bound, w.c_decay, w.c_capture_e, w.c_capture_h,
                                w.c_recomb_host, w.c_recomb_tadf, w.c_recomb_fluo, w.c_emit_tadf, w.c_emit_fluo};
#pragma unroll
        for (int i = 0; i < 11; ++i) if (c[i]) atomicAdd(P.totals + 5 + i, (unsigned long long)c[i]);
    }
}

#include "hf_frm_x.cuh"

// Re-inject (optionally) and regenerate the move event of the LAST charge of type t: queued (4 bits per request:
// 8 | 4 * reinject | t, at most two per event) and run in ONE place after the bookkeeping — by hf_step itself in the
// warp kernel, by the kernel's common tail in the thread-per-trajectory kernel (so that the lanes of a warp, each
// with its own trajectory, execute the candidate loop together).  One instance of the candidate loop per kernel
// instead of one per branch of the dispatcher.
template <int L>
__device__ __forceinline__ int hf_regen(const HfParams &P, HfWarp &w, int t, bool reinject, int lane, int *queue) {
    *queue |= (8 | (reinject ? 4 : 0) | t) << (*queue ? 4 : 0);
    return HF_ST_OK;
}

// ---- one KMC step: _first_reaction_method (reseau.py:483-562) ---------------------------------
// Returns HF_ST_OK when an event fired; ev_* describe it (packed coordinates).
template <int L, bool kX>
__device__ __forceinline__ int hf_step_core(const HfParams &P, HfWarp &w, int lane, int *ev_kind, int *ev_part,
                                            int *ev_init, int *ev_final, double *ev_tau, int *queue, int *xnew);

// xnew: the exciton slot (integrated modes) that needs a new pending event after this step; it is drawn in ONE
// place so that the generator is instantiated once per kernel.
template <int L = 32, bool kX = false>
__device__ __forceinline__ int hf_step(const HfParams &P, HfWarp &w, int lane, int *ev_kind, int *ev_part,
                                       int *ev_init, int *ev_final, double *ev_tau, int *queue = nullptr) {
    int xnew = -1, own_queue = 0;
    int st = hf_step_core<L, kX>(P, w, lane, ev_kind, ev_part, ev_init, ev_final, ev_tau, queue ? queue : &own_queue, &xnew);
    for (; st == HF_ST_OK && own_queue; own_queue >>= 4) {                       // the regenerations the step asked for
        const int t = own_queue & 1;
        if (own_queue & 4) st = hf_reinject<L>(P, w, t, lane);
        if (st == HF_ST_OK) st = hf_gen_move_event<L>(P, w, t, w.pos(t)[w.n_pos(t) - 1], true, lane);
    }
    if constexpr (kX) { if (st == HF_ST_OK && xnew >= 0) hfi_new_event<L>(P, w, xnew, lane); }
    return st;
}

template <int L, bool kX>
__device__ __forceinline__ int hf_step_core(const HfParams &P, HfWarp &w, int lane, int *ev_kind, int *ev_part,
                                            int *ev_init, int *ev_final, double *ev_tau, int *queue, int *xnew) {
    int kind, part, ini, fin;
    double tau;
    if (w.special) {
        // a tau = 0 bound / decay / capture event sits later in the concatenation order than any
        // move event and all residual taus are >= 0, so it is the pop() of reseau.py:486-487 (Q6)
        kind = w.special & 0xff; part = w.special >> 8; ini = fin = w.special_site; tau = 0.0;
    } else {
        // min tau over move_e + move_h, LAST one in concatenation order on ties (Q6)
        const int ne = w.n_ev(0), total = ne + w.n_ev(1);
        if (total == 0 && (!kX || w.n_x == 0)) return HF_ST_NO_EVENTS;           // 487-488
        double btau = INFINITY;
        int bidx = -1;
        for (int base = 0; base < total; base += L) {
            const int i = base + lane;
            if (i < total) {
                const double ti = i < ne ? w.ev_tau(0)[i] : w.ev_tau(1)[i - ne];
                if (bidx < 0 || ti <= btau) { btau = ti; bidx = i; }
            }
        }
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) {
            const double ot = hf_shfl_xor<L>(btau, off);
            const int oi = hf_shfl_xor<L>(bidx, off);
            if (oi >= 0 && (bidx < 0 || ot < btau || (ot == btau && oi > bidx))) { btau = ot; bidx = oi; }
        }
        // the excitons' events follow the carriers' in _get_all_events (reseau.py:608-612): move_x, isc, decay lists,
        // slots in creation order inside each; equal taus -> the later one wins
        int xsel = -1;
        if (kX && w.n_x > 0) {
            double xt = INFINITY;
            int xkey = -1;
#pragma unroll 1
            for (int i = lane; i < w.n_x; i += L) {
                const double ti = w.x_tau[i];
                const int key = (hfi_class(HFI_KIND(w.x_meta[i])) << 16) | i;
                if (xkey < 0 || ti < xt || (ti == xt && key > xkey)) { xt = ti; xkey = key; }
            }
#pragma unroll
            for (int off = L / 2; off > 0; off >>= 1) {
                const double ot = hf_shfl_xor<L>(xt, off);
                const int ok = hf_shfl_xor<L>(xkey, off);
                if (ok >= 0 && (xkey < 0 || ot < xt || (ot == xt && ok > xkey))) { xt = ot; xkey = ok; }
            }
            if (bidx < 0 || xt <= btau) { xsel = xkey & 0xffff; btau = xt; }
        }
        if (kX && xsel >= 0) {
            kind = HFI_KIND(w.x_meta[xsel]); part = HF_PART_EXCITON; tau = btau;
            ini = w.x_site[xsel]; fin = w.x_final[xsel];
            hf_sync<L>();
            if (lane == 0) {
                const int64_t at = w.cache_head * w.cache_stride;
                w.cache_init[at] = ini; w.cache_final[at] = fin;
                w.cache_kp[at] = kind | (part << 8); w.cache_tau[at] = tau;
            }
            w.cache_head = w.cache_head + 1 == HF_CACHE ? 0 : w.cache_head + 1;
            if (w.cache_n < HF_CACHE) w.cache_n += 1;
            *ev_kind = kind; *ev_part = part; *ev_init = ini; *ev_final = fin; *ev_tau = tau;
            if constexpr (kX) return hfi_fire<L>(P, w, xsel, kind, fin, tau, lane, queue, xnew);
        }
        const int t = bidx < ne ? 0 : 1, i = bidx - (t ? ne : 0);
        kind = HF_K_MOVE; part = t + 1; tau = btau;
        ini = w.ev_init(t)[i]; fin = w.ev_final(t)[i];
    }
    // _cache.popleft(); _cache.append(event)  (489-490)
    if (lane == 0) {
        const int64_t at = w.cache_head * w.cache_stride;
        w.cache_init[at] = ini; w.cache_final[at] = fin;
        w.cache_kp[at] = kind | (part << 8); w.cache_tau[at] = tau;
    }
    w.cache_head = w.cache_head + 1 == HF_CACHE ? 0 : w.cache_head + 1;
    if (w.cache_n < HF_CACHE) w.cache_n += 1;
    *ev_kind = kind; *ev_part = part; *ev_init = ini; *ev_final = fin; *ev_tau = tau;

    if (kind == HF_K_MOVE) {                                                     // 492-524
        const int t = part - 1, o = 1 - t;
        hf_remove_events<L>(w, t, ini, fin, lane);                                  // 495 / 511 (Q4)
        if (!hf_toggle_flag<L>(P, w, t, ini, lane)) return HF_ST_CAPACITY;          // 423-424 / 429-430
        if (!hf_toggle_flag<L>(P, w, t, fin, lane)) return HF_ST_CAPACITY;
        const int idx = hf_find_first<L>(w.pos(t), w.n_pos(t), ini, lane);          // 425-426 / 431-432
        if (idx < 0) return HF_ST_VALUE_ERROR;
        hf_remove_at<L>(w.pos(t), w.n_pos(t), idx, lane);
        if (lane == 0) w.pos(t)[w.n_pos(t) - 1] = fin;
        hf_sync<L>();
        const bool opposite_here = hf_contains<L>(w.flag(o), w.n_flag(o), fin, lane);
        const int capture_z = t == 0 ? 0 : P.Z - 1;
        if (!opposite_here && hf_pz(fin) != capture_z) {                         // 498-500 / 514-516
            hf_update_events<L, kX>(w, tau, lane);
            return hf_regen<L>(P, w, t, false, lane, queue);                    // the moved charge is the last of its list
        } else if (opposite_here) {                                              // 501-505 / 517-521
            hf_remove_events<L>(w, o, fin, fin, lane);
            // _update_events(0.) leaves every tau unchanged (Q5)
            w.special = HF_K_BOUND | (HF_PART_EXCITON << 8); w.special_site = fin;
        } else {                                                                 // 506-508 / 522-524
            hf_update_events<L, kX>(w, tau, lane);
            w.special = HF_K_CAPTURE | (part << 8); w.special_site = fin;
        }
        return HF_ST_OK;
    }
    const int site = w.special_site;
    w.special = 0;                                                               // _remove_*_event
    if (kind == HF_K_BOUND) {                                                    // 529-533
        const bool both = hf_contains<L>(w.flag(0), w.n_flag(0), site, lane) && hf_contains<L>(w.flag(1), w.n_flag(1), site, lane);
        if (both) {                                                              // molecule.py:88-96
            const double r = hf_philox_uniform(P.k0, P.k1, HF_STREAM_SPIN, w.ident, w.spins);
            w.spins += 1;
            w.xstate = (0 <= r && r < 0.25) ? HF_XS_SINGLET : HF_XS_TRIPLET;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {                                            // 434-438
            const int idx = hf_find_first<L>(w.pos(t), w.n_pos(t), site, lane);
            if (idx < 0) return HF_ST_VALUE_ERROR;
            hf_remove_at<L>(w.pos(t), w.n_pos(t), idx, lane);
            w.set_n_pos(t, w.n_pos(t) - 1);
        }
        if constexpr (kX) {                    // integrated: the exciton gets a slot and its first event (hf_frm_x.cuh)
            if (w.n_x >= P.cap_c) return HF_ST_CAPACITY;
            const int k = w.n_x;
            hf_sync<L>();
            if (lane == 0) { w.x_site[k] = site; w.x_meta[k] = both ? (w.xstate | HFI_ONSITE) : 0; }
            w.n_x = k + 1;
            w.xstate = 0;
            hf_sync<L>();
            *xnew = k;
            return HF_ST_OK;
        }
        w.special = HF_K_DECAY | (HF_PART_EXCITON << 8); w.special_site = site;
        return HF_ST_OK;
    }
    if (kind == HF_K_DECAY) {                                                    // 541-547
        const int sp = w.species[hf_site(P, site)];
        bool photon = false;
        if (w.xstate) {                                                          // molecule.py:185-199, 283-297, 387-400
            photon = sp != HF_SPECIES_HOST && w.xstate == HF_XS_SINGLET;
            w.xstate = 0;
            hf_clear_flag<L>(w, 0, site, lane);
            hf_clear_flag<L>(w, 1, site, lane);
        }
        w.recombination += 1;                                                    // 472-476
        w.c_recomb_host += sp == 0; w.c_recomb_tadf += sp == 1; w.c_recomb_fluo += sp == 2;
        if (photon) { w.emission += 1; w.c_emit_tadf += sp == 1; w.c_emit_fluo += sp == 2; }
        const int st = hf_regen<L>(P, w, 0, true, lane, queue);
        if (st != HF_ST_OK) return st;
        return hf_regen<L>(P, w, 1, true, lane, queue);
    }
    // capture (551-561)
    const int t = part - 1;
    if (!hf_toggle_flag<L>(P, w, t, site, lane)) return HF_ST_CAPACITY;             // 440-446
    const int idx = hf_find_first<L>(w.pos(t), w.n_pos(t), site, lane);
    if (idx < 0) return HF_ST_VALUE_ERROR;
    hf_remove_at<L>(w.pos(t), w.n_pos(t), idx, lane);
    w.set_n_pos(t, w.n_pos(t) - 1);
    return hf_regen<L>(P, w, t, true, lane, queue);
}

// ---- kernels -------------------------------------------------------------------------------------
// Lattice.operations(stop) (reseau.py:580-595) for every trajectory: one warp per trajectory,
// grid-stride over trajectories.
// min blocks: 512 threads per SM, i.e. at most 128 registers (what the kernel needed before the exciton generator
// became one of its callees; the generator's own needs must not halve the occupancy of the carrier path)
template <int kWarps, bool kX = false>
__global__ void __launch_bounds__(kWarps * 32, 16 / kWarps) hf_run_kernel(const HfParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HfWarp w;
    hf_carve(w, hf_smem + (size_t)warp * hf_warp_smem_bytes(P.cap_c, P.cap_f, hf_xcap(P)), P.cap_c, P.cap_f, hf_xcap(P));
    for (int64_t local = (int64_t)blockIdx.x * kWarps + warp; local < P.B; local += (int64_t)gridDim.x * kWarps) {
        hf_load_state<32, kX>(P, w, local, lane);
        // a ZeroDivisionError only ends the operations() call it happened in (reseau.py:587-589)
        if (w.status == HF_ST_ZERO_DIVISION || w.status == HF_ST_NO_EVENTS) w.status = HF_ST_OK;
        const bool traced = P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
        HfTraceRec *trace = traced ? P.trace + (local - P.trace_first) * P.trace_cap : nullptr;
        const uint64_t trace_base = traced ? P.trace_base[local - P.trace_first] : 0;   // events fired before the trace was armed
        if (w.status == HF_ST_OK) {
            for (int64_t i = 0; i < P.n_steps; ++i) {
                w.steps += 1;                                                    // 582
                int kind, part, ini, fin;
                double tau;
                const int st = hf_step<32, kX>(P, w, lane, &kind, &part, &ini, &fin, &tau);
                if (st != HF_ST_OK) { w.status = st; break; }                    // 587-591
                if (kind == HF_K_BOUND) w.c_bound += 1;
                else if (kind == HF_K_DECAY) w.c_decay += 1;
                else if (part == HF_PART_EXCITON) {}                             // transfers and spin flips: counted by hfi_fire
                else if (kind == HF_K_MOVE) { if (part == 1) w.c_move_e += 1; else w.c_move_h += 1; }
                else { if (part == 1) w.c_capture_e += 1; else w.c_capture_h += 1; }
                if (traced && lane == 0 && w.fired - trace_base < (uint64_t)P.trace_cap) {
                    HfTraceRec rec;
                    rec.initial = (int32_t)hf_site(P, ini); rec.final_ = (int32_t)hf_site(P, fin);
                    rec.tau = tau; rec.kind = (uint8_t)kind; rec.particule = (uint8_t)part;
                    rec.pad[0] = rec.pad[1] = 0;
                    rec.spins = (uint32_t)w.spins; rec.draws = w.draws;
                    trace[w.fired - trace_base] = rec;
                }
                w.fired += 1;
            }
        }
        hf_store_state<32, kX>(P, w, lane);
    }
}

// Lattice._charges_injection + _events_creation (reseau.py:207-276): random.sample of the
// injection sites (random.py:359-448, both the pool and the set variant), flag toggles, then one
// initial move event per charge.
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) hf_inject_kernel(const HfParams P, int32_t *pool_scratch) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HfWarp w;
    hf_carve(w, hf_smem + (size_t)warp * hf_warp_smem_bytes(P.cap_c, P.cap_f, hf_xcap(P)), P.cap_c, P.cap_f, hf_xcap(P));
    const int64_t molecules = (int64_t)P.X * P.Y;
    for (int64_t local = (int64_t)blockIdx.x * kWarps + warp; local < P.B; local += (int64_t)gridDim.x * kWarps) {
        hf_load_state(P, w, local, lane);      // zero-initialised by the host
        int st = HF_ST_OK;
        bool exhausted = false;
        int32_t *pool = pool_scratch ? pool_scratch + ((int64_t)blockIdx.x * kWarps + warp) * molecules : nullptr;
        for (int t = 0; t < 2; ++t) {                                            // 219, 225
            const int z = t == 0 ? P.Z - 1 : 0;
            if (molecules <= P.setsize) {                                        // pool variant
                for (int64_t i = lane; i < molecules; i += 32) pool[i] = (int32_t)i;
                __syncwarp();
                for (int i = 0; i < P.C; ++i) {
                    const int64_t j = hf_randbelow(P, w, molecules - i, &exhausted);
                    const int32_t pick = pool[j];
                    __syncwarp();
                    if (lane == 0) { pool[j] = pool[molecules - i - 1]; w.pos(t)[i] = hf_pack(pick / P.Y, pick % P.Y, z); }
                    __syncwarp();
                }
            } else {                                                             // set variant
                for (int i = 0; i < P.C; ++i) {
                    int64_t j = hf_randbelow(P, w, molecules, &exhausted);
                    int cand = hf_pack((int)(j / P.Y), (int)(j % P.Y), z);
                    while (hf_contains(w.pos(t), i, cand, lane) && !exhausted) {
                        j = hf_randbelow(P, w, molecules, &exhausted);
                        cand = hf_pack((int)(j / P.Y), (int)(j % P.Y), z);
                    }
                    if (lane == 0) w.pos(t)[i] = cand;
                    __syncwarp();
                }
            }
            w.set_n_pos(t, P.C);
        }
        if (exhausted) st = HF_ST_REPLAY_EXHAUSTED;
        for (int i = 0; i < P.C && st == HF_ST_OK; ++i)                          // 226-228
            for (int t = 0; t < 2; ++t)
                if (!hf_toggle_flag(P, w, t, w.pos(t)[i], lane)) st = HF_ST_CAPACITY;
        for (int t = 0; t < 2 && st == HF_ST_OK; ++t)                            // 240-276
            for (int i = 0; i < P.C && st == HF_ST_OK; ++i)
                st = hf_gen_move_event(P, w, t, w.pos(t)[i], false, lane);
        w.injection = 2ull * (uint64_t)P.C;                                      // 118
        w.status = st;
        hf_store_state(P, w, lane);
    }
}

// Known-answer entry: tau and k of hand-built hops in the trajectory's current state.
__global__ void hf_eval_hops_kernel(const HfParams P, int64_t local, int t, int n, const int32_t *initial,
                                    const int32_t *final_, const double *u, double *tau, double *rate, int32_t *zero_div) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31;
    HfWarp w;
    hf_carve(w, hf_smem, P.cap_c, P.cap_f, hf_xcap(P));
    hf_load_state(P, w, local, lane);
    for (int i = 0; i < n; ++i) {
        const int ini = hf_pack_site(P, initial[i]), fin = hf_pack_site(P, final_[i]);
        hf_fill_inv_old(P, w, t, ini, lane);
        if (lane == 0) {
            bool zd = false;
            double k = 0.0;
            tau[i] = hf_hop_time(P, w, t, ini, fin, u[i], &k, &zd);
            rate[i] = k;
            zero_div[i] = zd ? 1 : 0;
        }
        __syncwarp();
    }
}

// Sum the per-trajectory counters into totals[0..4] and count stopped trajectories.
__global__ void hf_reduce_kernel(const HfScalars *sc, int64_t B, unsigned long long *totals) {
    unsigned long long v[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const HfScalars s = sc[i];
        v[0] += s.emission; v[1] += s.recombination; v[2] += s.injection; v[3] += s.steps; v[4] += s.fired;
        v[5] += s.status != HF_ST_OK;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(HF_FULL, v[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) if (v[k]) atomicAdd(totals + k, v[k]);
        if (v[5]) atomicAdd(totals + 16, v[5]);
    }
}

// Per-realisation sums (one block per realisation): emission, recombination, injection, steps,
// fired — the fitness inputs of a composition sweep (one composition per realisation).
__global__ void hf_reduce_realisations_kernel(const HfScalars *sc, int64_t traj_per_real, unsigned long long *out) {
    __shared__ unsigned long long part[5][32];
    const int64_t base = (int64_t)blockIdx.x * traj_per_real;
    unsigned long long v[5] = {0, 0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < traj_per_real; i += blockDim.x) {
        const HfScalars s = sc[base + i];
        v[0] += s.emission; v[1] += s.recombination; v[2] += s.injection; v[3] += s.steps; v[4] += s.fired;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(HF_FULL, v[k], off);
        if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[threadIdx.x][w];
        out[(int64_t)blockIdx.x * 5 + threadIdx.x] = t;
    }
}

// ---- lattice generation ----------------------------------------------------------------------------
// Species shuffle, reseau.py:164-172: random.sample([0,1,2], k=total, counts=...) is a pool-based
// partial Fisher-Yates over range(total) (random.py:426-432) mapped through bisect over the
// cumulative counts; quirk Q2 uses only the first X*Y selections, replicated in every interior
// layer.  Inherently sequential: one thread per realisation; `pool` holds range(total).
__global__ void hf_species_pool_init(int32_t *pool, int64_t total, int n_real) {
    const int64_t n = total * n_real;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        pool[i] = (int32_t)(i % total);
}

__global__ void hf_species_shuffle(const HfParams P, int32_t *pool_all, int64_t total, const int64_t *n_host_all,
                                   const int64_t *n_tadf_all, uint8_t *layer_all, int real_base) {
    const int r = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int64_t n_host = n_host_all[real_base + r], n_tadf = n_tadf_all[real_base + r];   // per-realisation composition
    int32_t *pool = pool_all + (int64_t)r * total;
    uint8_t *layer = layer_all + (int64_t)(real_base + r) * P.X * P.Y;
    const uint32_t ident = P.first_real + (uint32_t)(real_base + r);
    const int64_t maxsize = (int64_t)1 << 53;
    uint64_t d = 0;
    const int64_t xy = (int64_t)P.X * P.Y;
    for (int64_t i = 0; i < xy && i < total; ++i) {
        const int64_t n = total - i;
        const int64_t rem = maxsize % n;
        const double limit = (double)(maxsize - rem) * 0x1p-53;
        double u = hf_philox_uniform(P.k0, P.k1, HF_STREAM_SPECIES, ident, d++);
        while (u >= limit) u = hf_philox_uniform(P.k0, P.k1, HF_STREAM_SPECIES, ident, d++);
        const int64_t j = (int64_t)floor(u * 0x1p53) % n;
        const int32_t s = pool[j];
        pool[j] = pool[n - 1];
        layer[i] = (uint8_t)(s < n_host ? 0 : (s < n_host + n_tadf ? 1 : 2));   // bisect over cum_counts
    }
}

// Site species (faces all Host, reseau.py:170,172) and the four Gaussian energies of the species
// constructors (molecule.py:180-183 etc.; random.gauss = Box-Muller pair, random.py:543-589).
__global__ void hf_fill_sites(const HfParams P, const uint8_t *layer_all, int n_real, uint8_t *species,
                              double *homo, double *lumo, double *s1, double *t1) {
    const double means[3][4] = {{6.0, -2.0, 3.50, 3.00}, {5.8, -2.6, 2.55, 2.52}, {5.25, -1.84, 2.69, 1.43}};
    const int64_t xy = (int64_t)P.X * P.Y, n = P.N * n_real;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = g / P.N, i = g % P.N;
        const int64_t z = i / xy, q = i % xy;
        const int sp = (z > 0 && z < P.Z - 1) ? layer_all[r * xy + q] : 0;
        const uint32_t ident = P.first_real + (uint32_t)r;
        double e[4];
#pragma unroll
        for (int pair = 0; pair < 2; ++pair) {
            const double ua = hf_philox_uniform(P.k0, P.k1, HF_STREAM_ENERGY, ident, (uint64_t)(4 * i + 2 * pair));
            const double ub = hf_philox_uniform(P.k0, P.k1, HF_STREAM_ENERGY, ident, (uint64_t)(4 * i + 2 * pair + 1));
            const double x2pi = __dmul_rn(ua, 0x1.921fb54442d18p+2);
            const double g2rad = __dsqrt_rn(__dmul_rn(-2.0, log(__dsub_rn(1.0, ub))));
            double sn, cs;
            sincos(x2pi, &sn, &cs);
            e[2 * pair] = __dadd_rn(means[sp][2 * pair], __dmul_rn(__dmul_rn(cs, g2rad), 0.1));
            e[2 * pair + 1] = __dadd_rn(means[sp][2 * pair + 1], __dmul_rn(__dmul_rn(sn, g2rad), 0.1));
        }
        species[g] = (uint8_t)sp;
        homo[g] = e[0]; lumo[g] = e[1]; s1[g] = e[2]; t1[g] = e[3];
    }
}
