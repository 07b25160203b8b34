// Thread-per-trajectory First-Reaction-Method kernel for charges == 1 (one electron and one hole
// in flight per trajectory: BASELINE configs 1, 2, 3 and 5).
//
// Why a second kernel: ncu on the warp-per-trajectory kernel (profiles/r01a_*) showed 1174 warp
// instructions per event at 26/32 active lanes and an fp64 pipe only 12.7 % busy — the step is
// bound by the warp-wide bookkeeping of the ordered lists, not by the rate arithmetic.  With a
// single charge of each type the lists degenerate to one slot each, so the whole trajectory state
// fits in registers and every lane can run its own trajectory: the 17-26 candidate hops become a
// per-thread loop, no shuffles, no shared memory, 32/32 lanes busy.
//
// Semantics are those of hf_frm.cuh (same citations), specialised by these facts about C = 1:
//   * the Molecule.electron / .hole flags always agree with the position lists (no second charge
//     of the same type can ever desynchronise the toggles), except on the exciton's site between
//     its `bound` and `decay` steps, when no candidate is evaluated;
//   * a candidate is therefore never blocked (the only same-type charge is the one moving): every
//     regenerated event draws exactly 26 (bulk) or 17 (z face) uniforms;
//   * the same-type Coulomb loop only meets the moving charge itself (skipped, reseau.py:300/332);
//     the opposite-type loop has at most one term;
//   * re-injection always finds its own list empty: choice() is randbelow(X*Y) (reseau.py:448-470).
// If any of these were ever violated the kernel stops the trajectory with HF_ST_CAPACITY instead
// of continuing with a wrong state.  State is exchanged through the same global arrays as the
// warp kernel (cap_c = 1, cap_f = 2), so the two kernels can be cross-checked on the same handle.
#pragma once
#include <type_traits>
#include "hf_frm.cuh"

struct HfC1 {
    // named scalars + select accessors: nothing is dynamically indexed, everything stays in registers
    int32_t e_pos, h_pos;      // packed position or -1
    int32_t e_fin, h_fin;      // pending move event: final site or -1
    double e_tau, h_tau;
    int32_t special, special_site, xstate, status, cache_head, cache_n;
    uint64_t draws, spins, emission, recombination, injection, steps, fired;

    __device__ __forceinline__ int32_t pos(int t) const { return t ? h_pos : e_pos; }
    __device__ __forceinline__ int32_t fin(int t) const { return t ? h_fin : e_fin; }
    __device__ __forceinline__ double tau(int t) const { return t ? h_tau : e_tau; }
    __device__ __forceinline__ void set_pos(int t, int32_t v) { if (t) h_pos = v; else e_pos = v; }
    __device__ __forceinline__ void set_fin(int t, int32_t v) { if (t) h_fin = v; else e_fin = v; }
    __device__ __forceinline__ void set_tau(int t, double v) { if (t) h_tau = v; else e_tau = v; }
};

struct HfC1Counters {
    unsigned move_e, move_h, bound, decay, capture_e, capture_h, recomb_host, recomb_tadf, recomb_fluo, emit_tadf, emit_fluo;
};

// kReplay = false: the kernel was launched for a handle without a replay table (P.replay == nullptr) and the tests on
// it disappear from the candidate loop (they were ~4 % of its instructions).
template <bool kReplay = true>
__device__ __forceinline__ double hf_c1_uniform(const HfParams &P, int64_t local, uint32_t ident, uint64_t d, bool *exhausted) {
    if (kReplay && P.replay) {
        if (d >= (uint64_t)P.n_replay) { *exhausted = true; return 0.5; }
        return P.replay[local * P.n_replay + (int64_t)d];
    }
    return hf_block_uniform(hf_philox_block_rk(P.rk, HF_STREAM_TRAJ, ident, d >> 1), d);
}

// _electron_reinjection / _hole_reinjection with an empty own list (reseau.py:448-470)
template <bool kReplay = true>
__device__ __forceinline__ int hf_c1_reinject(const HfParams &P, HfC1 &s, int t, int64_t local, uint32_t ident) {
    if (s.pos(t) >= 0) return HF_ST_CAPACITY;                                    // C = 1 invariant
    const int64_t n = (int64_t)P.X * P.Y;
    const int64_t maxsize = (int64_t)1 << 53;
    const int64_t rem = maxsize % n;
    const double limit = (double)(maxsize - rem) * 0x1p-53;
    bool exhausted = false;
    double r = hf_c1_uniform<kReplay>(P, local, ident, s.draws++, &exhausted);   // random.py:264-269
    while (r >= limit && !exhausted) r = hf_c1_uniform<kReplay>(P, local, ident, s.draws++, &exhausted);
    if (exhausted) return HF_ST_REPLAY_EXHAUSTED;
    const int64_t k = (int64_t)floor(r * 0x1p53) % n;
    s.set_pos(t, hf_pack((int)(k / P.Y), (int)(k % P.Y), t == 0 ? P.Z - 1 : 0));
    s.injection += 1;
    return HF_ST_OK;
}

#ifndef HF_C1_THREADS
#define HF_C1_THREADS 384
#endif
#ifndef HF_C1_MIN_BLOCKS
#define HF_C1_MIN_BLOCKS 2
#endif
#define HF_C1_MAXCAND 27
#ifndef HF_C1_PEEL
#define HF_C1_PEEL 1
#endif
#ifndef HF_C1_PIN_BOUNDS
#define HF_C1_PIN_BOUNDS 1
#endif
// relative half-width of the "could be the minimum" window of the screening pass, see hf_c1_gen
#define HF_C1_WINDOW 0x1p-9f

// This thread's column of the shared-memory scratch, addressed through a 32-bit shared-window address pinned in a
// register: with a generic pointer nvcc rebuilt the address (S2R tid, S2UR cta id, ULEA, LEA, IMAD) at every access
// of the candidate loop.  volatile: the column is only ever accessed through these, in program order.
#ifndef HF_C1_COL_ASM
#define HF_C1_COL_ASM 1
#endif
#if defined(__CUDA_ARCH__) && HF_C1_COL_ASM
typedef unsigned hf_c1_col;
__device__ __forceinline__ hf_c1_col hf_c1_col_of(float *p) { unsigned a = (unsigned)__cvta_generic_to_shared(p); asm volatile("" : "+r"(a)); return a; }
__device__ __forceinline__ float hf_c1_ld(hf_c1_col c, int off) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(c + 4u * (unsigned)off)); return v; }
__device__ __forceinline__ void hf_c1_st(hf_c1_col c, int off, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(c + 4u * (unsigned)off), "f"(v)); }
#else
typedef float *hf_c1_col;
__host__ __device__ inline hf_c1_col hf_c1_col_of(float *p) { return p; }
__host__ __device__ inline float hf_c1_ld(hf_c1_col c, int off) { return c[off]; }
__host__ __device__ inline void hf_c1_st(hf_c1_col c, int off, float v) { c[off] = v; }
#endif
__host__ __device__ inline uint32_t hf_c1_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
__host__ __device__ inline uint32_t hf_c1_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
#if defined(__CUDA_ARCH__)
#define HF_C1_PIN(x) asm volatile("" : "+r"(x))
// MUFU.LG2 / MUFU.EX2 without the denormal pre- and post-scaling __logf / __expf wrap around them (3 + 4 instructions
// per candidate): used where the argument is known to be a normal number / the result at least 1
__device__ __forceinline__ float hf_c1_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float hf_c1_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
#define HF_C1_PIN(x) (void)0
inline float hf_c1_lg2(float x) { return log2f(x); }
inline float hf_c1_ex2(float x) { return exp2f(x); }
#endif

// Exact waiting time of one candidate hop (reseau.py:283-292 / 315-324 with the Coulomb term of
// 297-313 / 329-345 reduced to the single opposite charge): the same operation sequence as
// hf_hop_time in hf_frm.cuh, hence the same bits.
__device__ __forceinline__ double hf_c1_exact_tau(const double *__restrict__ E, int64_t cand_site, double e_init, double fz,
                                                  int dz, bool other_has, int other, int cand, double inv_old,
                                                  double u, bool *rate_is_zero) {
    const double rng = __dsub_rn(1.0, u);                                        // 284 / 316
    double dE = __dsub_rn(__ldg(E + cand_site), e_init);                         // 286 / 318
    dE = __dadd_rn(dE, __dmul_rn(fz, (double)dz));                               // 287 / 319
    double out = 0.0;
    if (other_has) {                                                             // 305-312 / 337-344
        if (other == cand) {
            out = -INFINITY;
        } else {
            const double inv_new = __ddiv_rn(1.0, __dsqrt_rn((double)hf_dist2(other, cand)));
            out = __dsub_rn(0.0, __dmul_rn(HF_ELECTROSTATIC, __dsub_rn(inv_new, inv_old)));
        }
    }
    dE = __dadd_rn(dE, out);                                                     // 288 / 320
    double k = HF_RATE0;
    if (dE >= 0) k = __dmul_rn(k, exp(__ddiv_rn(-dE, HF_KT)));                   // 290-291 / 322-323
    *rate_is_zero = k == 0.0;
    return __ddiv_rn(-log(rng), k);                                              // 292 / 324
}

// _new_move_electron_events / _new_move_hole_events for the single charge of type t
// (reseau.py:380-396): min() over the 17-26 candidate hops, first minimum wins.
//
// t is a RUNTIME value on purpose: lanes regenerating an electron event and lanes regenerating a
// hole event run the same instructions (only the energy array and the sign of the field term
// differ), so a warp with a mix of both does not serialise.
//
// Screen-then-verify.  Evaluating exp(), log(), a square root and three divisions in fp64 for all
// 26 candidates costs ~420 instructions per candidate (ncu, profiles/r01b_*), yet only the
// minimum is kept.  Pass 1 computes a cheap fp32 estimate a_i of 1e13 * tau_i for every candidate
// (one MUFU log, one MUFU exp, no Coulomb term) whose relative error is bounded by eps = 1e-4:
//   -log(rng): rng >= 2^-6 away from 1: lg2.approx * ln 2 (= __logf) abs error 2^-21.41 on |log| >= 0.0157 -> 2.7e-5
//              plus its fp32 argument (rounded, or built from the high word of the draw)  -> 3.8e-6
//              otherwise the series u + u^2/2 + u^3/3 (truncation u^3/4 <= 9.5e-7)      -> 2e-6
//   exp(dE/kT) = 2^x2, x2 = (float)dE * (float)(log2 e / kT): three fp32 roundings, 1.8e-7 relative in x, i.e.
//              1.6e-5 at x = 88; ex2.approx 2 ulp                                       -> 1.6e-5
//              neglected Coulomb term of the one opposite charge: C_e |1/r' - 1/r| with r, r' >= 1 is below
//              C_e = 4.8e-7 eV whatever the geometry (adjacent charges: C_e (1 - 1/2) = 2.4e-7 eV), i.e.
//              below C_e / kT = 1.86e-5 in x                                            -> 1.9e-5
//   sum 6.6e-5 < eps.
// Every candidate whose estimate lies within a factor 1 + W, W = 2^-9 (9.7x the 2 eps / (1 - eps) = 2.0e-4 the
// argument needs) of the smallest estimate is then evaluated EXACTLY, in candidate order, by hf_c1_exact_tau;
// the first minimum among those is the reference's min().  A candidate outside the window has
// tau_i >= a_i (1 - eps) > a_min (1 + W)(1 - eps) > a_min (1 + eps) >= tau of the best
// estimate, so it can neither win nor tie.  Estimates that overflow (x > 88) are +inf and only
// matter when all are (then every candidate is verified); candidates with x > 700 are always
// verified because the reference raises ZeroDivisionError when any rate underflows to 0.
// tests/test_gpu_parity.py::test_adjacent_charges_screening_is_exact keeps the two carriers on neighbouring sites
// for > 10^6 regenerated events and compares every one with the all-exact C restatement.
// Result: bit-identical events and taus at ~1.06 exact evaluations per event instead of 26.
template <int kStride = HF_C1_THREADS, bool kReplay = true>
__device__ __forceinline__ int hf_c1_gen(const HfParams &P, const double *lumo, const double *homo, int t, HfC1 &s,
                                         int64_t local, uint32_t ident, float *approx_ptr /* [l * kStride]: this thread's column */) {
    const hf_c1_col approx = hf_c1_col_of(approx_ptr);
    const bool replay = kReplay && P.replay;
    const int pos = s.pos(t), other = s.pos(1 - t);
    const double *__restrict__ E = t ? homo : lumo;
    const int px = hf_px(pos), py = hf_py(pos), pz = hf_pz(pos);
    const int nz = (pz - 1 > -1 && pz + 1 < P.Z) ? 3 : 2;
    const int total = 9 * nz;
    const uint64_t d0 = s.draws;
    if (other >= 0 && other == pos) { s.draws = d0 + 1; return HF_ST_ZERO_DIVISION; }   // 1/old_r.norm(), first candidate
    const double e_init = __ldg(E + hf_site(P, pos));
    const double fz = t == 0 ? -P.field : P.field;                               // reseau.py:287 / 319
    const int64_t plane = (int64_t)P.X * P.Y;

    // ---- pass 0: gather the candidates' energies (all loads in flight together), store dE as
    //      fp32 in this thread's shared-memory column; remember where `self` sits ----
    int self_l = -1;
    {
        int l = 0;
        for (int ix = 0; ix < 3; ++ix) {
            const int cx = hf_bvk_xy(px, P.X, ix);
#pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
                const int cy = hf_bvk_xy(py, P.Y, iy);
                const double *row = E + ((int64_t)cy * P.X + cx);
#pragma unroll
                for (int iz = 0; iz < 3; ++iz) {
                    if (iz < nz) {
                        const int cz = hf_bvk_z(pz, P.Z, iz);
                        const int cand = hf_pack(cx, cy, cz);
                        if (cand == pos) self_l = l;
                        const double dE = (__ldg(row + cz * plane) - e_init) + fz * (double)(cz - pz);
                        hf_c1_st(approx, l * kStride, cand == other ? -INFINITY : (float)dE);   // onto the opposite charge: k = 1e13
                        ++l;
                    }
                }
            }
        }
    }
    // ---- pass 1: fp32 estimates a_l of 1e13 * tau_l.  Uniform d0 + r belongs to Philox block
    //      (d0 >> 1) + ((p + r) >> 1), half (p + r) & 1 with p = d0 & 1: looping over q = p + r in
    //      pairs makes every lane compute its blocks at the same iterations ----
    int n = total - 1;                                                           // never blocked when C = 1
    int p = (int)(d0 & 1);
#if HF_C1_PIN_BOUNDS & 1
    HF_C1_PIN(n); HF_C1_PIN(p);                                                  // in registers: the loop tests were rebuilding them from the position
#endif
#if HF_C1_PIN_BOUNDS & 2
    HF_C1_PIN(self_l);
#endif
    const float inv_kt_log2e = (float)(1.4426950408889634 / HF_KT);              // x / kT in units of ln 2: a = t0 * 2^x2
    const float x2_forced = (float)(700.0 * 1.4426950408889634);                 // x > 700: the rate may underflow to 0
    // the two smallest estimates as integer keys: the bits of a >= 0 order like a, the low 5 mantissa bits carry l
    // (estimates truncated by <= 2^-18 relative, far inside the window W).  Storing the scratch in uniform order
    // instead (no l = r + (r >= self) in this loop) was measured 9 % SLOWER: the conditional store serialises pass 0.
    uint32_t key1 = 0x7f800000u, key2 = 0x7f800000u;
    float x2max = 0.0f;
    bool exhausted = false;
    // One Philox block = two candidates.  Only the first and the last block of an event can hold a uniform that belongs to
    // no candidate (r < 0 when the draw counter is odd, r >= n at the end): they run the checked copy of the body, the
    // 11-12 blocks in between the unchecked one (5 instructions per candidate less).
    const int jn = (n + p + 1) >> 1;
    auto block = [&](const int j, auto checked) {
        constexpr bool kChecked = decltype(checked)::value;
        HfPhiloxBlock blk = {0, 0, 0, 0};
        if (!replay) blk = hf_philox_block_rk(P.rk, HF_STREAM_TRAJ, ident, (d0 >> 1) + (uint64_t)j);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 2 * j + h - p;
            if (kChecked && (r < 0 || r >= n)) continue;
            const int l = r + (r >= self_l ? 1 : 0);
            const float x2 = hf_c1_ld(approx, l * kStride) * inv_kt_log2e;
            x2max = fmaxf(x2max, x2);
            float a;
            const uint32_t hi = h ? blk.w3 : blk.w1;                             // u = hi * 2^-32 + (lower bits) * 2^-53
            if (!replay && hi - 0x00100000u < 0xffe00000u) {
                // 2^-12 <= u < 1 - 2^-12 (all but 5e-4 of the draws): the estimate straight from the high word, no fp64.
                // u_f = hi 2^-32 and v = (~hi) 2^-32 ~ 1 - u are off by < 2^-32 absolute and 2^-24 relative:
                // relative error of the series <= 1e-6, of -log(v) <= (2^-32 / v + 2^-24) / |log v| <= 3.8e-6.
                // t0 > 0 here (u >= 2^-12), so 0 * inf cannot occur.
                float t0;
                if (hi < 0x04000000u) {                                          // u < 2^-6
                    const float uf = (float)hi * 0x1p-32f;
                    t0 = uf * (1.0f + uf * (0.5f + uf * (1.0f / 3.0f)));
                } else {
                    t0 = hf_c1_lg2((float)(~hi) * 0x1p-32f) * -0.69314718055994531f;   // v in [2^-12, 0.985): a normal number
                }
                a = x2 > 0.0f ? t0 * hf_c1_ex2(x2) : t0;
            } else {
                const double u = replay ? hf_c1_uniform<kReplay>(P, local, ident, d0 + (uint64_t)r, &exhausted)
                                        : hf_block_uniform(blk, (uint64_t)h);
                float t0;
                if (u < 0x1p-6) {
                    const float uf = (float)u;
                    t0 = uf * (1.0f + uf * (0.5f + uf * (1.0f / 3.0f)));
                } else {
                    t0 = -__logf((float)(1.0 - u));
                }
                a = x2 > 0.0f ? t0 * hf_c1_ex2(x2) : t0;
                if (t0 == 0.0f) a = 0.0f;                                        // 0 * inf
            }
            hf_c1_st(approx, l * kStride, a);
            const uint32_t key = (__float_as_uint(a) & ~31u) | (uint32_t)l;
            key2 = hf_c1_umin(key2, hf_c1_umax(key1, key));
            key1 = hf_c1_umin(key1, key);
        }
    };
#if HF_C1_PEEL
    if (jn > 0) block(0, std::true_type{});
#pragma unroll 1
    for (int j = 1; j < jn - 1; ++j) block(j, std::false_type{});
    if (jn > 1) block(jn - 1, std::true_type{});
#else
#pragma unroll 1
    for (int j = 0; j < jn; ++j) block(j, std::true_type{});
#endif
    const float amin = __uint_as_float(key1 & ~31u), amin2 = __uint_as_float(key2 & ~31u);
    const int lmin = (int)(key1 & 31u);
    // candidates with x > 700 are always verified.  Their estimate overflowed to +inf (2^x2, x2 > 1009) or is 0 (t0 = 0;
    // zeros are verified as minima below); rare enough to find them by a second look, only when one was met
    unsigned forced = 0;
    if (x2max > x2_forced) {
        for (int i = 0; i < total; ++i)
            if (i != self_l && hf_c1_ld(approx, i * kStride) == INFINITY) forced |= 1u << i;
    }
    if (exhausted) { s.draws = d0 + (uint64_t)n; return HF_ST_REPLAY_EXHAUSTED; }
    // ---- candidates that could be the minimum: usually only the best estimate ----
    const float thr = amin * (1.0f + HF_C1_WINDOW);
    unsigned mask = forced | (1u << lmin);
    if (!(amin2 > thr)) {                                                        // a runner-up inside the window: scan
        for (int i = 0; i < total; ++i)
            if (i != self_l && !(hf_c1_ld(approx, i * kStride) > thr)) mask |= 1u << i;
    }
    // ---- pass 2: exact evaluation in candidate order, first minimum ----
    const bool other_has = other >= 0;
    double inv_old = 0.0;
    if (other_has) inv_old = __ddiv_rn(1.0, __dsqrt_rn((double)hf_dist2(other, pos)));
    double best = 0.0;
    int32_t bfin = -1;
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const int ix = i / (3 * nz), iy = (i / nz) % 3, iz = i % nz;
        const int cx = hf_bvk_xy(px, P.X, ix), cy = hf_bvk_xy(py, P.Y, iy), cz = hf_bvk_z(pz, P.Z, iz);
        const int cand = hf_pack(cx, cy, cz);
        const uint64_t di = d0 + (uint64_t)(i - (i > self_l ? 1 : 0));
        const double u = hf_c1_uniform<kReplay>(P, local, ident, di, &exhausted);
        bool rate_is_zero;
        const double tau = hf_c1_exact_tau(E, ((int64_t)cz * P.Y + cy) * P.X + cx, e_init, fz, cz - pz, other_has, other,
                                           cand, inv_old, u, &rate_is_zero);
        if (rate_is_zero) { s.draws = di + 1; return HF_ST_ZERO_DIVISION; }      // float division by zero
        if (bfin < 0 || tau < best) { best = tau; bfin = cand; }                 // min(): first minimum
    }
    s.draws = d0 + (uint64_t)n;
    s.set_fin(t, bfin);
    s.set_tau(t, best);
    return HF_ST_OK;
}

// _first_reaction_method (reseau.py:483-562) for one thread's trajectory.  Bookkeeping first;
// the (at most two) event regenerations it asks for are queued in `gen` (bits 0-1: first type,
// bit 2: re-inject it first; bits 4-5 / 6: second one, only after a decay) and executed by the
// caller's common tail so that the whole warp runs the candidate loop together.
#define HF_GEN_NONE 0
__device__ __forceinline__ int hf_c1_step(const HfParams &P, const uint8_t *species, HfC1 &s, HfC1Counters &c,
                                          int64_t local, uint32_t ident, int *gen,
                                          int *ev_kind, int *ev_part, int *ev_init, int *ev_final, double *ev_tau) {
    int kind, part, ini, fin;
    double tau;
    *gen = HF_GEN_NONE;
    if (s.special) {
        kind = s.special & 0xff; part = s.special >> 8; ini = fin = s.special_site; tau = 0.0;
    } else {
        const bool he = s.e_fin >= 0, hh = s.h_fin >= 0;
        if (!he && !hh) return HF_ST_NO_EVENTS;                                  // 487-488
        const int t = (hh && (!he || s.h_tau <= s.e_tau)) ? 1 : 0;               // ties: last in concat order (Q6)
        kind = HF_K_MOVE; part = t + 1; ini = s.pos(t); fin = s.fin(t); tau = s.tau(t);
    }
    // _cache (489-490): slot-major ring in global memory, coalesced across the warp's trajectories
    {
        const int64_t at = (int64_t)s.cache_head * P.B + local;
        P.cache_init[at] = ini; P.cache_final[at] = fin; P.cache_kp[at] = kind | (part << 8); P.cache_tau[at] = tau;
        s.cache_head = s.cache_head + 1 == HF_CACHE ? 0 : s.cache_head + 1;
        if (s.cache_n < HF_CACHE) s.cache_n += 1;
    }
    *ev_kind = kind; *ev_part = part; *ev_init = ini; *ev_final = fin; *ev_tau = tau;

    if (kind == HF_K_MOVE) {                                                     // 492-524
        const int t = part - 1, o = 1 - t;
        if (ini < 0) return HF_ST_VALUE_ERROR;
        s.set_fin(t, -1);                                                        // 495 / 511: its own event
        s.set_pos(t, fin);                                                       // 422-432
        const bool opposite_here = s.pos(o) == fin;
        const int capture_z = t == 0 ? 0 : P.Z - 1;
        if (!opposite_here && hf_pz(fin) != capture_z) {                         // 498-500 / 514-516
            if (s.fin(o) >= 0) s.set_tau(o, __dsub_rn(s.tau(o), tau));           // _update_events
            *gen = 8 | t;
        } else if (opposite_here) {                                              // 501-505 / 517-521 (Q5)
            s.set_fin(o, -1);
            s.special = HF_K_BOUND | (HF_PART_EXCITON << 8); s.special_site = fin;
        } else {                                                                 // 506-508 / 522-524
            if (s.fin(o) >= 0) s.set_tau(o, __dsub_rn(s.tau(o), tau));
            s.special = HF_K_CAPTURE | (part << 8); s.special_site = fin;
        }
        return HF_ST_OK;
    }
    const int site = s.special_site;
    s.special = 0;
    if (kind == HF_K_BOUND) {                                                    // 529-533
        if (s.e_pos != site || s.h_pos != site) return HF_ST_VALUE_ERROR;
        const double r = hf_philox_uniform(P.k0, P.k1, HF_STREAM_SPIN, ident, s.spins);   // molecule.py:88-96
        s.spins += 1;
        s.xstate = (0 <= r && r < 0.25) ? HF_XS_SINGLET : HF_XS_TRIPLET;
        s.e_pos = s.h_pos = -1;                                                  // 434-438
        s.special = HF_K_DECAY | (HF_PART_EXCITON << 8); s.special_site = site;
        return HF_ST_OK;
    }
    if (kind == HF_K_DECAY) {                                                    // 541-547
        const int sp = species[hf_site(P, site)];
        bool photon = false;
        if (s.xstate) { photon = sp != HF_SPECIES_HOST && s.xstate == HF_XS_SINGLET; s.xstate = 0; }
        s.recombination += 1;                                                    // 472-476
        c.recomb_host += sp == 0; c.recomb_tadf += sp == 1; c.recomb_fluo += sp == 2;
        if (photon) { s.emission += 1; c.emit_tadf += sp == 1; c.emit_fluo += sp == 2; }
        *gen = (8 | 4 | 0) | ((8 | 4 | 1) << 4);                                 // re-inject + regenerate e, then h
        return HF_ST_OK;
    }
    const int t = part - 1;                                                      // capture 551-561
    if (s.pos(t) != site) return HF_ST_VALUE_ERROR;
    s.set_pos(t, -1);                                                            // 440-446
    *gen = 8 | 4 | t;
    return HF_ST_OK;
}

__device__ __forceinline__ void hf_c1_load(const HfParams &P, HfC1 &s, int64_t local) {
    const HfScalars sc = P.sc[local];
    s.e_pos = sc.n_pos[0] > 0 ? P.pos[0][local] : -1;
    s.h_pos = sc.n_pos[1] > 0 ? P.pos[1][local] : -1;
    s.e_fin = sc.n_ev[0] > 0 ? P.ev_final[0][local] : -1;
    s.h_fin = sc.n_ev[1] > 0 ? P.ev_final[1][local] : -1;
    s.e_tau = sc.n_ev[0] > 0 ? P.ev_tau[0][local] : 0.0;
    s.h_tau = sc.n_ev[1] > 0 ? P.ev_tau[1][local] : 0.0;
    s.special = sc.special; s.special_site = sc.special_site; s.xstate = sc.xstate; s.status = sc.status;
    s.cache_head = sc.cache_head; s.cache_n = sc.cache_n;
    s.draws = sc.draws; s.spins = sc.spins; s.emission = sc.emission; s.recombination = sc.recombination;
    s.injection = sc.injection; s.steps = sc.steps; s.fired = sc.fired;
}

__device__ __forceinline__ void hf_c1_store(const HfParams &P, const HfC1 &s, const HfC1Counters &c, int64_t local) {
    HfScalars sc;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        sc.n_pos[t] = s.pos(t) >= 0;
        sc.n_ev[t] = s.fin(t) >= 0;
        if (s.pos(t) >= 0) P.pos[t][local] = s.pos(t);
        if (s.fin(t) >= 0) { P.ev_init[t][local] = s.pos(t); P.ev_final[t][local] = s.fin(t); P.ev_tau[t][local] = s.tau(t); }
        // flag sets (cap_f = 2): the charge's own site; on the exciton's site both flags stay set
        // until the decay step clears them (molecule.py:185-199)
        const bool lingering = (s.special & 0xff) == HF_K_DECAY;
        const int fsite = lingering ? s.special_site : s.pos(t);
        sc.n_flag[t] = fsite >= 0;
        if (fsite >= 0) P.flag[t][local * P.cap_f] = fsite;
    }
    sc.special = s.special; sc.special_site = s.special_site; sc.xstate = s.xstate; sc.status = s.status;
    sc.cache_head = s.cache_head; sc.cache_n = s.cache_n;
    sc.draws = s.draws; sc.spins = s.spins; sc.emission = s.emission; sc.recombination = s.recombination;
    sc.injection = s.injection; sc.steps = s.steps; sc.fired = s.fired;
    P.sc[local] = sc;
    // channel counters: warp-aggregate, one atomic per non-zero counter per warp per launch
    const unsigned v[11] = {c.move_e, c.move_h, c.bound, c.decay, c.capture_e, c.capture_h,
                            c.recomb_host, c.recomb_tadf, c.recomb_fluo, c.emit_tadf, c.emit_fluo};
#pragma unroll
    for (int i = 0; i < 11; ++i)
        if (v[i]) atomicAdd(P.totals + 5 + i, (unsigned long long)v[i]);
}

// Lattice.operations(stop) for charges == 1: one thread per trajectory.
// Two shapes: 384 threads x 2 CTAs per SM (80 registers: 113 664 trajectories per wave on 148 SMs, the fastest per
// event) and 288 x 3 (72 registers: 127 872 per wave, ~14 % slower per event).  The host picks the one that finishes
// the batch sooner: a batch just above a multiple of the first wave size (10^6 trajectories over 8 GPUs = 125 000
// each) would otherwise pay a whole extra wave for its last 10 %.
#define HF_C1_THREADS_B 288
#define HF_C1_MIN_BLOCKS_B 3
template <int kThreads = HF_C1_THREADS, int kMinBlocks = HF_C1_MIN_BLOCKS, bool kReplay = true>
__global__ void __launch_bounds__(kThreads, kMinBlocks) hf_run_c1_kernel(const HfParams P) {
    __shared__ float approx_all[HF_C1_MAXCAND * kThreads];
    float *approx = approx_all + threadIdx.x;
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= P.B) return;
    HfC1 s;
    HfC1Counters c = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    hf_c1_load(P, s, local);
    uint32_t ident = P.first_traj + (uint32_t)local;
    HF_C1_PIN(ident);                                                            // else rebuilt from tid / ctaid at every Philox block
    const int64_t r = local / P.traj_per_real;
    const uint8_t *species = P.species + r * P.N;
    const double *lumo = P.lumo + r * P.N, *homo = P.homo + r * P.N;
    if (s.status == HF_ST_ZERO_DIVISION || s.status == HF_ST_NO_EVENTS) s.status = HF_ST_OK;
    const bool traced = P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
    HfTraceRec *trace = traced ? P.trace + (local - P.trace_first) * P.trace_cap : nullptr;
    const uint64_t trace_base = traced ? P.trace_base[local - P.trace_first] : 0;   // events fired before the trace was armed
    if (s.status == HF_ST_OK) {
        for (int64_t i = 0; i < P.n_steps; ++i) {
            s.steps += 1;                                                        // 582
            int kind, part, ini, fin, gen;
            double tau;
            int st = hf_c1_step(P, species, s, c, local, ident, &gen, &kind, &part, &ini, &fin, &tau);
            // common tail: the queued regenerations (one for a move or a capture, two for a decay)
            for (; st == HF_ST_OK && gen; gen >>= 4) {
                const int t = gen & 1;
                if (gen & 4) st = hf_c1_reinject<kReplay>(P, s, t, local, ident);
                if (st == HF_ST_OK) st = hf_c1_gen<kThreads, kReplay>(P, lumo, homo, t, s, local, ident, approx);
            }
            if (st != HF_ST_OK) { s.status = st; break; }                        // 587-591
            if (kind == HF_K_MOVE) { if (part == 1) c.move_e += 1; else c.move_h += 1; }
            else if (kind == HF_K_BOUND) c.bound += 1;
            else if (kind == HF_K_DECAY) c.decay += 1;
            else { if (part == 1) c.capture_e += 1; else c.capture_h += 1; }
            if (traced && s.fired - trace_base < (uint64_t)P.trace_cap) {
                HfTraceRec rec;
                rec.initial = (int32_t)hf_site(P, ini); rec.final_ = (int32_t)hf_site(P, fin);
                rec.tau = tau; rec.kind = (uint8_t)kind; rec.particule = (uint8_t)part;
                rec.pad[0] = rec.pad[1] = 0;
                rec.spins = (uint32_t)s.spins; rec.draws = s.draws;
                trace[s.fired - trace_base] = rec;
            }
            s.fired += 1;
        }
    }
    hf_c1_store(P, s, c, local);
}

// Lattice._charges_injection + _events_creation (reseau.py:207-276) for charges == 1, one thread per trajectory.
// random.sample(positions, k = 1) is one randbelow(X * Y) in both of its variants (random.py:421-448: the pool
// variant returns pool[j] = j, the set variant j), i.e. hf_c1_reinject; the initial events are hf_c1_gen with the
// membership test of reseau.py:249 / 268, which for a single charge of each type blocks nothing.  Draw order:
// electron site, hole site, the electron's candidates, the hole's candidates — as in hf_inject_kernel, whose
// results this kernel reproduces bit for bit (the golden vectors' init_* arrays check both).
__global__ void __launch_bounds__(HF_C1_THREADS, HF_C1_MIN_BLOCKS) hf_inject_c1_kernel(const HfParams P) {
    __shared__ float approx_all[HF_C1_MAXCAND * HF_C1_THREADS];
    float *approx = approx_all + threadIdx.x;
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= P.B) return;
    HfC1 s;
    HfC1Counters c = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    s.e_pos = s.h_pos = s.e_fin = s.h_fin = -1;
    s.e_tau = s.h_tau = 0.0;
    s.special = s.special_site = s.xstate = s.status = s.cache_head = s.cache_n = 0;
    s.draws = s.spins = s.emission = s.recombination = s.injection = s.steps = s.fired = 0;
    const uint32_t ident = P.first_traj + (uint32_t)local;
    const int64_t r = local / P.traj_per_real;
    const double *lumo = P.lumo + r * P.N, *homo = P.homo + r * P.N;
    int st = hf_c1_reinject(P, s, 0, local, ident);                              // 219
    if (st == HF_ST_OK) st = hf_c1_reinject(P, s, 1, local, ident);              // 225
    if (st == HF_ST_OK) st = hf_c1_gen(P, lumo, homo, 0, s, local, ident, approx);   // 240-257
    if (st == HF_ST_OK) st = hf_c1_gen(P, lumo, homo, 1, s, local, ident, approx);   // 259-276
    s.injection = 2;                                                             // 118
    s.status = st;
    hf_c1_store(P, s, c, local);
}
