// Thread-per-trajectory First-Reaction-Method kernel for SMALL `charges` (2 <= C <= 12): the
// reference's own driver configuration (OLED.py:6-10: charges = 4) and the GA fitness sweeps.
//
// ncu on the warp-per-trajectory kernel (profiles/r01a_*): 1174 warp instructions per event, half of
// them bookkeeping on lists of <= 2C entries that 32 lanes cannot share, the other half ONE candidate's
// rate arithmetic (the 26 candidates run in parallel lanes).  With a handful of charges it is cheaper to
// give every lane its own trajectory: the bookkeeping is hf_step<1> — the SAME statements as the warp
// kernel, instantiated with one lane per trajectory (hf_frm.cuh) — on per-thread shared-memory lists,
// and the candidate loop below screens the 17-26 candidates in fp32 before verifying the possible
// minima exactly, like hf_frm_c1.cuh does for C = 1.
#pragma once
#include "hf_frm.cuh"
#include "hf_frm_c1.cuh"

#ifndef HF_SMALL_MAX_CHARGES
#define HF_SMALL_MAX_CHARGES 12
#endif
#ifndef HF_SMALL_THREADS
#define HF_SMALL_THREADS 64
#endif
#ifndef HF_SMALL_MIN_BLOCKS
#define HF_SMALL_MIN_BLOCKS 7
#endif

// bytes of shared memory per trajectory: the lists of hf_carve + the screening column, an odd number of
// 8-byte words so that the threads of a warp spread over the banks
// lists of one trajectory without the _cache ring (written straight to global memory by hf_step<1>)
// xcap: exciton slots of the integrated modes (hf_frm_x.cuh): taus + 8 prefix checkpoints, three int lists
__host__ __device__ inline size_t hf_small_list_bytes(int cap_c, int cap_f, int xcap = 0) {
    return 8u * (4u * (size_t)cap_c) + 4u * (6u * (size_t)cap_c + 2u * (size_t)cap_f) +
           (xcap ? 8u * ((size_t)xcap + 8u) + 4u * 3u * (size_t)xcap : 0u);
}
__device__ __forceinline__ void hf_small_carve(HfWarp &w, unsigned char *base, int cap_c, int cap_f, int xcap = 0) {
    w.cap_c = cap_c; w.cap_f = cap_f; w.cache_stride = 1;
    double *d = reinterpret_cast<double *>(base);
    w.ev_tau_b = d; d += 2 * cap_c;
    w.inv_old_b = d; d += 2 * cap_c;
    w.term_b = nullptr;                                                          // only the warp kernel shares a Coulomb sum
    w.x_tau = w.x_scratch = nullptr; w.x_site = w.x_meta = w.x_final = nullptr;
    if (xcap) { w.x_tau = d; d += xcap; w.x_scratch = d; d += 8; }
    int32_t *i = reinterpret_cast<int32_t *>(d);
    if (xcap) { w.x_site = i; i += xcap; w.x_meta = i; i += xcap; w.x_final = i; i += xcap; }
    w.pos_b = i; i += 2 * cap_c;
    w.ev_init_b = i; i += 2 * cap_c;
    w.ev_final_b = i; i += 2 * cap_c;
    w.flag_b = i; i += 2 * cap_f;
    w.approx = reinterpret_cast<float *>(i);
    w.cache_init = w.cache_final = w.cache_kp = nullptr; w.cache_tau = nullptr;   // set by hf_load_state<1>
}
__host__ __device__ inline size_t hf_small_thread_bytes(int cap_c, int cap_f, int xcap = 0) {
    size_t b = hf_small_list_bytes(cap_c, cap_f, xcap) + 4u * 28u;
    b = (b + 7u) & ~(size_t)7u;
    if (((b >> 3) & 1u) == 0) b += 8;
    return b;
}

// _new_move_*_events (reseau.py:380-396) and the per-charge body of _init_move_*_events (240-276) for
// one thread's trajectory: min() over the unblocked candidates, first minimum wins, one TRAJ uniform
// per unblocked candidate in neighbour order.
//
// Screen-then-verify as in hf_c1_gen, with these differences: candidates can be blocked (they draw
// nothing), and the Coulomb sum has up to 2C - 1 terms.  The fp32 estimate still ignores it: every
// term is |1/r' - 1/r| <= 1 times HF_ELECTROSTATIC = 4.8e-7 eV, so n terms shift dE / kT by <= 1.86e-5 n
// and the estimate of 1e13 * tau by the same relative amount on top of hf_c1_gen's 5e-5 (its 6.6e-5 less the one
// Coulomb term it counts itself) =: eps; candidates
// outside a window of 1 + max(2^-9, 4 eps) around the best estimate cannot win or tie
// ((1 + 4 eps)(1 - eps) > 1 + eps).  (Measured: faster than the warp kernel up to 12 charges of each type.)  Whenever the reference could raise ZeroDivisionError for
// geometric reasons (a same-type charge on a candidate: the flag/list desynchronisation of DESIGN.md §0;
// an opposite charge on the moving charge's own site) every candidate is evaluated exactly, in order,
// as the warp kernel does.
__device__ __forceinline__ int hf_gen_move_event_serial(const HfParams &P, HfWarp &w, int t, int pos, bool use_flags) {
    const int px = hf_px(pos), py = hf_py(pos), pz = hf_pz(pos);
    const int nz = (pz - 1 > -1 && pz + 1 < P.Z) ? 3 : 2;
    const int total = 9 * nz;
    const int32_t *blk = use_flags ? w.flag(t) : w.pos(t);
    const int nb = use_flags ? w.n_flag(t) : w.n_pos(t);
    const int ns = w.n_pos(t), no = w.n_pos(1 - t);
    const double *__restrict__ E = w.energy(t);
    float *approx = w.approx;
    const double e_init = __ldg(E + hf_site(P, pos));
    const double fz = t == 0 ? -P.field : P.field;
    const int64_t plane = (int64_t)P.X * P.Y;
    // ---- pass 0: which candidates are blocked, lead onto an opposite charge or onto a same-type charge (the
    //      flag / list desynchronisation of DESIGN.md §0: ZeroDivisionError in the reference) — found from the
    //      short lists, not by testing every candidate against every entry; then the candidates' dE in fp32
    unsigned blocked = 0, onto_opp = 0, onto_same = 0;
    bool slow = false;
    {
        int xs = 0, ys = 0, zs = 0;                                              // the three columns / rows / planes, 10 bits each
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xs |= hf_bvk_xy(px, P.X, i) << (10 * i); ys |= hf_bvk_xy(py, P.Y, i) << (10 * i);
            zs |= (i < nz ? hf_bvk_z(pz, P.Z, i) : 1023) << (10 * i);             // 1023: no such plane (z < 1023 on any lattice)
        }
#pragma unroll 1
        for (int j = 0; j < nb; ++j) blocked |= hf_cand_bits(xs, ys, zs, nz, blk[j]);   // 385 / 394 ; 249 / 268
#pragma unroll 1
        for (int j = 0; j < no; ++j) { const int q = w.pos(1 - t)[j]; onto_opp |= hf_cand_bits(xs, ys, zs, nz, q); slow |= q == pos; }   // 1 / |loc - initial| = 1 / 0
#pragma unroll 1
        for (int j = 0; j < ns; ++j) { const int q = w.pos(t)[j]; if (q != pos) onto_same |= hf_cand_bits(xs, ys, zs, nz, q); }
    }
    unsigned amask = ((1u << total) - 1u) & ~blocked;                           // total <= 27; the charge's own site is cleared below
    {
        int l = 0;
        for (int ix = 0; ix < 3; ++ix) {
            const int cx = hf_bvk_xy(px, P.X, ix);
            for (int iy = 0; iy < 3; ++iy) {
                const int cy = hf_bvk_xy(py, P.Y, iy);
                const double *row = E + ((int64_t)cy * P.X + cx);
                for (int iz = 0; iz < nz; ++iz, ++l) {
                    const int cz = hf_bvk_z(pz, P.Z, iz);
                    if (hf_pack(cx, cy, cz) == pos) { amask &= ~(1u << l); continue; }
                    if (!(amask >> l & 1u)) continue;
                    const double dE = (__ldg(row + cz * plane) - e_init) + fz * (double)(cz - pz);
                    approx[l] = (onto_opp >> l & 1u) ? -INFINITY : (float)dE;    // onto an opposite charge: k = 1e13
                }
            }
        }
    }
    slow |= (onto_same & amask) != 0;
    const int n = __popc(amask);
    if (n == 0) return HF_ST_VALUE_ERROR;                                        // min() of an empty sequence
    hf_fill_inv_old<1>(P, w, t, pos, 0);
    const uint64_t d0 = w.draws;
    bool exhausted = false;
    double best = 0.0;
    int32_t bfin = -1;
    unsigned verify = amask;
    if (!slow) {
        // ---- pass 1: fp32 estimates of 1e13 * tau; uniform d0 + r is half (p + r) & 1 of Philox block (d0 >> 1) + ((p + r) >> 1).
        //      The arithmetic of hf_c1_gen (hf_frm_c1.cuh has the error budget): the estimate straight from the high word
        //      of the draw, MUFU.LG2 / MUFU.EX2 issued directly, candidates whose rate could underflow found by a max and
        //      a rare second look, the two smallest estimates as integer keys (candidate index in the 5 low mantissa bits)
        int p = (int)(d0 & 1), nn = n;
        int replay = P.replay != nullptr;
        HF_C1_PIN(p); HF_C1_PIN(nn); HF_C1_PIN(replay);                          // else rebuilt / reloaded at every use in the loop
        const float inv_kt_log2e = (float)(1.4426950408889634 / HF_KT), x2_forced = (float)(700.0 * 1.4426950408889634);
        uint32_t key1 = 0x7f800000u, key2 = 0x7f800000u;
        float x2max = 0.0f;
        unsigned rem = amask;
        // (peeling the first / last block as hf_c1_gen does was measured 11 % SLOWER here: this kernel is bound by
        //  instruction fetch and the extra copy of the body costs more than the tests it removes)
        for (int j = 0; 2 * j < nn + p; ++j) {
            HfPhiloxBlock rnd = {0, 0, 0, 0};
            if (!replay) rnd = hf_philox_block_rk(P.rk, HF_STREAM_TRAJ, w.ident, (d0 >> 1) + (uint64_t)j);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 2 * j + h - p;
                if (r < 0 || r >= nn) continue;
                const int l = __ffs(rem) - 1;
                rem &= rem - 1;
                const float x2 = approx[l] * inv_kt_log2e;
                x2max = fmaxf(x2max, x2);
                float a;
                const uint32_t hi = h ? rnd.w3 : rnd.w1;
                if (!replay && hi - 0x00100000u < 0xffe00000u) {                 // 2^-12 <= u < 1 - 2^-12: no fp64, t0 > 0
                    float t0;
                    if (hi < 0x04000000u) {
                        const float uf = (float)hi * 0x1p-32f;
                        t0 = uf * (1.0f + uf * (0.5f + uf * (1.0f / 3.0f)));
                    } else {
                        t0 = hf_c1_lg2((float)(~hi) * 0x1p-32f) * -0.69314718055994531f;
                    }
                    a = x2 > 0.0f ? t0 * hf_c1_ex2(x2) : t0;
                } else {
                    const double u = replay ? hf_traj_uniform(P, w, d0 + (uint64_t)r, &exhausted) : hf_block_uniform(rnd, (uint64_t)h);
                    float t0;
                    if (u < 0x1p-6) {
                        const float uf = (float)u;
                        t0 = uf * (1.0f + uf * (0.5f + uf * (1.0f / 3.0f)));
                    } else {
                        t0 = -__logf((float)(1.0 - u));
                    }
                    a = x2 > 0.0f ? t0 * hf_c1_ex2(x2) : t0;
                    if (t0 == 0.0f) a = 0.0f;                                    // 0 * inf
                }
                approx[l] = a;
                const uint32_t key = (__float_as_uint(a) & ~31u) | (uint32_t)l;
                key2 = hf_c1_umin(key2, hf_c1_umax(key1, key));
                key1 = hf_c1_umin(key1, key);
            }
        }
        if (exhausted) { w.draws = d0 + (uint64_t)n; return HF_ST_REPLAY_EXHAUSTED; }
        const float amin = __uint_as_float(key1 & ~31u), amin2 = __uint_as_float(key2 & ~31u);
        const int lmin = (int)(key1 & 31u);
        unsigned forced = 0;                                                     // x > 700: estimate +inf (or 0, verified as a minimum)
        if (x2max > x2_forced) {
            for (unsigned m = amask; m; m &= m - 1) {
                const int i = __ffs(m) - 1;
                if (approx[i] == INFINITY) forced |= 1u << i;
            }
        }
        // window: 4 x the estimate's error bound, 5e-5 + C_e (ns + no) / kT (2^-9 up to 8 charges of each type)
        const float thr = amin * (1.0f + fmaxf(HF_C1_WINDOW, 4.0f * (5e-5f + (float)(HF_ELECTROSTATIC / HF_KT) * (float)(ns + no))));
        verify = forced | (1u << lmin);
        if (!(amin2 > thr)) {
            for (unsigned m = amask; m; m &= m - 1) {
                const int i = __ffs(m) - 1;
                if (!(approx[i] > thr)) verify |= 1u << i;
            }
        }
    }
    // ---- pass 2: exact evaluation in candidate order, first minimum (387 / 396)
    for (unsigned m = verify; m; m &= m - 1) {
        const int i = __ffs(m) - 1;
        const int ix = i / (3 * nz), iy = (i / nz) % 3, iz = i % nz;
        const int cand = hf_pack(hf_bvk_xy(px, P.X, ix), hf_bvk_xy(py, P.Y, iy), hf_bvk_z(pz, P.Z, iz));
        const int rank = __popc(amask & ((1u << i) - 1u));
        const double u = hf_traj_uniform(P, w, d0 + (uint64_t)rank, &exhausted);
        if (exhausted) return HF_ST_REPLAY_EXHAUSTED;
        double rate = 0.0;
        bool zero_div = false;
        const double tau = hf_hop_time<true>(P, w, t, pos, cand, u, &rate, &zero_div);
        if (zero_div) { w.draws = d0 + (uint64_t)rank + 1; return HF_ST_ZERO_DIVISION; }   // draws up to and including the failing candidate
        if (bfin < 0 || tau < best) { best = tau; bfin = cand; }
    }
    w.draws = d0 + (uint64_t)n;
    if (w.n_ev(t) >= P.cap_c) return HF_ST_CAPACITY;
    w.ev_init(t)[w.n_ev(t)] = pos; w.ev_final(t)[w.n_ev(t)] = bfin; w.ev_tau(t)[w.n_ev(t)] = best;
    w.set_n_ev(t, w.n_ev(t) + 1);
    return HF_ST_OK;
}

// Lattice.operations(stop) for 2 <= charges <= HF_SMALL_MAX_CHARGES: one thread per trajectory.
template <bool kX = false>
__global__ void __launch_bounds__(HF_SMALL_THREADS, HF_SMALL_MIN_BLOCKS) hf_run_small_kernel(const HfParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= P.B) return;
    const size_t stride = hf_small_thread_bytes(P.cap_c, P.cap_f, hf_xcap(P));
    HfWarp w;
    hf_small_carve(w, hf_smem + threadIdx.x * stride, P.cap_c, P.cap_f, hf_xcap(P));
    hf_load_state<1, kX>(P, w, local, 0);
    if (w.status == HF_ST_ZERO_DIVISION || w.status == HF_ST_NO_EVENTS) w.status = HF_ST_OK;   // reseau.py:587-589
    const bool traced = P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
    HfTraceRec *trace = traced ? P.trace + (local - P.trace_first) * P.trace_cap : nullptr;
    const uint64_t trace_base = traced ? P.trace_base[local - P.trace_first] : 0;   // events fired before the trace was armed
    if (w.status == HF_ST_OK) {
        for (int64_t i = 0; i < P.n_steps; ++i) {
            w.steps += 1;                                                        // 582
            int kind, part, ini, fin;
            double tau;
            int queue = 0;
            int st = hf_step<1, kX>(P, w, 0, &kind, &part, &ini, &fin, &tau, &queue);
            // common tail: the regenerations the step asked for (one after a move or a capture, two after a decay)
            for (; st == HF_ST_OK && queue; queue >>= 4) {
                const int t = queue & 1;
                if (queue & 4) st = hf_reinject<1>(P, w, t, 0);
                if (st == HF_ST_OK) st = hf_gen_move_event_serial(P, w, t, w.pos(t)[w.n_pos(t) - 1], true);
            }
            if (st != HF_ST_OK) { w.status = st; break; }                        // 587-591
            if (kind == HF_K_BOUND) w.c_bound += 1;
            else if (kind == HF_K_DECAY) w.c_decay += 1;
            else if (part == HF_PART_EXCITON) {}                                 // transfers and spin flips: counted by hfi_fire
            else if (kind == HF_K_MOVE) { if (part == 1) w.c_move_e += 1; else w.c_move_h += 1; }
            else { if (part == 1) w.c_capture_e += 1; else w.c_capture_h += 1; }
            if (traced && w.fired - trace_base < (uint64_t)P.trace_cap) {
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
    hf_store_state<1, kX>(P, w, 0);
}

// Lattice._charges_injection + _events_creation (reseau.py:207-276) for 2 <= charges <= HF_SMALL_MAX_CHARGES, one thread
// per trajectory: the statements of hf_inject_kernel with L = 1, for the SET variant of random.sample (X * Y > setsize,
// random.py:421-448; the pool variant needs X * Y integers of scratch per trajectory and keeps the warp kernel).
// Same draws in the same order, hence the same sites, flags and initial events (golden init_* arrays under emulation,
// test_thread_inject_equals_warp_inject on the device).  2.56e6 trajectories of the GA leg: 57 ms -> see DESIGN.md section 3.
__global__ void __launch_bounds__(HF_SMALL_THREADS, HF_SMALL_MIN_BLOCKS) hf_inject_small_kernel(const HfParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= P.B) return;
    const size_t stride = hf_small_thread_bytes(P.cap_c, P.cap_f, hf_xcap(P));
    HfWarp w;
    hf_small_carve(w, hf_smem + threadIdx.x * stride, P.cap_c, P.cap_f, hf_xcap(P));
    hf_load_state<1, false>(P, w, local, 0);   // zero-initialised by the host
    const int64_t molecules = (int64_t)P.X * P.Y;
    int st = HF_ST_OK;
    bool exhausted = false;
    for (int t = 0; t < 2; ++t) {                                                // 219, 225
        const int z = t == 0 ? P.Z - 1 : 0;
        for (int i = 0; i < P.C; ++i) {
            int64_t j = hf_randbelow(P, w, molecules, &exhausted);
            int cand = hf_pack((int)(j / P.Y), (int)(j % P.Y), z);
            while (hf_contains<1>(w.pos(t), i, cand, 0) && !exhausted) {
                j = hf_randbelow(P, w, molecules, &exhausted);
                cand = hf_pack((int)(j / P.Y), (int)(j % P.Y), z);
            }
            w.pos(t)[i] = cand;
        }
        w.set_n_pos(t, P.C);
    }
    if (exhausted) st = HF_ST_REPLAY_EXHAUSTED;
    for (int i = 0; i < P.C && st == HF_ST_OK; ++i)                              // 226-228
        for (int t = 0; t < 2; ++t)
            if (!hf_toggle_flag<1>(P, w, t, w.pos(t)[i], 0)) st = HF_ST_CAPACITY;
    for (int t = 0; t < 2 && st == HF_ST_OK; ++t)                                // 240-276
        for (int i = 0; i < P.C && st == HF_ST_OK; ++i)
            st = hf_gen_move_event_serial(P, w, t, w.pos(t)[i], false);
    w.injection = 2ull * (uint64_t)P.C;                                          // 118
    w.status = st;
    hf_store_state<1, false>(P, w, 0);
}
