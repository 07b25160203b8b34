// emu_frm.cpp — compiles the CUDA kernel source (hf_frm.cuh) for the host under warp_emu.h and
// exposes a small C API for tests/test_kernel_emulated.py.  TEST INFRASTRUCTURE ONLY (see
// warp_emu.h): one trajectory, one warp, no product code path leads here.
#include "warp_emu.h"

#include "../../hyperfluorescence_b200/csrc/hf_frm.cuh"
#include "../../hyperfluorescence_b200/csrc/hf_frm_c1.cuh"
#include "../../hyperfluorescence_b200/csrc/hf_frm_small.cuh"
#include "../../hyperfluorescence_b200/csrc/hf_exciton.cuh"
#include "../../hyperfluorescence_b200/csrc/hf_exciton_dense.cuh"

#include <vector>

namespace {

struct Buffers {
    std::vector<int32_t> pos, flag, ev_init, ev_final, cache_i;
    std::vector<double> ev_tau, cache_tau;
    std::vector<HfScalars> sc;
    std::vector<unsigned long long> totals;
    std::vector<HfTraceRec> trace;
    std::vector<uint64_t> trace_base;
    std::vector<int32_t> pool;
    std::vector<int32_t> x_site, x_meta, x_final, x_n;
    std::vector<double> x_tau, x_clock;
    std::vector<uint64_t> x_draws;
};

void wire(HfParams &P, Buffers &b, int64_t trace_cap) {
    const int64_t B = P.B;
    b.pos.assign(2 * B * P.cap_c, 0); b.flag.assign(2 * B * P.cap_f, 0);
    b.ev_init.assign(2 * B * P.cap_c, 0); b.ev_final.assign(2 * B * P.cap_c, 0); b.ev_tau.assign(2 * B * P.cap_c, 0.0);
    b.cache_i.assign(3 * B * HF_CACHE, 0); b.cache_tau.assign(B * HF_CACHE, 0.0);
    b.sc.assign(B, HfScalars{}); b.totals.assign(32, 0);
    b.trace.assign(B * trace_cap + 1, HfTraceRec{});
    for (int t = 0; t < 2; ++t) {
        P.pos[t] = b.pos.data() + t * B * P.cap_c;
        P.flag[t] = b.flag.data() + t * B * P.cap_f;
        P.ev_init[t] = b.ev_init.data() + t * B * P.cap_c;
        P.ev_final[t] = b.ev_final.data() + t * B * P.cap_c;
        P.ev_tau[t] = b.ev_tau.data() + t * B * P.cap_c;
    }
    P.sc = b.sc.data();
    P.cache_init = b.cache_i.data(); P.cache_final = b.cache_i.data() + B * HF_CACHE; P.cache_kp = b.cache_i.data() + 2 * B * HF_CACHE;
    P.cache_tau = b.cache_tau.data();
    P.totals = b.totals.data();
    b.trace_base.assign(B, 0); P.trace_base = b.trace_base.data();
    P.trace = b.trace.data(); P.trace_first = 0; P.trace_n = B; P.trace_cap = trace_cap;
    if (P.xi.mode) {
        b.x_site.assign(B * P.cap_c, 0); b.x_meta.assign(B * P.cap_c, 0); b.x_final.assign(B * P.cap_c, 0);
        b.x_tau.assign(B * P.cap_c, 0.0); b.x_n.assign(B, 0); b.x_clock.assign(B, 0.0); b.x_draws.assign(B, 0);
        P.xi.x_site = b.x_site.data(); P.xi.x_meta = b.x_meta.data(); P.xi.x_final = b.x_final.data();
        P.xi.x_tau = b.x_tau.data(); P.xi.x_n = b.x_n.data(); P.xi.x_clock = b.x_clock.data(); P.xi.x_draws = b.x_draws.data();
    }
}

// Integrated exciton model of an emulated run (P.xi): mode 0 / 1 need nothing else; mode 2 takes the tables and
// the haloed, tagged weights exactly as hf_x_configure would have built them on the device.
struct EmuXModel {
    int32_t mode, M;
    const int32_t *offsets; const double *g6, *gd, *kf, *kd, *k_isc, *k_risc, *k_rad, *k_nonrad, *ws1, *wt1;
    int64_t w_stride, w_halo;
};
HfiTables g_emu_tab;
void apply_xmodel(HfParams &P, const EmuXModel *m) {
    memset(&P.xi, 0, sizeof(P.xi));
    if (!m) return;
    P.xi.mode = m->mode;
    if (m->mode != 2) return;
    HfiTables &T = g_emu_tab;
    memset(&T, 0, sizeof(T));
    T.M = m->M; T.X = P.X; T.Y = P.Y; T.Z = P.Z; T.k0 = P.k0; T.k1 = P.k1;
    T.offsets = m->offsets; T.g6 = m->g6; T.gd = m->gd;
    memcpy(T.k, m->kf, 72); memcpy(T.k + 9, m->kd, 72); memcpy(T.k + 18, m->k_isc, 24); memcpy(T.k + 21, m->k_risc, 24);
    memcpy(T.k + 24, m->k_rad, 48); memcpy(T.k + 30, m->k_nonrad, 48);
    T.ws1 = m->ws1; T.wt1 = m->wt1; T.w_stride = m->w_stride; T.w_halo = m->w_halo;
    P.xi.tab = &T;
}

struct InjectArgs { const HfParams *P; int32_t *pool; };
void inject_body(void *a) { auto *x = (InjectArgs *)a; hf_inject_kernel<1>(*x->P, x->pool); }
void run_body(void *a) { const HfParams &P = *(const HfParams *)a; if (P.xi.mode) hf_run_kernel<1, true>(P); else hf_run_kernel<1, false>(P); }
void run_c1_body(void *a) { hf_run_c1_kernel<HF_C1_THREADS, HF_C1_MIN_BLOCKS>(*(const HfParams *)a); }
void inject_c1_body(void *a) { hf_inject_c1_kernel(*(const HfParams *)a); }
void inject_small_body(void *a) { hf_inject_small_kernel(*(const HfParams *)a); }
void run_small_body(void *a) { const HfParams &P = *(const HfParams *)a; if (P.xi.mode) hf_run_small_kernel<true>(P); else hf_run_small_kernel<false>(P); }

struct ShuffleArgs { const HfParams *P; int32_t *pool; int64_t total, n_host, n_tadf; uint8_t *layer; };
void shuffle_body(void *a) { auto *x = (ShuffleArgs *)a; hf_species_shuffle(*x->P, x->pool, x->total, &x->n_host, &x->n_tadf, x->layer, 0); }
struct FillArgs { const HfParams *P; const uint8_t *layer; uint8_t *species; double *homo, *lumo, *s1, *t1; };
void fill_body(void *a) { auto *x = (FillArgs *)a; hf_fill_sites(*x->P, x->layer, 1, x->species, x->homo, x->lumo, x->s1, x->t1); }

int64_t sample_setsize(int64_t k) {
    int64_t setsize = 21;
    if (k > 5) {
        const int64_t e = (int64_t)ceil(log((double)(k * 3)) / log(4.0));
        int64_t p4 = 1;
        for (int64_t i = 0; i < e; ++i) p4 *= 4;
        setsize += p4;
    }
    return setsize;
}

}  // namespace

extern "C" {

// Lattice generation kernels for one realisation.
int emu_generate(const int32_t dims[3], const double props[3], uint64_t seed, uint32_t realisation,
                 uint8_t *species, double *homo, double *lumo, double *s1, double *t1) {
    HfParams P;
    memset(&P, 0, sizeof(P));
    P.X = dims[0]; P.Y = dims[1]; P.Z = dims[2]; P.N = (int64_t)P.X * P.Y * P.Z;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); hf_philox_round_keys(P.k0, P.k1, P.rk); P.first_real = realisation;
    const int64_t xy = (int64_t)P.X * P.Y;
    const int64_t n_fluo = (int64_t)((double)P.N * props[2]), n_tadf = (int64_t)((double)P.N * props[1]);
    const int64_t n_host = P.N - 2 * xy - n_fluo - n_tadf, total = n_host + n_tadf + n_fluo;
    if (total != xy * (P.Z - 2) || total <= 0) return 1;
    std::vector<int32_t> pool((size_t)total);
    for (int64_t i = 0; i < total; ++i) pool[(size_t)i] = (int32_t)i;
    std::vector<uint8_t> layer((size_t)xy);
    blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
    ShuffleArgs sa{&P, pool.data(), total, n_host, n_tadf, layer.data()};
    emu_run_warp(shuffle_body, &sa);
    FillArgs fa{&P, layer.data(), species, homo, lumo, s1, t1};
    emu_run_warp(fill_body, &fa);
    return 0;
}

// hf_inject + n_calls x hf_run(stop) for one trajectory on a given lattice.
// ints_out:  n_e, n_h, n_fe, n_fh, n_me, n_mh, special, special_site(packed), xstate, status, cache_head, cache_n
// u64_out:   draws, spins, emission, recombination, injection, steps, fired, init_draws
// lists_out: 6 consecutive blocks of cap ints (packed coordinates): e_pos, h_pos, me_init, mh_init,
//            me_final, mh_final; flags_out: 2 blocks of cap_f; taus_out: 2 blocks of cap doubles.
int emu_run(const int32_t dims[3], double field, int32_t charges, uint64_t seed, uint32_t trajectory,
            const uint8_t *species, const double *homo, const double *lumo,
            const double *replay, int64_t n_replay, int64_t stop, int32_t n_calls, int32_t small_kernel, int32_t distance,
            HfTraceRec *trace_out, int64_t trace_cap,
            int32_t *ints_out, uint64_t *u64_out, int32_t *lists_out, int32_t *flags_out, double *taus_out,
            unsigned long long *totals_out, int32_t *cache_i_out, double *cache_tau_out,
            int32_t *init_lists_out, double *init_taus_out, int32_t *init_counts_out,
            const EmuXModel *xmodel, int32_t *x_ints_out /* [3][cap]: site, meta, final */, double *x_tau_out /* [cap] */,
            double *x_scalars_out /* n_x, draws, clock */) {
    HfParams P;
    memset(&P, 0, sizeof(P));
    P.X = dims[0]; P.Y = dims[1]; P.Z = dims[2]; P.N = (int64_t)P.X * P.Y * P.Z;
    P.field = field; P.C = charges; P.cap_c = charges; P.cap_f = (charges == 1 && distance == 1) ? 2 : 2 * charges + 8;
    P.dist = distance;
    P.B = 1; P.traj_per_real = 1;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); hf_philox_round_keys(P.k0, P.k1, P.rk); P.first_traj = trajectory; P.first_real = 0;
    P.setsize = sample_setsize(charges);
    P.species = species; P.homo = homo; P.lumo = lumo;
    P.replay = replay; P.n_replay = n_replay;
    apply_xmodel(P, xmodel);
    Buffers b;
    wire(P, b, trace_cap);
    if (hf_warp_smem_bytes(P.cap_c, P.cap_f, hf_xcap(P)) > sizeof(hf_smem)) return 2;
    blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
    b.pool.assign((size_t)((int64_t)P.X * P.Y), 0);
    InjectArgs ia{&P, b.pool.data()};
    // the product's choice: thread-per-trajectory injection with the thread kernels when random.sample takes its set variant
    const bool small_inject = small_kernel && charges >= 1 && charges <= HF_SMALL_MAX_CHARGES && distance == 1 &&
                              (int64_t)P.X * P.Y > P.setsize && hf_small_thread_bytes(P.cap_c, P.cap_f, hf_xcap(P)) * 32 <= sizeof(hf_smem);
    if (small_inject) emu_run_warp(inject_small_body, &P);
    else emu_run_warp(inject_body, &ia);
    const int cap = P.cap_c;
    init_counts_out[0] = b.sc[0].n_pos[0]; init_counts_out[1] = b.sc[0].n_pos[1];
    init_counts_out[2] = b.sc[0].n_ev[0]; init_counts_out[3] = b.sc[0].n_ev[1];
    init_counts_out[4] = b.sc[0].status;
    u64_out[7] = b.sc[0].draws;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < cap; ++i) {
            init_lists_out[(0 + t) * cap + i] = P.pos[t][i];
            init_lists_out[(2 + t) * cap + i] = P.ev_init[t][i];
            init_lists_out[(4 + t) * cap + i] = P.ev_final[t][i];
            init_taus_out[t * cap + i] = P.ev_tau[t][i];
        }
    if (small_kernel && (charges < 1 || charges > HF_SMALL_MAX_CHARGES || hf_small_thread_bytes(P.cap_c, P.cap_f, hf_xcap(P)) * 32 > sizeof(hf_smem))) return 4;
    for (int c = 0; c < n_calls; ++c) {
        P.n_steps = stop;
        emu_run_warp(small_kernel ? run_small_body : run_body, &P);      // thread kernel: lane 0 owns the trajectory
    }
    const HfScalars &sc = b.sc[0];
    const int32_t ints[12] = {sc.n_pos[0], sc.n_pos[1], sc.n_flag[0], sc.n_flag[1], sc.n_ev[0], sc.n_ev[1],
                              sc.special, sc.special_site, sc.xstate, sc.status, sc.cache_head, sc.cache_n};
    memcpy(ints_out, ints, sizeof(ints));
    const uint64_t u64s[7] = {sc.draws, sc.spins, sc.emission, sc.recombination, sc.injection, sc.steps, sc.fired};
    memcpy(u64_out, u64s, sizeof(u64s));
    for (int t = 0; t < 2; ++t) {
        for (int i = 0; i < cap; ++i) {
            lists_out[(0 + t) * cap + i] = P.pos[t][i];
            lists_out[(2 + t) * cap + i] = P.ev_init[t][i];
            lists_out[(4 + t) * cap + i] = P.ev_final[t][i];
            taus_out[t * cap + i] = P.ev_tau[t][i];
        }
        for (int i = 0; i < P.cap_f; ++i) flags_out[t * P.cap_f + i] = P.flag[t][i];
    }
    memcpy(totals_out, b.totals.data(), sizeof(unsigned long long) * 24);
    memcpy(cache_i_out, b.cache_i.data(), sizeof(int32_t) * 3 * HF_CACHE);
    memcpy(cache_tau_out, b.cache_tau.data(), sizeof(double) * HF_CACHE);
    if (P.xi.mode && x_ints_out) {
        for (int i = 0; i < cap; ++i) {
            x_ints_out[i] = b.x_site[i]; x_ints_out[cap + i] = b.x_meta[i]; x_ints_out[2 * cap + i] = b.x_final[i];
            x_tau_out[i] = b.x_tau[i];
        }
        x_scalars_out[0] = (double)b.x_n[0]; x_scalars_out[1] = (double)b.x_draws[0]; x_scalars_out[2] = b.x_clock[0];
    }
    const int64_t n = (int64_t)sc.fired < trace_cap ? (int64_t)sc.fired : trace_cap;
    memcpy(trace_out, b.trace.data(), sizeof(HfTraceRec) * (size_t)n);
    return 0;
}

// Batch of B <= 32 trajectories (global ids first_traj ..) with charges == 1 on one lattice:
// hf_inject, then n_calls x hf_run(stop) with either the thread-per-trajectory kernel (use_c1, each
// emulated lane owns one trajectory) or the warp kernel (one trajectory after the other).
// Outputs per trajectory b: trace_out[b*trace_cap ..], sc_out[b] (raw HfScalars), and the raw
// global arrays pos/ev_init/ev_final ([2][B]), ev_tau ([2][B]), flag ([2][B][2]), cache ([3][10][B], tau [10][B]).
int emu_run_batch_c1(const int32_t dims[3], double field, uint64_t seed, uint32_t first_traj, int32_t B, int32_t use_c1,
                     const uint8_t *species, const double *homo, const double *lumo, int64_t stop, int32_t n_calls,
                     HfTraceRec *trace_out, int64_t trace_cap, HfScalars *sc_out, int32_t *pos_out, int32_t *ev_init_out,
                     int32_t *ev_final_out, double *ev_tau_out, int32_t *flag_out, int32_t *cache_i_out,
                     double *cache_tau_out, unsigned long long *totals_out) {
    if (B < 1 || B > 32) return 3;
    HfParams P;
    memset(&P, 0, sizeof(P));
    P.X = dims[0]; P.Y = dims[1]; P.Z = dims[2]; P.N = (int64_t)P.X * P.Y * P.Z;
    P.field = field; P.C = 1; P.cap_c = 1; P.cap_f = 2; P.dist = 1;
    P.B = B; P.traj_per_real = B;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); hf_philox_round_keys(P.k0, P.k1, P.rk); P.first_traj = first_traj; P.first_real = 0;
    P.setsize = sample_setsize(1);
    P.species = species; P.homo = homo; P.lumo = lumo;
    Buffers b;
    wire(P, b, trace_cap);
    blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
    b.pool.assign((size_t)((int64_t)P.X * P.Y), 0);
    InjectArgs ia{&P, b.pool.data()};
    if (use_c1) emu_run_warp(inject_c1_body, &P);                  // thread-per-trajectory injection, like the product path
    else emu_run_warp(inject_body, &ia);
    for (int c = 0; c < n_calls; ++c) {
        P.n_steps = stop;
        emu_run_warp(use_c1 ? run_c1_body : run_body, &P);
    }
    memcpy(trace_out, b.trace.data(), sizeof(HfTraceRec) * (size_t)(B * trace_cap));
    memcpy(sc_out, b.sc.data(), sizeof(HfScalars) * (size_t)B);
    memcpy(pos_out, b.pos.data(), sizeof(int32_t) * 2 * (size_t)B);
    memcpy(ev_init_out, b.ev_init.data(), sizeof(int32_t) * 2 * (size_t)B);
    memcpy(ev_final_out, b.ev_final.data(), sizeof(int32_t) * 2 * (size_t)B);
    memcpy(ev_tau_out, b.ev_tau.data(), sizeof(double) * 2 * (size_t)B);
    memcpy(flag_out, b.flag.data(), sizeof(int32_t) * 4 * (size_t)B);
    memcpy(cache_i_out, b.cache_i.data(), sizeof(int32_t) * 3 * HF_CACHE * (size_t)B);
    memcpy(cache_tau_out, b.cache_tau.data(), sizeof(double) * HF_CACHE * (size_t)B);
    memcpy(totals_out, b.totals.data(), sizeof(unsigned long long) * 20);
    return 0;
}

static void x_spawn_body(void *a) { hfx_spawn_kernel(*(const HfxParams *)a); }
// the same two variants the library picks (x_pick_kernel): register tables when M / 32 is instantiated, generic otherwise
static void x_run_body_generic(void *a) { hfx_run_kernel<1, 0, -1, 1, false>(*(const HfxParams *)a); }
static void x_run_body_r2(void *a) { hfx_run_kernel<1, 2, -1, 1, true>(*(const HfxParams *)a); }           // run-time M variants
static void x_run_body_r5(void *a) { hfx_run_kernel<1, 5, -1, 1, true>(*(const HfxParams *)a); }
static void x_run_body_r12(void *a) { hfx_run_kernel<1, 12, -1, 1, false, 6>(*(const HfxParams *)a); }
static void x_run_body_r16(void *a) { hfx_run_kernel<1, 16, -1, 1, false, 8>(*(const HfxParams *)a); }
static void x_run_body_k3(void *a) { hfx_run_kernel<1, 3, 26, 1, true>(*(const HfxParams *)a); }
static void x_run_body_k8(void *a) { hfx_run_kernel<1, 8, 0, 1, true>(*(const HfxParams *)a); }
static void x_run_body_k16(void *a) { hfx_run_kernel<1, 16, 2, 1, false, 8>(*(const HfxParams *)a); }
static void x_run_body_b3(void *a) { hfx_run_kernel<1, 3, 26, 1, true, 3, true>(*(const HfxParams *)a); }  // brick-layout weights
static void x_run_body_b8(void *a) { hfx_run_kernel<1, 8, 0, 1, true, 8, true>(*(const HfxParams *)a); }
static void x_run_body_b16(void *a) { hfx_run_kernel<1, 16, 2, 1, false, 8, true>(*(const HfxParams *)a); }

// Exciton path: B <= 32 trajectories on one lattice, hfx_spawn + n_calls x hfx_run(n_events).
int emu_x_run(const int32_t dims[3], const uint8_t *species, const double *ws1, const double *wt1, int32_t M,
              const int8_t *offsets, const double *g6, const double *gd, const double *kf, const double *kd,
              const double *k_isc, const double *k_risc, const double *k_rad, const double *k_nonrad,
              double singlet_fraction, uint64_t seed, uint32_t first_traj, int32_t B, int64_t n_events, int32_t n_calls,
              int32_t generic, HfTraceRec *trace_out, int64_t trace_cap, int32_t *site_out, int32_t *spin_out, double *time_out,
              double *life_out, uint64_t *draws_out, unsigned long long *totals_out) {
    if (B < 1 || B > 32) return 3;
    HfxParams P;
    memset(&P, 0, sizeof(P));
    P.X = dims[0]; P.Y = dims[1]; P.Z = dims[2]; P.N = (int64_t)P.X * P.Y * P.Z; P.B = B; P.traj_per_real = B;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); P.first_traj = first_traj;
    P.M = M; P.block = 32;
    std::vector<int32_t> packed((size_t)M);
    for (int m = 0; m < M; ++m) packed[(size_t)m] = (offsets[3 * m] + 128) | ((offsets[3 * m + 1] + 128) << 8) | ((offsets[3 * m + 2] + 128) << 16);
    std::vector<int32_t> lin((size_t)M);
    int maxr = 0;
    for (int m = 0; m < M; ++m) {
        lin[(size_t)m] = (int32_t)((int64_t)offsets[3 * m + 2] * P.X * P.Y + (int64_t)offsets[3 * m + 1] * P.X + offsets[3 * m]);
        for (int k = 0; k < 3; ++k) { const int v = offsets[3 * m + k] < 0 ? -offsets[3 * m + k] : offsets[3 * m + k]; if (v > maxr) maxr = v; }
    }
    P.R = maxr; P.lin = lin.data();
    P.offsets = packed.data(); P.g6 = g6; P.gd = gd;
    memcpy(P.kf, kf, sizeof(P.kf)); memcpy(P.kd, kd, sizeof(P.kd));
    memcpy(P.k_isc, k_isc, sizeof(P.k_isc)); memcpy(P.k_risc, k_risc, sizeof(P.k_risc));
    memcpy(P.k_rad, k_rad, sizeof(P.k_rad)); memcpy(P.k_nonrad, k_nonrad, sizeof(P.k_nonrad));
    P.singlet_fraction = singlet_fraction;
    (void)species;
    // the device layout: R zero planes below and above the lattice
    P.w_halo = (int64_t)maxr * P.X * P.Y; P.w_stride = P.N + 2 * P.w_halo;
    std::vector<double> hs1((size_t)P.w_stride, 0.0), ht1((size_t)P.w_stride, 0.0);
    const bool brick = generic == 2;                       // the padded 4 x 4 x 1 brick layout of the weights (compiled cutoffs only)
    if (brick) {
        if (M != 122 && M != 256 && M != 514) return 4;
        const int R = maxr, bX = (P.X + 2 * R + 3) & ~3, bY = (P.Y + 2 * R + 3) & ~3;
        P.brick = 1; P.bX = bX; P.bY = bY;
        P.w_halo = (int64_t)R * bX * bY; P.w_stride = (int64_t)P.Z * bX * bY + 2 * P.w_halo;
        hs1.assign((size_t)P.w_stride, 0.0); ht1.assign((size_t)P.w_stride, 0.0);
        for (int z = 0; z < P.Z; ++z)
            for (int yp = 0; yp < bY; ++yp)
                for (int xp = 0; xp < bX; ++xp) {
                    const int x = ((xp - R) % P.X + P.X) % P.X, y = ((yp - R) % P.Y + P.Y) % P.Y;
                    const int64_t src = ((int64_t)z * P.Y + y) * P.X + x, o = P.w_halo + (int64_t)z * bX * bY + hfx_brick2(xp, yp, bX);
                    hs1[(size_t)o] = ws1[src]; ht1[(size_t)o] = wt1[src];
                }
        for (int m = 0; m < M; ++m) lin[(size_t)m] = hfx_brick_lin(offsets[3 * m], offsets[3 * m + 1], offsets[3 * m + 2], bX, bY);
    } else {
        memcpy(hs1.data() + P.w_halo, ws1, sizeof(double) * (size_t)P.N);
        memcpy(ht1.data() + P.w_halo, wt1, sizeof(double) * (size_t)P.N);
    }
    P.ws1 = hs1.data(); P.wt1 = ht1.data();
    std::vector<int32_t> site((size_t)B), spin((size_t)B);
    std::vector<double> time((size_t)B), life((size_t)B);
    std::vector<uint64_t> draws((size_t)B), tlen((size_t)B, 0);
    std::vector<unsigned long long> totals(HFX_T_N, 0);
    std::vector<HfTraceRec> trace((size_t)(B * trace_cap + 1));
    P.site = site.data(); P.spin = spin.data(); P.time = time.data(); P.lifetime = life.data(); P.draws = draws.data();
    P.totals = totals.data();
    P.trace = trace.data(); P.trace_first = 0; P.trace_n = B; P.trace_cap = trace_cap; P.trace_len = tlen.data();
    if (hfx_smem_bytes(M, 1) > sizeof(hf_smem)) return 2;
    blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
    emu_run_warp(x_spawn_body, &P);
    void (*body)(void *) = brick ? (M == 122 ? x_run_body_b3 : M == 256 ? x_run_body_b8 : x_run_body_b16) :
                           generic ? x_run_body_generic : M == 122 ? x_run_body_k3 : M == 256 ? x_run_body_k8 : M == 514 ? x_run_body_k16 :
                           (M >> 5) == 2 ? x_run_body_r2 : (M >> 5) == 5 ? x_run_body_r5 : (M >> 5) == 12 ? x_run_body_r12 :
                           (M >> 5) >= 16 ? x_run_body_r16 : x_run_body_generic;
    for (int c = 0; c < n_calls; ++c) { P.n_events = n_events; emu_run_warp(body, &P); }
    memcpy(trace_out, trace.data(), sizeof(HfTraceRec) * (size_t)(B * trace_cap));
    for (int b = 0; b < B; ++b) {
        const int p = site[(size_t)b];
        site_out[b] = (int32_t)((((int64_t)(p >> 20)) * P.Y + ((p >> 10) & 1023)) * P.X + (p & 1023));
        spin_out[b] = spin[(size_t)b]; time_out[b] = time[(size_t)b]; life_out[b] = life[(size_t)b]; draws_out[b] = draws[(size_t)b];
    }
    memcpy(totals_out, totals.data(), sizeof(unsigned long long) * HFX_T_N);
    return 0;
}

// Exhaustive check of the packed brick offsets (hfx_brick_lin) against the brick addresses themselves: every offset
// with |d| <= r_max, every site of a bX x bY plane whose neighbour stays inside it.  Returns the number of mismatches.
int emu_brick_selftest(int32_t bX, int32_t bY, int32_t r_max) {
    int bad = 0;
    const int plane = bX * bY;
    for (int dz = -r_max; dz <= r_max; ++dz)
        for (int dy = -r_max; dy <= r_max; ++dy)
            for (int dx = -r_max; dx <= r_max; ++dx) {
                const int32_t t = hfx_brick_lin(dx, dy, dz, bX, bY);
                for (int y = r_max; y < bY - r_max; ++y)
                    for (int x = r_max; x < bX - r_max; ++x) {
                        int d = t >> 8;                                  // what hfx_gather<.., kBrick> computes
                        if (t & (1 << (x & 3))) d += 12;
                        if (t & (16 << (y & 3))) d += 4 * bX - 16;
                        const int want = dz * plane + hfx_brick2(x + dx, y + dy, bX) - hfx_brick2(x, y, bX);
                        bad += d != want;
                    }
            }
    return bad;
}

static void xd_run_body(void *a) { hfxd_run_kernel<1>(*(const HfxdParams *)a); }

// High-density devices: ONE device (one warp), fill + n_calls x hfxd_run(n_events).
int emu_xd_run(const int32_t dims[3], const double *ws1, const double *wt1, int32_t M, const int8_t *offsets, const double *g6,
               const double *gd, const double *kf, const double *kd, const double *k_isc, const double *k_risc,
               const double *k_rad, const double *k_nonrad, double singlet_fraction, double cutoff, const double *ann,
               double p_tta, uint64_t seed, uint32_t device, int32_t n_excitons, int64_t n_events, int32_t n_calls,
               HfTraceRec *trace_out, int64_t trace_cap, int32_t *site_out, int32_t *spin_out, double *total_out,
               double *scalars_out, unsigned long long *totals_out, uint8_t *occ_out) {
    HfxdParams D;
    memset(&D, 0, sizeof(D));
    HfxParams &P = D.X;
    P.X = dims[0]; P.Y = dims[1]; P.Z = dims[2]; P.N = (int64_t)P.X * P.Y * P.Z; P.B = 1; P.traj_per_real = 1;
    P.k0 = (uint32_t)seed; P.k1 = (uint32_t)(seed >> 32); P.first_traj = device;
    P.M = M;
    std::vector<int32_t> packed((size_t)M), lin((size_t)M);
    int maxr = 0;
    for (int m = 0; m < M; ++m) {
        packed[(size_t)m] = (offsets[3 * m] + 128) | ((offsets[3 * m + 1] + 128) << 8) | ((offsets[3 * m + 2] + 128) << 16);
        lin[(size_t)m] = (int32_t)((int64_t)offsets[3 * m + 2] * P.X * P.Y + (int64_t)offsets[3 * m + 1] * P.X + offsets[3 * m]);
        for (int k = 0; k < 3; ++k) { const int v = offsets[3 * m + k] < 0 ? -offsets[3 * m + k] : offsets[3 * m + k]; if (v > maxr) maxr = v; }
    }
    P.R = maxr; P.lin = lin.data(); P.offsets = packed.data(); P.g6 = g6; P.gd = gd;
    memcpy(P.kf, kf, sizeof(P.kf)); memcpy(P.kd, kd, sizeof(P.kd));
    memcpy(P.k_isc, k_isc, sizeof(P.k_isc)); memcpy(P.k_risc, k_risc, sizeof(P.k_risc));
    memcpy(P.k_rad, k_rad, sizeof(P.k_rad)); memcpy(P.k_nonrad, k_nonrad, sizeof(P.k_nonrad));
    P.singlet_fraction = singlet_fraction;
    P.w_halo = (int64_t)maxr * P.X * P.Y; P.w_stride = P.N + 2 * P.w_halo;
    std::vector<double> hs1((size_t)P.w_stride, 0.0), ht1((size_t)P.w_stride, 0.0);
    memcpy(hs1.data() + P.w_halo, ws1, sizeof(double) * (size_t)P.N);
    memcpy(ht1.data() + P.w_halo, wt1, sizeof(double) * (size_t)P.N);
    P.ws1 = hs1.data(); P.wt1 = ht1.data();
    D.n = n_excitons; D.n_pad = (n_excitons + 31) & ~31;
    memcpy(D.ann, ann, sizeof(D.ann));
    D.p_tta = p_tta; (void)cutoff;
    D.occ_words = (P.w_stride + 15) / 16;
    std::vector<uint32_t> pos((size_t)D.n_pad, 0), occ((size_t)D.occ_words, 0);
    std::vector<double> T((size_t)D.n_pad, 0.0), birth((size_t)D.n_pad, 0.0);
    double time = 0.0, life = 0.0;
    uint64_t draws = 0, tlen = 0;
    std::vector<unsigned long long> totals(HFXD_T_N, 0);
    std::vector<HfTraceRec> trace((size_t)(trace_cap + 1));
    std::vector<double> rec_L((size_t)D.n_pad * 32, 0.0);
    std::vector<uint32_t> rec_seen((size_t)D.n_pad * 32, 0);
    D.rec_L = rec_L.data(); D.rec_seen = rec_seen.data();
    D.pos = pos.data(); D.T = T.data(); D.birth = birth.data(); D.occ = occ.data();
    D.time = &time; D.life = &life; D.draws = &draws; D.totals = totals.data();
    P.trace = trace.data(); P.trace_first = 0; P.trace_n = 1; P.trace_cap = trace_cap; P.trace_len = &tlen;
    if (hfxd_smem_bytes(M, D.n_pad, 1) > sizeof(hf_smem)) return 2;
    blockDim.x = 32; gridDim.x = 1; blockIdx.x = 0;
    D.fill = 1; P.n_events = 0;
    emu_run_warp(xd_run_body, &D);
    D.fill = 0;
    for (int c = 0; c < n_calls; ++c) { P.n_events = n_events; emu_run_warp(xd_run_body, &D); }
    memcpy(trace_out, trace.data(), sizeof(HfTraceRec) * (size_t)trace_cap);
    for (int k = 0; k < n_excitons; ++k) {
        const int p = (int)(pos[(size_t)k] & 0x3fffffffu);
        site_out[k] = (int32_t)((((int64_t)(p >> 20)) * P.Y + ((p >> 10) & 1023)) * P.X + (p & 1023));
        spin_out[k] = (int32_t)(pos[(size_t)k] >> 30);
        total_out[k] = T[(size_t)k];
    }
    scalars_out[0] = time; scalars_out[1] = life; scalars_out[2] = (double)draws;
    memcpy(totals_out, totals.data(), sizeof(unsigned long long) * HFXD_T_N);
    for (int64_t i = 0; i < P.N; ++i) {
        const int64_t idx = i + P.w_halo;
        occ_out[i] = (uint8_t)((occ[(size_t)(idx >> 4)] >> (2 * (int)(idx & 15))) & 3u);
    }
    return 0;
}

}  // extern "C"
