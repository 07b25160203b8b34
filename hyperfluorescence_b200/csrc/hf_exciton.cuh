// Exciton trajectories ("Path X"): one warp per exciton, rejection-free (BKL)
// selection by a warp-shuffle inclusive prefix scan over the transfer / spin-flip / decay rates.
//
// PARITY UNPINNED: the reference has no exciton dynamics (only the event codes event.py:43-51,
// the empty lists reseau.py:233-235, the `...` stubs reseau.py:526-550, TADF.intersystem_crossing
// molecule.py:299-307, the 25/75 spin statistics molecule.py:93 and per-site S1/T1 energies
// molecule.py:182-183).  The model is specified by oracle/exciton_oracle.c (see its header and
// DESIGN.md §9); this kernel is tested against that specification, bit for bit.
//
// Channels of an exciton on site i (species a, spin s), in canonical order:
//   m < M    transfer to neighbour j = i + offset[m] (periodic x, y; open z)
//            singlet: Forster (kF[a][b] * r^-6) * min(1, ws1[j]/ws1[i])
//            triplet: Dexter  (kD[a][b] * exp(-2r/L)) * min(1, wt1[j]/wt1[i])
//   M        singlet: ISC k_isc[a]        triplet: RISC k_risc[a] * min(1, ws1[i]/wt1[i])
//   M+1      radiative decay k_rad[a][s]        M+2  non-radiative decay k_nonrad[a][s]
// ws1 / wt1 are per-site Boltzmann weights exp(-E/kT) precomputed once per lattice, with the site's
// species code stored in the two least-significant mantissa bits (one 8-byte load delivers the
// acceptor's weight AND its species; the tag moves a weight by <= 3 ulp).  The selection path then
// contains only IEEE multiplies, adds, compares and one division: the device and the oracle
// produce identical sums and therefore identical event sequences.  The neighbour table is ordered
// dz-major, dx fastest, so consecutive lanes read consecutive x: a warp's gather touches each
// (dy, dz) row of the cutoff sphere as one contiguous run.
//
// Canonical summation order (specification): channel c -> lane c % 32, slot c / 32; each lane
// sums its slots in order (S_l), one Kogge-Stone scan over the 32 lane sums gives L_l and the total
// R; the chosen lane is the first with L_l > u_sel * R, and inside it the prefix restarts from
// L_{l-1} and walks the lane's slots.  One scan per event instead of one per 32 channels.
#pragma once
#include "hf_frm.cuh"

#define HF_STREAM_XTRAJ 4u
#define HFX_SINGLET 1
#define HFX_TRIPLET 3
#define HFX_K_MOVE 1
#define HFX_K_ISC 3
#define HFX_K_FORSTER 4
#define HFX_K_DECAY 5
// tallies, same indices as HF_XT_* (include/hfkmc.h) and XT_* (oracle/exciton_oracle.c)
#define HFX_T_GENERATED 0
#define HFX_T_EVENTS 1
#define HFX_T_FORSTER 2
#define HFX_T_DEXTER 3
#define HFX_T_ISC 4
#define HFX_T_RISC 5
#define HFX_T_RAD 6
#define HFX_T_NONRAD_S 9
#define HFX_T_NONRAD_T 12
#define HFX_T_N 16

struct HfxParams {
    int32_t X, Y, Z;
    int64_t N, B, traj_per_real;
    uint32_t k0, k1, first_traj;
    int32_t M;                        // neighbour offsets
    int32_t R;                        // ceil(cutoff): sites at least R away from the x / y seams need no wrap
    const int32_t *offsets;           // [M] packed (dx+128) | (dy+128) << 8 | (dz+128) << 16
    const int32_t *lin;               // [M] dz*X*Y + dy*X + dx
    const double *g6, *gd;            // [M]
    double kf[9], kd[9];              // [donor][acceptor]
    double k_isc[3], k_risc[3], k_rad[6], k_nonrad[6];
    double singlet_fraction;
    const double *ws1, *wt1;          // [n_real][N] tagged weights
    // state [B]
    int32_t *site, *spin;
    double *time, *lifetime;
    uint64_t *draws;
    unsigned long long *totals;       // [n_real][HFX_T_N]
    HfTraceRec *trace;
    int64_t trace_first, trace_n, trace_cap;
    uint64_t *trace_len;              // [trace_n] events recorded so far
    int64_t n_events;
};

__host__ __device__ inline size_t hfx_padded_channels(int M) { return (size_t)((M + 3 + 31) & ~31); }

// per-site tagged Boltzmann weights: exp(-E / kT) with the species code in the two low mantissa bits
__device__ __forceinline__ double hfx_tag(double w, unsigned species) {
    return __longlong_as_double((__double_as_longlong(w) & ~3ll) | (long long)species);
}
__device__ __forceinline__ int hfx_tag_of(double w) { return (int)(__double_as_longlong(w) & 3ll); }

__global__ void hfx_weights_kernel(const uint8_t *species, const double *s1, const double *t1, double *ws1, double *wt1, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ws1[i] = hfx_tag(exp(-(s1[i] / HF_KT)), species[i]);
        wt1[i] = hfx_tag(exp(-(t1[i] / HF_KT)), species[i]);
    }
}

// A new exciton: uniformly random interior site, singlet with probability singlet_fraction
// (molecule.py:93).  Consumes Philox block draws/2 of the trajectory's XTRAJ stream.
__device__ __forceinline__ void hfx_spawn(const HfxParams &P, uint32_t ident, uint64_t &draws, int &site, int &spin,
                                          double &time) {
    const HfPhiloxBlock blk = hf_philox_block(P.k0, P.k1, HF_STREAM_XTRAJ, ident, draws >> 1);
    draws += 2;
    const double ua = hf_block_uniform(blk, 0), ub = hf_block_uniform(blk, 1);
    const int64_t n_interior = (int64_t)P.X * P.Y * (P.Z - 2);
    int64_t k = (int64_t)floor(ua * (double)n_interior);
    if (k >= n_interior) k = n_interior - 1;
    site = hf_pack((int)(k % P.X), (int)((k / P.X) % P.Y), 1 + (int)(k / ((int64_t)P.X * P.Y)));
    spin = ub < P.singlet_fraction ? HFX_SINGLET : HFX_TRIPLET;
    time = 0.0;
}

__global__ void hfx_spawn_kernel(const HfxParams P) {
    const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= P.B) return;
    uint64_t draws = 0;
    int site, spin;
    double time;
    hfx_spawn(P, P.first_traj + (uint32_t)local, draws, site, spin, time);
    P.site[local] = site; P.spin[local] = spin; P.time[local] = time; P.draws[local] = draws; P.lifetime[local] = 0.0;
    atomicAdd(P.totals + (local / P.traj_per_real) * HFX_T_N + HFX_T_GENERATED, 1ull);
}

__host__ __device__ inline size_t hfx_smem_bytes(int M, int warps) {
    // g6, gd (f64) | per-warp rates (f64) | per-warp donor prefactors (4 f64) | lin, off (i32) | per-warp counters (u32)
    return 8 * (2 * (size_t)M + (size_t)warps * (hfx_padded_channels(M) + 4)) + 4 * (2 * (size_t)M + (size_t)warps * HFX_T_N);
}

// Pin a value in a register: nvcc otherwise re-derives kernel parameters and pointers built from
// them inside the hot loop (LDC + IMAD chains) to save registers, which costs more issue slots than
// it saves (ncu, profiles/r01e_*).  An empty asm with a read-write operand makes the value opaque.
#if defined(__CUDA_ARCH__)
#define HFX_PIN32(x) asm volatile("" : "+r"(x))
#define HFX_PIN64(x) asm volatile("" : "+l"(x))
#define HFX_PINF64(x) asm volatile("" : "+d"(x))
#else
#define HFX_PIN32(x) (void)0
#define HFX_PIN64(x) (void)0
#define HFX_PINF64(x) (void)0
#endif

#ifndef HFX_MIN_BLOCKS
#define HFX_MIN_BLOCKS 4
#endif
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, HFX_MIN_BLOCKS) hfx_run_kernel(const HfxParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int M = P.M;
    HFX_PIN32(M);
    const int padded = (int)hfx_padded_channels(M), K = padded >> 5, Kfull = M >> 5;
    double *s_g6 = reinterpret_cast<double *>(hf_smem);
    double *s_gd = s_g6 + M;
    double *s_rates = s_gd + M + (size_t)warp * padded;
    double *s_pref = s_gd + M + (size_t)kWarps * padded + 4 * warp;            // this event's kF[a][.] or kD[a][.]
    int32_t *s_lin = reinterpret_cast<int32_t *>(s_gd + M + (size_t)kWarps * (padded + 4));
    int32_t *s_off = s_lin + M;
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_off + M) + warp * HFX_T_N;
    for (int m = threadIdx.x; m < M; m += kWarps * 32) { s_g6[m] = P.g6[m]; s_gd[m] = P.gd[m]; s_lin[m] = P.lin[m]; s_off[m] = P.offsets[m]; }
    __syncthreads();
    int plane = P.X * P.Y, dim_x = P.X, dim_y = P.Y;
    unsigned n_sites = (unsigned)P.N;
    HFX_PIN32(plane); HFX_PIN32(dim_x); HFX_PIN32(dim_y); HFX_PIN32(n_sites);

    for (int64_t local = (int64_t)blockIdx.x * kWarps + warp; local < P.B; local += (int64_t)gridDim.x * kWarps) {
        const uint32_t ident = P.first_traj + (uint32_t)local;
        const int64_t r = local / P.traj_per_real;
        const double *__restrict__ ws1 = P.ws1 + r * P.N, *__restrict__ wt1 = P.wt1 + r * P.N;
        int site = P.site[local], spin = P.spin[local];
        double time = P.time[local], lifetime = P.lifetime[local];
        uint64_t draws = P.draws[local];
        if (lane < HFX_T_N) s_cnt[lane] = 0;
        const bool traced = P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
        uint64_t trace_at = traced ? P.trace_len[local - P.trace_first] : 0;
        __syncwarp();

        for (int64_t e = 0; e < P.n_events; ++e) {
            const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
            const int i = pz * plane + py * dim_x + px;
            const bool singlet = spin == HFX_SINGLET;
            const double *w = singlet ? ws1 : wt1;
            HFX_PIN64(w);
            const double wi = __ldg(w + i);
            const int a = hfx_tag_of(wi);
            double inv_wi = __ddiv_rn(1.0, wi);
            HFX_PINF64(inv_wi);
            if (lane < 3) s_pref[lane] = singlet ? P.kf[3 * a + lane] : P.kd[3 * a + lane];
            __syncwarp();
            const bool interior = px >= P.R && px < dim_x - P.R && py >= P.R && py < dim_y - P.R;   // no periodic wrap needed
            double S = 0.0;
            {   // full slots: every lane owns a transfer channel
                const double *g = (singlet ? s_g6 : s_gd) + lane;
                const int32_t *lin = s_lin + lane, *off = s_off + lane;
                double *out = s_rates + lane;
                for (int k = 0; k < Kfull; ++k, g += 32, lin += 32, off += 32, out += 32) {
                    int j;
                    bool ok;
                    if (interior) {
                        j = i + *lin;
                        ok = (unsigned)j < n_sites;                             // open z: out of range <=> outside [0, N)
                    } else {
                        const int o = *off;
                        const int zz = pz + ((o >> 16) & 255) - 128;
                        int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                        if (xx < 0) xx += dim_x; else if (xx >= dim_x) xx -= dim_x;
                        if (yy < 0) yy += dim_y; else if (yy >= dim_y) yy -= dim_y;
                        ok = zz >= 0 && zz < P.Z;
                        j = zz * plane + yy * dim_x + xx;
                    }
                    double rate = 0.0;
                    if (ok) {
                        const double wj = __ldg(w + j);
                        rate = __dmul_rn(__dmul_rn(s_pref[hfx_tag_of(wj)], *g), fmin(__dmul_rn(wj, inv_wi), 1.0));
                    }
                    *out = rate;
                    S = __dadd_rn(S, rate);
                }
            }
            for (int k = Kfull; k < K; ++k) {   // tail slot(s): the remaining transfer channels and the three on-site ones
                const int c = (k << 5) + lane;
                double rate = 0.0;
                if (c < M) {
                    const int o = s_off[c];
                    const int zz = pz + ((o >> 16) & 255) - 128;
                    int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                    if (xx < 0) xx += P.X; else if (xx >= P.X) xx -= P.X;
                    if (yy < 0) yy += P.Y; else if (yy >= P.Y) yy -= P.Y;
                    if (zz >= 0 && zz < P.Z) {
                        const double wj = __ldg(w + (zz * plane + yy * P.X + xx));
                        rate = __dmul_rn(__dmul_rn(s_pref[hfx_tag_of(wj)], (singlet ? s_g6 : s_gd)[c]), fmin(__dmul_rn(wj, inv_wi), 1.0));
                    }
                } else if (c == M) {
                    rate = singlet ? P.k_isc[a] : __dmul_rn(P.k_risc[a], fmin(__dmul_rn(__ldg(ws1 + i), inv_wi), 1.0));   // inv_wi = 1/wt1[i] for a triplet
                } else if (c == M + 1) {
                    rate = P.k_rad[2 * a + (singlet ? 0 : 1)];
                } else if (c == M + 2) {
                    rate = P.k_nonrad[2 * a + (singlet ? 0 : 1)];
                }
                s_rates[c] = rate;
                S = __dadd_rn(S, rate);
            }
            // one inclusive Kogge-Stone scan over the 32 lane sums
            double L = S;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double up = __shfl_up_sync(HF_FULL, L, off);
                if (lane >= off) L = __dadd_rn(L, up);
            }
            const double R = __shfl_sync(HF_FULL, L, 31);
            const HfPhiloxBlock blk = hf_philox_block(P.k0, P.k1, HF_STREAM_XTRAJ, ident, draws >> 1);
            draws += 2;
            const double u_sel = hf_block_uniform(blk, 0), u_time = hf_block_uniform(blk, 1);
            const double target = __dmul_rn(u_sel, R);
            const unsigned over = __ballot_sync(HF_FULL, L > target);
            const unsigned nonzero = __ballot_sync(HF_FULL, S > 0.0);
            const int lstar = over ? __ffs(over) - 1 : 31 - __clz(nonzero);
            const double below = __shfl_sync(HF_FULL, L, (lstar + 31) & 31);
            // in-lane walk: lane k accumulates the chosen lane's rates of slots 0..k in slot order, so
            // lane k holds the running sum after slot k and one ballot finds the first that passes target
            double acc = lstar > 0 ? below : 0.0, mine = 0.0;
            {
                const double *rk = s_rates + lstar;
                for (int k = 0; k < K; ++k, rk += 32) {
                    const double v = *rk;
                    if (k <= lane) acc = __dadd_rn(acc, v);
                    if (k == lane) mine = v;
                }
            }
            const unsigned in_k = K >= 32 ? HF_FULL : ((1u << K) - 1u);
            const unsigned passed = __ballot_sync(HF_FULL, acc > target) & in_k;
            const unsigned live = __ballot_sync(HF_FULL, mine > 0.0) & in_k;
            const int slot = passed ? __ffs(passed) - 1 : 31 - __clz(live);
            const int chosen = (slot << 5) + lstar;
            const double dt = __ddiv_rn(-log(__dsub_rn(1.0, u_time)), R);
            time = __dadd_rn(time, dt);
            int kind, fin = site, tally;
            if (chosen < M) {
                const int o = s_off[chosen];
                int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                if (xx < 0) xx += P.X; else if (xx >= P.X) xx -= P.X;
                if (yy < 0) yy += P.Y; else if (yy >= P.Y) yy -= P.Y;
                fin = hf_pack(xx, yy, pz + ((o >> 16) & 255) - 128);
                kind = singlet ? HFX_K_FORSTER : HFX_K_MOVE;
                tally = singlet ? HFX_T_FORSTER : HFX_T_DEXTER;
            } else if (chosen == M) {
                kind = HFX_K_ISC;                                               // molecule.py:299-307
                tally = singlet ? HFX_T_ISC : HFX_T_RISC;
                spin = singlet ? HFX_TRIPLET : HFX_SINGLET;
            } else {
                kind = HFX_K_DECAY;
                tally = (chosen == M + 1 ? HFX_T_RAD : (singlet ? HFX_T_NONRAD_S : HFX_T_NONRAD_T)) + a;
            }
            int site_after = fin;
            if (kind == HFX_K_DECAY) {
                lifetime = __dadd_rn(lifetime, time);
                hfx_spawn(P, ident, draws, site_after, spin, time);
            }
            if (lane == 0) {
                s_cnt[HFX_T_EVENTS] += 1;
                s_cnt[tally] += 1;
                if (kind == HFX_K_DECAY) s_cnt[HFX_T_GENERATED] += 1;
                if (traced && trace_at < (uint64_t)P.trace_cap) {
                    HfTraceRec rec;
                    rec.initial = (int32_t)i; rec.final_ = (int32_t)hf_site(P.X, P.Y, fin);
                    rec.tau = dt; rec.kind = (uint8_t)kind; rec.particule = HF_PART_EXCITON;
                    rec.pad[0] = (uint8_t)spin; rec.pad[1] = 0;     // spin of the trajectory's exciton after the step
                    rec.spins = (uint32_t)chosen; rec.draws = draws;
                    P.trace[(local - P.trace_first) * P.trace_cap + trace_at] = rec;
                }
            }
            trace_at += 1;
            site = site_after;
            __syncwarp();                                                       // s_rates / s_pref are rewritten by the next event
        }
        if (lane == 0) {
            P.site[local] = site; P.spin[local] = spin; P.time[local] = time; P.lifetime[local] = lifetime; P.draws[local] = draws;
            if (traced) P.trace_len[local - P.trace_first] = trace_at;
        }
        __syncwarp();
        if (lane < HFX_T_N && s_cnt[lane]) atomicAdd(P.totals + r * HFX_T_N + lane, (unsigned long long)s_cnt[lane]);
        __syncwarp();
    }
}
