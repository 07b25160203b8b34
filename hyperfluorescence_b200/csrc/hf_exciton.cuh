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
// ws1 / wt1 are per-site Boltzmann weights exp(-E/kT) precomputed once per lattice, so the
// selection path contains only IEEE multiplies, adds, compares and one division: the device and
// the oracle produce identical prefix sums and therefore identical event sequences.  The
// neighbour table is ordered dz-major, dx fastest, so consecutive lanes read consecutive x: a
// warp's gather touches each (dy, dz) row of the cutoff sphere as one contiguous run.
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
    const int32_t *offsets;           // [M] packed (dx+128) | (dy+128) << 8 | (dz+128) << 16
    const double *g6, *gd;            // [M]
    double kf[9], kd[9];              // [donor][acceptor]
    double k_isc[3], k_risc[3], k_rad[6], k_nonrad[6];
    double singlet_fraction;
    const uint8_t *species;           // [n_real][N]
    const double *ws1, *wt1;          // [n_real][N]
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

// per-site Boltzmann weights exp(-E / kT)
__global__ void hfx_weights_kernel(const double *s1, const double *t1, double *ws1, double *wt1, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ws1[i] = exp(-(s1[i] / HF_KT));
        wt1[i] = exp(-(t1[i] / HF_KT));
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

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) hfx_run_kernel(const HfxParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int padded = (int)hfx_padded_channels(P.M);
    // shared: neighbour tables (per CTA) then one prefix array per warp
    double *s_g6 = reinterpret_cast<double *>(hf_smem);
    double *s_gd = s_g6 + P.M;
    double *s_prefix = s_gd + P.M + (size_t)warp * padded;
    int32_t *s_off = reinterpret_cast<int32_t *>(s_gd + P.M + (size_t)kWarps * padded);
    for (int m = threadIdx.x; m < P.M; m += kWarps * 32) { s_g6[m] = P.g6[m]; s_gd[m] = P.gd[m]; s_off[m] = P.offsets[m]; }
    __syncthreads();
    const int64_t plane = (int64_t)P.X * P.Y;
    const int C = P.M + 3;

    for (int64_t local = (int64_t)blockIdx.x * kWarps + warp; local < P.B; local += (int64_t)gridDim.x * kWarps) {
        const uint32_t ident = P.first_traj + (uint32_t)local;
        const int64_t r = local / P.traj_per_real;
        const uint8_t *__restrict__ species = P.species + r * P.N;
        const double *__restrict__ ws1 = P.ws1 + r * P.N, *__restrict__ wt1 = P.wt1 + r * P.N;
        int site = P.site[local], spin = P.spin[local];
        double time = P.time[local], lifetime = P.lifetime[local];
        uint64_t draws = P.draws[local];
        unsigned cnt[HFX_T_N];
#pragma unroll
        for (int k = 0; k < HFX_T_N; ++k) cnt[k] = 0;
        const bool traced = P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
        uint64_t trace_at = traced ? P.trace_len[local - P.trace_first] : 0;

        for (int64_t e = 0; e < P.n_events; ++e) {
            const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
            const int64_t i = (int64_t)pz * plane + (int64_t)py * P.X + px;
            const int a = species[i];
            const bool singlet = spin == HFX_SINGLET;
            const double *__restrict__ w = singlet ? ws1 : wt1;
            const double *s_g = singlet ? s_g6 : s_gd;
            const double wi = w[i];
            const double inv_wi = __ddiv_rn(1.0, wi);
            const double p0 = singlet ? P.kf[3 * a] : P.kd[3 * a], p1 = singlet ? P.kf[3 * a + 1] : P.kd[3 * a + 1],
                         p2 = singlet ? P.kf[3 * a + 2] : P.kd[3 * a + 2];
            double base = 0.0;
            for (int q = 0; q < padded; q += 32) {
                const int c = q + lane;
                double rate = 0.0;
                if (c < P.M) {
                    const int o = s_off[c];
                    const int zz = pz + ((o >> 16) & 255) - 128;
                    if (zz >= 0 && zz < P.Z) {
                        int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                        if (xx < 0) xx += P.X; else if (xx >= P.X) xx -= P.X;
                        if (yy < 0) yy += P.Y; else if (yy >= P.Y) yy -= P.Y;
                        const int64_t j = (int64_t)zz * plane + (int64_t)yy * P.X + xx;
                        const int b = __ldg(species + j);
                        const double wj = __ldg(w + j);
                        const double pref = b == 0 ? p0 : (b == 1 ? p1 : p2);
                        const double boltz = fmin(__dmul_rn(wj, inv_wi), 1.0);
                        rate = __dmul_rn(__dmul_rn(pref, s_g[c]), boltz);
                    }
                } else if (c == P.M) {
                    rate = singlet ? P.k_isc[a] : __dmul_rn(P.k_risc[a], fmin(__dmul_rn(ws1[i], __ddiv_rn(1.0, wt1[i])), 1.0));
                } else if (c == P.M + 1) {
                    rate = P.k_rad[2 * a + (singlet ? 0 : 1)];
                } else if (c == P.M + 2) {
                    rate = P.k_nonrad[2 * a + (singlet ? 0 : 1)];
                }
                // inclusive Kogge-Stone scan of the chunk (canonical order of the specification)
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const double up = __shfl_up_sync(HF_FULL, rate, off);
                    if (lane >= off) rate = __dadd_rn(rate, up);
                }
                s_prefix[c] = __dadd_rn(base, rate);
                base = __dadd_rn(base, __shfl_sync(HF_FULL, rate, 31));
            }
            __syncwarp();
            const double R = base;
            const HfPhiloxBlock blk = hf_philox_block(P.k0, P.k1, HF_STREAM_XTRAJ, ident, draws >> 1);
            draws += 2;
            const double u_sel = hf_block_uniform(blk, 0), u_time = hf_block_uniform(blk, 1);
            const double target = __dmul_rn(u_sel, R);
            int chosen = -1;
            for (int q = 0; q < padded && chosen < 0; q += 32) {
                const int c = q + lane;
                const unsigned m = __ballot_sync(HF_FULL, c < C && s_prefix[c] > target);
                if (m) chosen = q + __ffs(m) - 1;
            }
            if (chosen < 0) {      // rounding left u_sel * R >= P_last: the last channel with a non-zero rate
                for (int q = padded - 32; q >= 0 && chosen < 0; q -= 32) {
                    const int c = q + lane;
                    const double prev = c > 0 ? s_prefix[c - 1] : 0.0;
                    const unsigned m = __ballot_sync(HF_FULL, c < C && s_prefix[c] > prev);
                    if (m) chosen = q + 31 - __clz(m);
                }
            }
            const double dt = __ddiv_rn(-log(__dsub_rn(1.0, u_time)), R);
            time = __dadd_rn(time, dt);
            cnt[HFX_T_EVENTS] += 1;
            int kind, fin = site;
            if (chosen < P.M) {
                const int o = s_off[chosen];
                int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                if (xx < 0) xx += P.X; else if (xx >= P.X) xx -= P.X;
                if (yy < 0) yy += P.Y; else if (yy >= P.Y) yy -= P.Y;
                fin = hf_pack(xx, yy, pz + ((o >> 16) & 255) - 128);
                if (singlet) { kind = HFX_K_FORSTER; cnt[HFX_T_FORSTER] += 1; }
                else { kind = HFX_K_MOVE; cnt[HFX_T_DEXTER] += 1; }
            } else if (chosen == P.M) {
                kind = HFX_K_ISC;                                        // molecule.py:299-307
                if (singlet) { spin = HFX_TRIPLET; cnt[HFX_T_ISC] += 1; }
                else { spin = HFX_SINGLET; cnt[HFX_T_RISC] += 1; }
            } else {
                kind = HFX_K_DECAY;
                if (chosen == P.M + 1) { cnt[HFX_T_RAD] += a == 0; cnt[HFX_T_RAD + 1] += a == 1; cnt[HFX_T_RAD + 2] += a == 2; }
                else if (singlet) { cnt[HFX_T_NONRAD_S] += a == 0; cnt[HFX_T_NONRAD_S + 1] += a == 1; cnt[HFX_T_NONRAD_S + 2] += a == 2; }
                else { cnt[HFX_T_NONRAD_T] += a == 0; cnt[HFX_T_NONRAD_T + 1] += a == 1; cnt[HFX_T_NONRAD_T + 2] += a == 2; }
            }
            int site_after = fin;
            if (kind == HFX_K_DECAY) {
                lifetime = __dadd_rn(lifetime, time);
                hfx_spawn(P, ident, draws, site_after, spin, time);
                cnt[HFX_T_GENERATED] += 1;
            }
            if (traced && lane == 0 && trace_at < (uint64_t)P.trace_cap) {
                HfTraceRec rec;
                rec.initial = (int32_t)i; rec.final_ = (int32_t)hf_site(P.X, P.Y, fin);
                rec.tau = dt; rec.kind = (uint8_t)kind; rec.particule = HF_PART_EXCITON;
                rec.pad[0] = (uint8_t)spin; rec.pad[1] = 0;     // spin of the trajectory's exciton after the step
                rec.spins = (uint32_t)chosen; rec.draws = draws;
                P.trace[(local - P.trace_first) * P.trace_cap + trace_at] = rec;
            }
            trace_at += 1;
            site = site_after;
            __syncwarp();
        }
        if (lane == 0) {
            P.site[local] = site; P.spin[local] = spin; P.time[local] = time; P.lifetime[local] = lifetime; P.draws[local] = draws;
            if (traced) P.trace_len[local - P.trace_first] = trace_at;
#pragma unroll
            for (int k = 0; k < HFX_T_N; ++k)
                if (cnt[k]) atomicAdd(P.totals + r * HFX_T_N + k, (unsigned long long)cnt[k]);
        }
    }
}
