// High-density exciton devices ("Path XD", SURVEY.md §8 f4, BASELINE configs[3]): n interacting excitons
// per device, a 2-bit-per-site occupancy mask, annihilation on occupied acceptors, one warp per device.
//
// PARITY UNPINNED: the reference has neither exciton dynamics nor annihilation; its only occupancy
// rule is same-type charge blocking (reseau.py:385, 394).  The model is specified by
// oracle/exciton_dense_oracle.c (read its header first: channels, canonical sums, selection rule,
// replacement rule); this kernel is tested against it bit for bit.
//
// Device state lives in global memory between launches and in the warp's shared memory during one:
//   pos[n_pad]   u32  packed x | y << 10 | z << 20 | spin << 30 (spin code 1 singlet / 3 triplet, molecule.py:16-21)
//   T[n_pad]     f64  cached total rate of every exciton (0 for the padding slots)
//   birth[n_pad] f64  device clock when the exciton was created
//   occ[]        u32  16 sites per word, 2 bits per site, with the same R-plane halo as the weights
//   rec[n_pad]   32 x (f64 L, u32 seen) per exciton, global: the scanned lane sums of its channel rates at its last
//                evaluation and, per lane, which of the lane's channels pointed at occupied sites (bit 31: lane sum > 0)
// Per event the warp (1) scans the cached totals to pick an exciton, (2) picks that exciton's channel from its
// RECORD — the 32 scanned lane sums name the lane, and only that lane's K channels are evaluated again, by K lanes
// in parallel, bit for bit as at the last evaluation (nothing the exciton can see has changed since: that is what
// keeps its cached total valid) — (3) applies it — lane 0 edits the occupancy words — and (4) re-evaluates the
// totals (M + 3 channels, lane = channel % 32; each lane also fetches the acceptor's occupancy code) and records of
// the excitons that can see a changed site.  Those are found without scanning the device: every evaluation of an
// exciton ON a changed site (the chosen exciton before the event — its record —, the same slot after it, a promoted
// acceptor) knows the occupancy codes of all sites in its cutoff sphere, so the lanes that met an occupied site
// name the neighbours; only those are looked up in the position list.  One full evaluation per event instead of two
// (round 2: 2 316 -> see profiles/ warp instructions per event).
#pragma once
#include "hf_exciton.cuh"

#define HF_STREAM_XDENSE 5u
#define HFXD_K_ANNIHILATION 6          // the `unbound` slot of event.py:49
#define HFXD_T_ANN 16                  // + 2 * donor spin index + acceptor spin index
#define HFXD_T_TTA_UP 20
#define HFXD_T_N 24
#define HFXD_MAX_EXCITONS 1024
#ifndef HFXD_MIN_BLOCKS
#define HFXD_MIN_BLOCKS 7              // 7 CTAs x 4 warps = 28 devices per SM: the 4096 devices of BASELINE configs[3] are ONE wave on 148 SMs
#endif
#define HFXD_SCRATCH 32                // doubles of per-warp scratch: the rates of one lane's channels
#ifndef HFXD_PIN_RECORDS
#define HFXD_PIN_RECORDS 0                 // measured: pinning the two record pointers costs more (registers) than the multiplies it saves: 3.79e8 -> 3.66e8
#endif
#ifndef HFXD_CHUNK
#define HFXD_CHUNK 3                   // slots whose gathers (weight + occupancy word) are in flight together
#endif

struct HfxdParams {
    HfxParams X;                       // lattice, tables, rate constants, Philox key; B = devices, first_traj = first device id
    int32_t n, n_pad;                  // excitons per device, rounded up to 32
    double ann[4];                     // [donor spin index][acceptor spin index]
    double p_tta;                      // triplet-triplet promotion probability
    uint32_t *pos;                     // [B][n_pad]
    double *T, *birth;                 // [B][n_pad]
    double *rec_L;                     // [B][n_pad][32] scanned lane sums of every exciton's last evaluation
    uint32_t *rec_seen;                // [B][n_pad][32] occupied-acceptor bits per lane | (lane sum > 0) << 31
    uint32_t *occ;                     // [B][occ_words]
    int64_t occ_words;
    double *time, *life;               // [B] device clock, summed lifetimes of the lost excitons
    uint64_t *draws;                   // [B]
    unsigned long long *totals;        // [n_real][HFXD_T_N]
    int32_t fill;                      // 1: (re)fill the devices before running (hf_xd_spawn)
};

__host__ __device__ inline size_t hfxd_smem_bytes(int M, int n_pad, int warps) {
    // CTA: g6, gd (f64) | k[36], ann[4] (f64) | lin, off (i32)
    // per warp: T (f64) | the chosen lane's K <= 31 rates (32 f64) | pos (u32) | counters (u32)
    // (birth[] stays in global memory: it is touched once per LOST exciton, by one lane)
    return 8 * (2 * (size_t)M + HFX_N_K + 4) + 4 * 2 * (size_t)((M + 1) & ~1) +
           (size_t)warps * (8 * ((size_t)n_pad + HFXD_SCRATCH) + 4 * ((size_t)n_pad + 64));
}

__device__ __forceinline__ int hfxd_occ_get(const uint32_t *occ, int64_t idx) {
    return (int)((occ[idx >> 4] >> (2 * (int)(idx & 15))) & 3u);
}
__device__ __forceinline__ void hfxd_occ_set(uint32_t *occ, int64_t idx, int code) {
    const int sh = 2 * (int)(idx & 15);
    occ[idx >> 4] = (occ[idx >> 4] & ~(3u << sh)) | ((uint32_t)code << sh);
}

struct HfxdWarp {                      // one warp's view of its device
    const HfxdParams *P;
    const double *s_g6, *s_gd, *s_k, *s_ann;
    const int32_t *s_lin, *s_off;
    double *s_T, *s_rates;
    uint32_t *s_pos, *s_cnt, *s_dirty;
    const double *ws1, *wt1;           // this device's realisation, site 0
    uint32_t *occ;                     // this device's mask (index = site + halo)
    int lane, K, M;
    bool aligned;                      // M == 32 * HFXD_KF and every mask index fits 31 bits: the static fast path
};

// Lane sums of the channel rates of an exciton (packed site + spin).
// `seen` gets bit k set when this lane's channel of slot k points at an occupied site.  HFXD_CHUNK slots are
// gathered (weight + occupancy word) before any of them is consumed.  ONE inlined instance per kernel (the event
// loop is a small state machine around a single call): with five copies the kernel was 7 100 instructions and a
// third of its issue stalls were instruction fetches.
//
// Fast path (KF = M / 32 full slots of transfer channels followed by the three on-site channels on lanes 0-2 of slot
// KF — cutoff 4 gives M = 256 = 8 x 32): no per-slot channel-range tests, 32-bit offsets, and the periodic-wrap arithmetic
// only for sites next to the x / y seams (a warp-uniform branch).  Every other cutoff takes the generic loop.
#define HFXD_KF 8
#ifndef HFXD_FAST_CHUNK
#define HFXD_FAST_CHUNK 2               // must divide HFXD_KF; measured: 2 -> 2.67e8, 4 -> 2.56e8, 8 (fully unrolled) -> 2.19e8 events/s on 4096 x 256
#endif
__device__ __forceinline__ double hfxd_lane_sum(const HfxdWarp &W, uint32_t ps, unsigned &seen) {
    const HfxParams &X = W.P->X;
    const int site = (int)(ps & 0x3fffffffu), spin = (int)(ps >> 30);
    const bool singlet = spin == HFX_SINGLET;
    const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
    const int plane = X.X * X.Y;
    const int64_t i = (int64_t)pz * plane + py * X.X + px;
    const double *w = (singlet ? W.ws1 : W.wt1) + i;
    const double wi = __ldg(w);
    const int a = hfx_tag_of(wi);
    const double inv_wi = __ddiv_rn(1.0, wi);
    const double *pref = W.s_k + (singlet ? 0 : 9) + 3 * a;
    const double *g = singlet ? W.s_g6 : W.s_gd;
    const double *ann = W.s_ann + (singlet ? 0 : 2);
    const bool nowrap = px >= X.R && px < X.X - X.R && py >= X.R && py < X.Y - X.R;
    const int64_t oi = i + X.w_halo;
    double S = 0.0;
    seen = 0;
    if (W.aligned) {
        // what the loop reads through, pinned in registers: nvcc otherwise rebuilds the pointers (a 64-bit multiply for
        // the mask, the shared-window base for the tables) at every use
        const uint32_t *occ = W.occ;
        const double *wq = w;
        unsigned oi32 = (unsigned)oi;                                            // the host checked that the mask index fits
        hfx_sptr lin = hfx_saddr(W.s_lin + W.lane), gl = hfx_saddr(g + W.lane), pr = hfx_saddr(pref), an = hfx_saddr(ann);
        const hfx_sptr lin_to_off = hfx_saddr(W.s_off) - hfx_saddr(W.s_lin);      // a site next to a seam takes its offsets from the packed table
        HFX_PIN64(occ); HFX_PIN64(wq); HFX_PIN32(oi32); HFX_PIN32(lin); HFX_PIN32(gl); HFX_PIN32(pr); HFX_PIN32(an);
        // a ROLLED loop over chunks of HFXD_FAST_CHUNK slots: the hot code stays small enough for the instruction
        // caches (seven warps per scheduler, each somewhere else in a 5 000-instruction kernel)
#pragma unroll 1
        for (int k0 = 0; k0 < HFXD_KF; k0 += HFXD_FAST_CHUNK, lin += 4 * 32 * HFXD_FAST_CHUNK, gl += 8 * 32 * HFXD_FAST_CHUNK) {
            double wj[HFXD_FAST_CHUNK];
            uint32_t ow[HFXD_FAST_CHUNK];
            int sh[HFXD_FAST_CHUNK];
#pragma unroll
            for (int u = 0; u < HFXD_FAST_CHUNK; ++u) {
                int d = hfx_lds_s32(lin + 4 * 32 * u);
                if (!nowrap) d = hfx_delta(hfx_lds_s32(lin + lin_to_off + 4 * 32 * u), px, py, X.X, X.Y, plane);   // warp-uniform branch
                wj[u] = __ldg(wq + d);
                const unsigned oj = oi32 + (unsigned)d;
                ow[u] = occ[oj >> 4];
                sh[u] = 2 * (int)(oj & 15u);
            }
            double rate[HFXD_FAST_CHUNK];
            unsigned any = 0;
#pragma unroll
            for (int u = 0; u < HFXD_FAST_CHUNK; ++u) {
                rate[u] = __dmul_rn(__dmul_rn(hfx_lds_f64(pr + 8 * hfx_tag_of(wj[u])), hfx_lds_f64(gl + 8 * 32 * u)),
                                    hfx_min1(__dmul_rn(wj[u], inv_wi)));         // hfx_rate()
                ow[u] = (ow[u] >> sh[u]) & 3u;
                any |= ow[u];
            }
            if (any) {                                                           // an occupied acceptor: rare, one branch per chunk
#pragma unroll
                for (int u = 0; u < HFXD_FAST_CHUNK; ++u)
                    if (ow[u]) { rate[u] = __dmul_rn(rate[u], hfx_lds_f64(an + (ow[u] == HFX_SINGLET ? 0 : 8))); seen |= 1u << (k0 + u); }
            }
#pragma unroll
            for (int u = 0; u < HFXD_FAST_CHUNK; ++u) S = __dadd_rn(S, rate[u]);
        }
        double rate = 0.0;                                                       // slot KF: ISC / RISC, radiative, non-radiative on lanes 0, 1, 2
        if (W.lane == 0) rate = singlet ? W.s_k[18 + a] : __dmul_rn(W.s_k[21 + a], hfx_min1(__dmul_rn(__ldg(W.ws1 + i), inv_wi)));
        else if (W.lane == 1) rate = W.s_k[24 + 2 * a + (singlet ? 0 : 1)];
        else if (W.lane == 2) rate = W.s_k[30 + 2 * a + (singlet ? 0 : 1)];
        return __dadd_rn(S, rate);
    }
#pragma unroll 1
    for (int k0 = 0; k0 < W.K; k0 += HFXD_CHUNK) {
        double wj[HFXD_CHUNK];
        uint32_t ow[HFXD_CHUNK];
        int sh[HFXD_CHUNK];
#pragma unroll
        for (int u = 0; u < HFXD_CHUNK; ++u) {
            const int c = ((k0 + u) << 5) + W.lane;
            wj[u] = 0.0; ow[u] = 0; sh[u] = 0;
            if (c < W.M) {
                const int delta = nowrap ? W.s_lin[c] : hfx_delta(W.s_off[c], px, py, X.X, X.Y, plane);
                wj[u] = __ldg(w + delta);
                const int64_t oj = oi + delta;
                ow[u] = W.occ[oj >> 4];
                sh[u] = 2 * (int)(oj & 15);
            }
        }
#pragma unroll
        for (int u = 0; u < HFXD_CHUNK; ++u) {
            const int k = k0 + u, c = (k << 5) + W.lane;
            if (k < W.K) {
                double rate = 0.0;
                if (c < W.M) {
                    rate = hfx_rate(pref, g[c], wj[u], inv_wi);
                    const int o = (int)((ow[u] >> sh[u]) & 3u);
                    if (o) { rate = __dmul_rn(rate, ann[o == HFX_SINGLET ? 0 : 1]); seen |= 1u << k; }
                } else if (c == W.M) {
                    rate = singlet ? W.s_k[18 + a] : __dmul_rn(W.s_k[21 + a], hfx_min1(__dmul_rn(__ldg(W.ws1 + i), inv_wi)));
                } else if (c == W.M + 1) {
                    rate = W.s_k[24 + 2 * a + (singlet ? 0 : 1)];
                } else if (c == W.M + 2) {
                    rate = W.s_k[30 + 2 * a + (singlet ? 0 : 1)];
                }
                S = __dadd_rn(S, rate);
            }
        }
    }
    return S;
}

// Slot of the exciton sitting on packed site `nb` (-1 if none); warp-uniform.
__device__ __forceinline__ int hfxd_find(const HfxdWarp &W, int nb) {
    for (int q0 = 0; q0 < W.P->n; q0 += 32) {
        const int q = q0 + W.lane;
        const unsigned hit = __ballot_sync(HF_FULL, q < W.P->n && (int)(W.s_pos[q] & 0x3fffffffu) == nb);
        if (hit) return q0 + __ffs(hit) - 1;
    }
    return -1;
}

// Mark as dirty the excitons on the occupied sites an evaluation at packed site `site` has met; dmask (warp-uniform)
// gets the bit of every word of s_dirty that was touched.
__device__ __forceinline__ void hfxd_mark_seen(const HfxdWarp &W, int site, unsigned seen, unsigned &dmask) {
    const HfxParams &X = W.P->X;
    unsigned any = __ballot_sync(HF_FULL, seen != 0);
    while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        unsigned m = __shfl_sync(HF_FULL, seen, src);
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const int o = W.s_off[(k << 5) + src];
            int xx = hf_px(site) + (o & 255) - 128, yy = hf_py(site) + ((o >> 8) & 255) - 128;
            if (xx < 0) xx += X.X; else if (xx >= X.X) xx -= X.X;
            if (yy < 0) yy += X.Y; else if (yy >= X.Y) yy -= X.Y;
            const int q = hfxd_find(W, hf_pack(xx, yy, hf_pz(site) + ((o >> 16) & 255) - 128));
            if (q >= 0) {
                if (W.lane == 0) W.s_dirty[q >> 5] |= 1u << (q & 31);
                dmask |= 1u << (q >> 5);
            }
        }
    }
    __syncwarp();
}

// The canonical pick (oracle: pick()) in two steps.  Lane: L = this lane's scanned sum, positive = its own sum > 0;
// returns the chosen lane and the running sum below it.  Slot: walks v[slot][lstar] in shared memory from there; returns
// the chosen index (slot * 32 + lane) and the running sum before it.  Both warp-uniform.
__device__ __forceinline__ int hfxd_pick_lane(double L, bool positive, double target, double &acc0) {
    const unsigned over = __ballot_sync(HF_FULL, L > target);
    const unsigned nonzero = __ballot_sync(HF_FULL, positive);
    const int lstar = over ? __ffs(over) - 1 : (nonzero ? 31 - __clz(nonzero) : 0);
    const double below = __shfl_sync(HF_FULL, L, (lstar + 31) & 31);
    acc0 = lstar > 0 ? below : 0.0;
    return lstar;
}
template <bool kCompact = false>      // kCompact: v[slot] holds the chosen lane's values only
__device__ __forceinline__ int hfxd_pick_slot(const double *v, int n_slots, int lstar, double acc, double target, double &base) {
    double b_slot = acc, b_last = acc;
    int slot = -1, last = -1;
    for (int q = 0; q < n_slots; ++q) {
        const double r = v[kCompact ? q : (q << 5) + lstar], before = acc;
        acc = __dadd_rn(acc, r);
        if (r > 0.0) { last = q; b_last = before; }
        if (slot < 0 && acc > target) { slot = q; b_slot = before; }
    }
    if (slot < 0) { slot = last < 0 ? n_slots - 1 : last; b_slot = b_last; }
    base = b_slot;
    return (slot << 5) + lstar;
}

// The K channels of lane `lstar` of an exciton (packed site + spin), evaluated by lanes 0 .. K - 1 with the arithmetic
// of hfxd_lane_sum and left in s_rates[k] (K <= 31, checked by the host).
__device__ __forceinline__ void hfxd_column(const HfxdWarp &W, uint32_t ps, int lstar) {
    const HfxParams &X = W.P->X;
    const int site = (int)(ps & 0x3fffffffu), spin = (int)(ps >> 30);
    const bool singlet = spin == HFX_SINGLET;
    const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
    const int plane = X.X * X.Y;
    const int64_t i = (int64_t)pz * plane + py * X.X + px;
    const double *w = (singlet ? W.ws1 : W.wt1) + i;
    const double wi = __ldg(w);
    const int a = hfx_tag_of(wi);
    const int c = (W.lane << 5) + lstar;
    if (W.lane < W.K) {
        const double inv_wi = __ddiv_rn(1.0, wi);
        double rate = 0.0;
        if (c < W.M) {
            const bool nowrap = px >= X.R && px < X.X - X.R && py >= X.R && py < X.Y - X.R;
            const int delta = nowrap ? W.s_lin[c] : hfx_delta(W.s_off[c], px, py, X.X, X.Y, plane);
            const double wj = __ldg(w + delta);
            const int o = hfxd_occ_get(W.occ, i + X.w_halo + delta);
            rate = hfx_rate(W.s_k + (singlet ? 0 : 9) + 3 * a, (singlet ? W.s_g6 : W.s_gd)[c], wj, inv_wi);
            if (o) rate = __dmul_rn(rate, W.s_ann[(singlet ? 0 : 2) + (o == HFX_SINGLET ? 0 : 1)]);
        } else if (c == W.M) {
            rate = singlet ? W.s_k[18 + a] : __dmul_rn(W.s_k[21 + a], hfx_min1(__dmul_rn(__ldg(W.ws1 + i), inv_wi)));
        } else if (c == W.M + 1) {
            rate = W.s_k[24 + 2 * a + (singlet ? 0 : 1)];
        } else if (c == W.M + 2) {
            rate = W.s_k[30 + 2 * a + (singlet ? 0 : 1)];
        }
        W.s_rates[W.lane] = rate;
    }
    __syncwarp();
}

// A new exciton for slot k: the first empty interior site along the device's stream (oracle: replace()).
__device__ __forceinline__ uint32_t hfxd_replace(const HfxdWarp &W, uint32_t ident, uint64_t &draws) {
    const HfxParams &X = W.P->X;
    const int64_t n_interior = (int64_t)X.X * X.Y * (X.Z - 2);
    for (;;) {
        const HfPhiloxBlock blk = hf_philox_block(X.k0, X.k1, HF_STREAM_XDENSE, ident, draws >> 1);
        draws += 2;
        const double ua = hf_block_uniform(blk, 0), ub = hf_block_uniform(blk, 1);
        int64_t q = (int64_t)floor(ua * (double)n_interior);
        if (q >= n_interior) q = n_interior - 1;
        const int x = (int)(q % X.X), y = (int)((q / X.X) % X.Y), z = 1 + (int)(q / ((int64_t)X.X * X.Y));
        const int64_t oi = ((int64_t)z * X.Y + y) * X.X + x + X.w_halo;
        if (hfxd_occ_get(W.occ, oi)) continue;                                   // same word for every lane: uniform
        const int spin = ub < X.singlet_fraction ? HFX_SINGLET : HFX_TRIPLET;
        __syncwarp();
        if (W.lane == 0) hfxd_occ_set(W.occ, oi, spin);
        __syncwarp();
        return (uint32_t)hf_pack(x, y, z) | ((uint32_t)spin << 30);
    }
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, HFXD_MIN_BLOCKS) hfxd_run_kernel(const HfxdParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const HfxParams &X = P.X;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = X.M, padded = (int)hfx_padded_channels(M), K = padded >> 5, n_pad = P.n_pad, n_slots = n_pad >> 5;
    double *s_g6 = reinterpret_cast<double *>(hf_smem);
    double *s_gd = s_g6 + M;
    double *s_k = s_gd + M;
    double *s_ann = s_k + HFX_N_K;
    int32_t *s_lin = reinterpret_cast<int32_t *>(s_ann + 4);
    int32_t *s_off = s_lin + ((M + 1) & ~1);
    unsigned char *wsm = reinterpret_cast<unsigned char *>(s_off + ((M + 1) & ~1)) +
                         (size_t)warp * (8 * ((size_t)n_pad + HFXD_SCRATCH) + 4 * ((size_t)n_pad + 64));
    HfxdWarp W;
    W.P = &P; W.s_g6 = s_g6; W.s_gd = s_gd; W.s_k = s_k; W.s_ann = s_ann; W.s_lin = s_lin; W.s_off = s_off;
    W.s_T = reinterpret_cast<double *>(wsm); W.s_rates = W.s_T + n_pad;
    W.s_pos = reinterpret_cast<uint32_t *>(W.s_rates + HFXD_SCRATCH); W.s_cnt = W.s_pos + n_pad; W.s_dirty = W.s_cnt + 32;
    W.lane = lane; W.K = K; W.M = M;
    W.aligned = M == 32 * HFXD_KF && X.w_stride < 0x7fffffff;
    for (int m = threadIdx.x; m < M; m += kWarps * 32) { s_g6[m] = X.g6[m]; s_gd[m] = X.gd[m]; s_lin[m] = X.lin[m]; s_off[m] = X.offsets[m]; }
    if (threadIdx.x < 9) { s_k[threadIdx.x] = X.kf[threadIdx.x]; s_k[9 + threadIdx.x] = X.kd[threadIdx.x]; }
    if (threadIdx.x < 3) { s_k[18 + threadIdx.x] = X.k_isc[threadIdx.x]; s_k[21 + threadIdx.x] = X.k_risc[threadIdx.x]; }
    if (threadIdx.x < 6) { s_k[24 + threadIdx.x] = X.k_rad[threadIdx.x]; s_k[30 + threadIdx.x] = X.k_nonrad[threadIdx.x]; }
    if (threadIdx.x < 4) s_ann[threadIdx.x] = P.ann[threadIdx.x];
    __syncthreads();

    for (int64_t dev = (int64_t)blockIdx.x * kWarps + warp; dev < X.B; dev += (int64_t)gridDim.x * kWarps) {
        const uint32_t ident = X.first_traj + (uint32_t)dev;
        const int64_t r = dev / X.traj_per_real;
        W.ws1 = X.ws1 + r * X.w_stride + X.w_halo; W.wt1 = X.wt1 + r * X.w_stride + X.w_halo;
        W.occ = P.occ + dev * P.occ_words;
        uint32_t *g_pos = P.pos + dev * n_pad;
        double *g_T = P.T + dev * n_pad, *g_birth = P.birth + dev * n_pad;
        double *g_L = P.rec_L + dev * n_pad * 32 + lane;                         // this lane's column of the device's records
        uint32_t *g_seen = P.rec_seen + dev * n_pad * 32 + lane;
#if HFXD_PIN_RECORDS
        HFX_PIN64(g_L); HFX_PIN64(g_seen);                                       // else rebuilt (two 64-bit multiplies) at every use
#endif
        double time, life;
        uint64_t draws;
        W.s_cnt[lane] = 0; W.s_dirty[lane] = 0;
        __syncwarp();
        unsigned seen, dmask = 0;                                                // dmask: words of s_dirty that may be non-zero
        if (P.fill) {
            // hf_xd_spawn: an empty mask (memset by the host), slots filled in order; their totals are computed by the
            // refresh pass below (ev = -1: every slot marked dirty)
            time = 0.0; life = 0.0; draws = 0;
            for (int q = lane; q < n_pad; q += 32) { W.s_pos[q] = 0; W.s_T[q] = 0.0; g_birth[q] = 0.0; }
            __syncwarp();
            for (int k = 0; k < P.n; ++k) {
                const uint32_t ps = hfxd_replace(W, ident, draws);
                if (lane == 0) { W.s_pos[k] = ps; W.s_cnt[HFX_T_GENERATED] += 1; }
            }
            __syncwarp();
            if (lane < n_slots) W.s_dirty[lane] = (lane << 5) + 32 <= P.n ? 0xffffffffu : ((lane << 5) < P.n ? (1u << (P.n & 31)) - 1u : 0u);
            dmask = n_slots >= 32 ? 0xffffffffu : (1u << n_slots) - 1u;
        } else {
            time = P.time[dev]; life = P.life[dev]; draws = P.draws[dev];
            for (int q = lane; q < n_pad; q += 32) { W.s_pos[q] = g_pos[q]; W.s_T[q] = g_T[q]; }
        }
        __syncwarp();
        const bool traced = X.trace && dev >= X.trace_first && dev < X.trace_first + X.trace_n;
        uint64_t trace_at = traced ? X.trace_len[dev - X.trace_first] : 0;

        for (int64_t ev = P.fill ? -1 : 0; ev < X.n_events; ++ev) {
            // refresh stage 1: a slot whose own site / spin changed (its evaluation names the excitons it meets)
            //               2: the slots marked dirty
            int stage = 2, tgt = -1, promoted = -1;
            uint32_t job = 0;
            if (ev >= 0) {
                // (1) the exciton: canonical scan over the cached totals
                double A = 0.0;
                for (int q = 0; q < n_slots; ++q) A = __dadd_rn(A, W.s_T[(q << 5) + lane]);
                const double G = hfx_scan(A, lane);
                const double R = __shfl_sync(HF_FULL, G, 31);
                const HfPhiloxBlock rnd = hf_philox_block(X.k0, X.k1, HF_STREAM_XDENSE, ident, draws >> 1);
                draws += 2;
                const double u_sel = hf_block_uniform(rnd, 0), u_time = hf_block_uniform(rnd, 1);
                const double target = __dmul_rn(u_sel, R);
                double acc0, base;
                const int xl = hfxd_pick_lane(G, A > 0.0, target, acc0);
                const int kx = hfxd_pick_slot(W.s_T, n_slots, xl, acc0, target, base);
                const uint32_t ps = W.s_pos[kx];
                // (2) its channel, from the record of its last evaluation: the lane by the scanned sums, then that lane's
                //     K channels again (same inputs, same arithmetic)
                const double Lk = g_L[kx * 32];
                const uint32_t sk = g_seen[kx * 32];
                const double target2 = __dsub_rn(target, base);
                const int lstar = hfxd_pick_lane(Lk, (sk >> 31) != 0u, target2, acc0);
                hfxd_column(W, ps, lstar);
                double base2;
                const int chosen = hfxd_pick_slot<true>(W.s_rates, K, lstar, acc0, target2, base2);
                hfxd_mark_seen(W, (int)(ps & 0x3fffffffu), sk & 0x7fffffffu, dmask);     // the neighbours of the site that changes
                const double dt = __ddiv_rn(-log(__dsub_rn(1.0, u_time)), R);
                time = __dadd_rn(time, dt);
                // (3) apply
                const int site = (int)(ps & 0x3fffffffu), spin = (int)(ps >> 30);
                const bool singlet = spin == HFX_SINGLET;
                const int sidx = singlet ? 0 : 1;
                const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
                const int64_t i = hf_site(X.X, X.Y, site);
                const int a = hfx_tag_of(__ldg((singlet ? W.ws1 : W.wt1) + i));
                int kind, tally, fin = site;
                bool lost = false;
                uint32_t ps_after = ps;
                if (chosen < M) {
                    const int o = W.s_off[chosen];
                    int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                    if (xx < 0) xx += X.X; else if (xx >= X.X) xx -= X.X;
                    if (yy < 0) yy += X.Y; else if (yy >= X.Y) yy -= X.Y;
                    fin = hf_pack(xx, yy, pz + ((o >> 16) & 255) - 128);
                    const int64_t oj = hf_site(X.X, X.Y, fin) + X.w_halo;
                    const int occ_j = hfxd_occ_get(W.occ, oj);
                    __syncwarp();
                    if (occ_j == 0) {
                        if (lane == 0) { hfxd_occ_set(W.occ, i + X.w_halo, 0); hfxd_occ_set(W.occ, oj, spin); }
                        ps_after = (uint32_t)fin | ((uint32_t)spin << 30);
                        kind = singlet ? HFX_K_FORSTER : HFX_K_MOVE;
                        tally = singlet ? HFX_T_FORSTER : HFX_T_DEXTER;
                    } else {
                        const int oidx = occ_j == HFX_SINGLET ? 0 : 1;
                        kind = HFXD_K_ANNIHILATION;
                        tally = HFXD_T_ANN + 2 * sidx + oidx;
                        if (sidx == 1 && oidx == 1) {                           // triplet on triplet: the survivor may be promoted
                            const HfPhiloxBlock b2 = hf_philox_block(X.k0, X.k1, HF_STREAM_XDENSE, ident, draws >> 1);
                            draws += 2;
                            if (hf_block_uniform(b2, 0) < P.p_tta) {
                                promoted = hfxd_find(W, fin);
                                if (lane == 0) {
                                    if (promoted >= 0) W.s_pos[promoted] = (uint32_t)fin | ((uint32_t)HFX_SINGLET << 30);
                                    hfxd_occ_set(W.occ, oj, HFX_SINGLET);
                                    W.s_cnt[HFXD_T_TTA_UP] += 1;
                                }
                            }
                        }
                        lost = true;
                    }
                } else if (chosen == M) {
                    kind = HFX_K_ISC;                                           // molecule.py:299-307
                    tally = singlet ? HFX_T_ISC : HFX_T_RISC;
                    const int flipped = singlet ? HFX_TRIPLET : HFX_SINGLET;
                    ps_after = (uint32_t)site | ((uint32_t)flipped << 30);
                    if (lane == 0) hfxd_occ_set(W.occ, i + X.w_halo, flipped);
                } else {
                    kind = HFX_K_DECAY;
                    tally = (chosen == M + 1 ? HFX_T_RAD : (singlet ? HFX_T_NONRAD_S : HFX_T_NONRAD_T)) + a;
                    lost = true;
                }
                if (lost) {
                    if (lane == 0) hfxd_occ_set(W.occ, i + X.w_halo, 0);
                    life = __dadd_rn(life, __dsub_rn(time, g_birth[kx]));        // every lane reads the same word
                    __syncwarp();
                    ps_after = hfxd_replace(W, ident, draws);
                    if (lane == 0) { g_birth[kx] = time; W.s_cnt[HFX_T_GENERATED] += 1; }
                }
                if (lane == 0) {
                    W.s_pos[kx] = ps_after;
                    W.s_cnt[HFX_T_EVENTS] += 1;
                    W.s_cnt[tally] += 1;
                    if (traced && trace_at < (uint64_t)X.trace_cap) {
                        HfTraceRec rec;
                        rec.initial = (int32_t)i; rec.final_ = (int32_t)hf_site(X.X, X.Y, fin);
                        rec.tau = dt; rec.kind = (uint8_t)kind; rec.particule = HF_PART_EXCITON;
                        rec.pad[0] = (uint8_t)(ps_after >> 30); rec.pad[1] = 0;
                        rec.spins = (uint32_t)chosen | ((uint32_t)kx << 16); rec.draws = draws;
                        X.trace[(dev - X.trace_first) * X.trace_cap + trace_at] = rec;
                    }
                }
                trace_at += 1;
                __syncwarp();
                // (4) refresh the cached totals and records: the slot itself in its new state, a promoted acceptor (whose
                //     neighbours see a different spin), then everybody marked
                stage = 1; tgt = kx; job = ps_after;
            }
            for (;;) {
                if (stage == 2) {                                                // next slot marked dirty, ascending
                    if (!dmask) break;
                    const int wq = __ffs(dmask) - 1;
                    const unsigned todo = W.s_dirty[wq];
                    if (!(todo & (todo - 1u))) dmask &= dmask - 1u;              // the word's last bit (or none): done with it after this
                    if (!todo) continue;
                    tgt = (wq << 5) + __ffs(todo) - 1; job = W.s_pos[tgt];
                }
                const double S = hfxd_lane_sum(W, job, seen);
                const double L = hfx_scan(S, lane);
                __syncwarp();
                if (stage == 1) hfxd_mark_seen(W, (int)(job & 0x3fffffffu), seen, dmask);   // the neighbours of the site that changed
                g_L[tgt * 32] = L;
                g_seen[tgt * 32] = seen | (S > 0.0 ? 0x80000000u : 0u);
                const double tot = __shfl_sync(HF_FULL, L, 31);
                if (lane == 0) { W.s_T[tgt] = tot; W.s_dirty[tgt >> 5] &= ~(1u << (tgt & 31)); }
                __syncwarp();
                if (stage == 1 && promoted >= 0 && tgt != promoted) { tgt = promoted; job = W.s_pos[promoted]; continue; }
                stage = 2;
            }
            __syncwarp();
        }
        for (int q = lane; q < n_pad; q += 32) { g_pos[q] = W.s_pos[q]; g_T[q] = W.s_T[q]; }
        if (lane == 0) {
            P.time[dev] = time; P.life[dev] = life; P.draws[dev] = draws;
            if (traced) X.trace_len[dev - X.trace_first] = trace_at;
        }
        __syncwarp();
        if (lane < HFXD_T_N && W.s_cnt[lane]) atomicAdd(P.totals + r * HFXD_T_N + lane, (unsigned long long)W.s_cnt[lane]);
        __syncwarp();
    }
}
