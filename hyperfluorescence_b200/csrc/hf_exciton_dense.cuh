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
//   birth[n_pad] f64  device clock when the exciton was created (global memory only: touched once per lost exciton)
// The OCCUPANCY CODE of a site (0 empty / 1 singlet / 3 triplet) is not stored per site: a device holds n <= 1024
// excitons on millions of sites, so the mask is a hash set of the device's own excitons in shared memory
// (open addressing, linear probing, load factor <= 1/2: keys = the pos[] words, values = their slots), rebuilt
// from pos[] at the start of a launch.  Round 1 kept 2 bits per site in global memory (0.55 MB per device, 2.2 GB
// for BASELINE configs[3]): every rate evaluation gathered 257 words of it, 37 % of them DRAM misses (ncu r01l),
// and the kernel waited on those.  A lookup is now a shared-memory probe and names the neighbour's SLOT as well,
// so the excitons that can see a changed site are marked for re-evaluation directly.
// Per event the warp (1) scans the cached totals to pick an exciton, (2) evaluates that exciton's
// M + 3 channels (lane = channel % 32; each lane also looks up the acceptor's occupancy) and scans
// them to pick the channel, (3) applies it — lane 0 edits the hash set — and (4) re-evaluates the
// totals of the excitons that can see a changed site.
#pragma once
#include "hf_exciton.cuh"

#define HF_STREAM_XDENSE 5u
#define HFXD_K_ANNIHILATION 6          // the `unbound` slot of event.py:49
#define HFXD_T_ANN 16                  // + 2 * donor spin index + acceptor spin index
#define HFXD_T_TTA_UP 20
#define HFXD_T_N 24
#define HFXD_MAX_EXCITONS 1024
#ifndef HFXD_MIN_BLOCKS
#define HFXD_MIN_BLOCKS 2              // 2 CTAs x 14 warps = 28 devices per SM: the 4096 devices of BASELINE configs[3] are ONE wave on 148 SMs
#endif
#define HFXD_KREG 9                    // slots of 32 channels whose rates stay in registers between evaluation and pick (cutoff <= 4)
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
    int32_t hash_size;                 // power of two >= 2 * n_pad
    double *time, *life;               // [B] device clock, summed lifetimes of the lost excitons
    uint64_t *draws;                   // [B]
    unsigned long long *totals;        // [n_real][HFXD_T_N]
    int32_t fill;                      // 1: (re)fill the devices before running (hf_xd_spawn)
};

__host__ __device__ inline int hfxd_hash_size(int n_pad) { int h = 64; while (h < 2 * n_pad) h <<= 1; return h; }
__host__ __device__ inline bool hfxd_rates_in_registers(int M) { return (M + 3 + 31) / 32 <= HFXD_KREG; }
__host__ __device__ inline size_t hfxd_warp_bytes(int M, int n_pad) {
    // T (f64) | rates (f64, only when they do not fit the registers) | pos, keys (u32) | counters, dirty (u32) | bitmap (u32) | vals (u16)
    const size_t h = (size_t)hfxd_hash_size(n_pad);
    size_t b = 8 * ((size_t)n_pad + (hfxd_rates_in_registers(M) ? 0 : hfx_padded_channels(M))) + 4 * ((size_t)n_pad + h + 64 + h / 2) + 2 * h;
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t hfxd_smem_bytes(int M, int n_pad, int warps) {
    // CTA: g6, gd (f64) | k[36], ann[4] (f64) | off, lin, pk, pkc, pkc2 (i32)   + per warp
    return 8 * (2 * (size_t)M + HFX_N_K + 4) + 4 * 5 * (size_t)((M + 3) & ~3) + (size_t)warps * hfxd_warp_bytes(M, n_pad);
}

struct HfxdWarp {                      // one warp's view of its device
    const HfxdParams *P;
    const double *s_g6, *s_gd, *s_k, *s_ann;
    const int32_t *s_off, *s_lin, *s_pk, *s_pkc, *s_pkc2;   // packed offsets; linear, packed and hashed-packed site deltas (valid away from the x / y seams)
    double *s_T, *s_rates;
    uint32_t *s_pos, *s_cnt, *s_dirty, *s_keys, *s_bits;
    uint16_t *s_vals;
    const double *ws1, *wt1;           // this device's realisation, site 0
    uint32_t hmask;
    int hshift;
    int lane, K, M;
};

// ---- the occupancy hash set -------------------------------------------------------------------------------------
// keys[H] / vals[H]: open addressing, linear probing, bucket = top bits of site * C1.  bits[16 H]: a Bloom filter with
// two hash functions (top bits of site * C1 and of site * C2), set on insert and never cleared before the next rebuild
// (stale bits only cost a wasted exact lookup): a clear bit proves the site empty.  With n = H / 2 entries a lane's
// lookup passes the filter with probability (1 - exp(-1/16))^2 = 0.4 %, a warp's with 11 %: the divergent exact probe
// is the exception, not the rule.
#define HFXD_C1 0x9E3779B1u
#define HFXD_C2 0x85EBCA6Bu
__device__ __forceinline__ uint32_t hfxd_hslot(const HfxdWarp &W, uint32_t site) { return (site * HFXD_C1) >> W.hshift; }
__device__ __forceinline__ bool hfxd_bloom(const HfxdWarp &W, uint32_t p1, uint32_t p2) {    // p1 = site * C1, p2 = site * C2
    const uint32_t f1 = p1 >> (W.hshift - 4), f2 = p2 >> (W.hshift - 4);
    return ((W.s_bits[f1 >> 5] >> (f1 & 31)) & (W.s_bits[f2 >> 5] >> (f2 & 31)) & 1u) != 0;
}
__device__ __forceinline__ bool hfxd_maybe(const HfxdWarp &W, uint32_t site) { return hfxd_bloom(W, site * HFXD_C1, site * HFXD_C2); }
__device__ __forceinline__ void hfxd_bit_set(const HfxdWarp &W, uint32_t site) {
    const uint32_t f1 = (site * HFXD_C1) >> (W.hshift - 4), f2 = (site * HFXD_C2) >> (W.hshift - 4);
    W.s_bits[f1 >> 5] |= 1u << (f1 & 31);
    W.s_bits[f2 >> 5] |= 1u << (f2 & 31);
}
// entry of the exciton on packed `site`, -1 if the site is empty (any lane, any site)
__device__ __forceinline__ int hfxd_hfind(const HfxdWarp &W, uint32_t site) {
    uint32_t h = hfxd_hslot(W, site);
    for (;;) {
        const uint32_t e = W.s_keys[h];
        if (e == 0) return -1;                                                   // packed site 0 is a face site: never occupied
        if ((e & 0x3fffffffu) == site) return (int)h;
        h = (h + 1) & W.hmask;
    }
}
__device__ __forceinline__ int hfxd_occ_code(const HfxdWarp &W, uint32_t site) {
    const int h = hfxd_hfind(W, site);
    return h < 0 ? 0 : (int)(W.s_keys[h] >> 30);
}
// one lane: insert / erase (backward-shift deletion keeps every probe sequence intact) / change the spin
__device__ __forceinline__ void hfxd_hinsert(const HfxdWarp &W, uint32_t ps, int slot) {
    uint32_t h = hfxd_hslot(W, ps & 0x3fffffffu);
    while (W.s_keys[h]) h = (h + 1) & W.hmask;
    W.s_keys[h] = ps; W.s_vals[h] = (uint16_t)slot;
    hfxd_bit_set(W, ps & 0x3fffffffu);
}
// all lanes: the bitmap from the slots (start of a launch, and every so often to shed the stale bits of erased entries)
__device__ __forceinline__ void hfxd_bits_rebuild(const HfxdWarp &W, int n, int H) {
    __syncwarp();
    for (int q = W.lane; q < H / 2; q += 32) W.s_bits[q] = 0;
    __syncwarp();
    for (int q = W.lane; q < n; q += 32) {
        const uint32_t site = W.s_pos[q] & 0x3fffffffu;
        if (site) {
            const uint32_t f1 = (site * HFXD_C1) >> (W.hshift - 4), f2 = (site * HFXD_C2) >> (W.hshift - 4);
            atomicOr(&W.s_bits[f1 >> 5], 1u << (f1 & 31));
            atomicOr(&W.s_bits[f2 >> 5], 1u << (f2 & 31));
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void hfxd_herase(const HfxdWarp &W, uint32_t site) {
    int i = hfxd_hfind(W, site);
    if (i < 0) return;
    uint32_t j = (uint32_t)i;
    for (;;) {
        j = (j + 1) & W.hmask;
        const uint32_t e = W.s_keys[j];
        if (!e) break;
        const uint32_t k = hfxd_hslot(W, e & 0x3fffffffu);
        const bool stays = (uint32_t)i <= j ? ((uint32_t)i < k && k <= j) : ((uint32_t)i < k || k <= j);
        if (stays) continue;
        W.s_keys[i] = e; W.s_vals[i] = W.s_vals[j];
        i = (int)j;
    }
    W.s_keys[i] = 0;
}
__device__ __forceinline__ void hfxd_hset_spin(const HfxdWarp &W, uint32_t site, int spin) {
    const int h = hfxd_hfind(W, site);
    if (h >= 0) W.s_keys[h] = site | ((uint32_t)spin << 30);
}

// One transfer channel the slow way: periodic wrap in x and y, exact occupancy lookup, marking.  Returns the rate.
__device__ __forceinline__ double hfxd_channel(const HfxdWarp &W, const double *w, const double *pref, const double *g, const double *ann,
                                               double inv_wi, int px, int py, int pz, int c, bool mark) {
    const HfxParams &X = W.P->X;
    const int o = W.s_off[c];
    const int zz = pz + ((o >> 16) & 255) - 128;
    int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
    if (xx < 0) xx += X.X; else if (xx >= X.X) xx -= X.X;
    if (yy < 0) yy += X.Y; else if (yy >= X.Y) yy -= X.Y;
    double rate = hfx_rate(pref, g[c], __ldg(w + ((zz - pz) * X.X * X.Y + (yy - py) * X.X + (xx - px))), inv_wi);
    if (zz >= 0 && zz < X.Z) {                                                   // outside: the halo's weight 0 gave rate 0
        const uint32_t nb = (uint32_t)hf_pack(xx, yy, zz);
        if (hfxd_maybe(W, nb)) {
            const int h = hfxd_hfind(W, nb);
            if (h >= 0) {
                rate = __dmul_rn(rate, ann[(W.s_keys[h] >> 30) == HFX_SINGLET ? 0 : 1]);
                if (mark) { const int q = W.s_vals[h]; atomicOr(&W.s_dirty[q >> 5], 1u << (q & 31)); }
            }
        }
    }
    return rate;
}

// Lane sums of the channel rates of an exciton (packed site + spin): S = the lane's rates added in slot order.
// keep: leave the rates for the pick — in rr[] (registers, K <= HFXD_KREG) or in s_rates.  mark: the excitons met on
// occupied acceptor sites are marked dirty (they see the site that is about to change / has just changed).
// ONE inlined instance per kernel (keep / mark are run-time flags).
// Fast path (site away from the x / y seams, K <= HFXD_KREG — the usual case): linear offsets and hashed packed
// offsets from shared-memory tables, HFXD_CHUNK gathers in flight, occupancy through the bitmap only; the rare lanes
// whose bitmap test passes resolve it exactly afterwards and re-add their nine rates in slot order.
__device__ __forceinline__ double hfxd_lane_sum(const HfxdWarp &W, uint32_t ps, double (&rr)[HFXD_KREG], const bool keep, const bool mark) {
    const HfxParams &X = W.P->X;
    const int site = (int)(ps & 0x3fffffffu), spin = (int)(ps >> 30);
    const bool singlet = spin == HFX_SINGLET;
    const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
    const int64_t i = ((int64_t)pz * X.Y + py) * X.X + px;
    const double *w = (singlet ? W.ws1 : W.wt1) + i;
    const double wi = __ldg(w);
    const int a = hfx_tag_of(wi);
    const double inv_wi = __ddiv_rn(1.0, wi);
    const double *pref = W.s_k + (singlet ? 0 : 9) + 3 * a;
    const double *g = singlet ? W.s_g6 : W.s_gd;
    const double *ann = W.s_ann + (singlet ? 0 : 2);
    const double on0 = singlet ? W.s_k[18 + a] : __dmul_rn(W.s_k[21 + a], hfx_min1(__dmul_rn(__ldg(W.ws1 + i), inv_wi)));
    const double on1 = W.s_k[24 + 2 * a + (singlet ? 0 : 1)], on2 = W.s_k[30 + 2 * a + (singlet ? 0 : 1)];
    const bool fast = hfxd_rates_in_registers(W.M) && px >= X.R && px < X.X - X.R && py >= X.R && py < X.Y - X.R;
    double S = 0.0;
    if (fast) {
        const uint32_t site_c1 = (uint32_t)site * HFXD_C1, site_c2 = (uint32_t)site * HFXD_C2;
        unsigned hits = 0;
#pragma unroll
        for (int k0 = 0; k0 < HFXD_KREG; k0 += HFXD_CHUNK) {
            double wj[HFXD_CHUNK];
#pragma unroll
            for (int u = 0; u < HFXD_CHUNK; ++u) {
                const int c = ((k0 + u) << 5) + W.lane;
                wj[u] = c < W.M ? __ldg(w + W.s_lin[c]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < HFXD_CHUNK; ++u) {
                const int k = k0 + u, c = (k << 5) + W.lane;
                double rate;
                if (c < W.M) {
                    rate = hfx_rate(pref, g[c], wj[u], inv_wi);
                    // (site + pk) * C == site * C + pk * C (mod 2^32): the neighbour's two hashes without a multiply
                    hits |= (hfxd_bloom(W, site_c1 + (uint32_t)W.s_pkc[c], site_c2 + (uint32_t)W.s_pkc2[c]) ? 1u : 0u) << k;
                } else {
                    rate = c == W.M ? on0 : (c == W.M + 1 ? on1 : (c == W.M + 2 ? on2 : 0.0));
                }
                rr[k] = rate;
                S = __dadd_rn(S, rate);
            }
        }
        if (__any_sync(HF_FULL, hits != 0)) {                                    // somebody may have a neighbour: exact lookups
            if (hits) {
                bool changed = false;
#pragma unroll
                for (int k = 0; k < HFXD_KREG; ++k) {
                    if (hits >> k & 1u) {
                        const int c = (k << 5) + W.lane;
                        const int zz = pz + ((W.s_off[c] >> 16) & 255) - 128;
                        const int h = (zz >= 0 && zz < X.Z) ? hfxd_hfind(W, (uint32_t)(site + W.s_pk[c])) : -1;
                        if (h >= 0) {
                            rr[k] = __dmul_rn(rr[k], ann[(W.s_keys[h] >> 30) == HFX_SINGLET ? 0 : 1]);
                            changed = true;
                            if (mark) { const int q = W.s_vals[h]; atomicOr(&W.s_dirty[q >> 5], 1u << (q & 31)); }
                        }
                    }
                }
                if (changed) {
                    S = 0.0;
#pragma unroll
                    for (int k = 0; k < HFXD_KREG; ++k) S = __dadd_rn(S, rr[k]);  // slots >= K hold +0.0: S + 0.0 == S
                }
            }
        }
    } else {
        const bool in_regs = hfxd_rates_in_registers(W.M);
#pragma unroll 1
        for (int k = 0; k < W.K; ++k) {
            const int c = (k << 5) + W.lane;
            double rate;
            if (c < W.M) rate = hfxd_channel(W, w, pref, g, ann, inv_wi, px, py, pz, c, mark);
            else rate = c == W.M ? on0 : (c == W.M + 1 ? on1 : (c == W.M + 2 ? on2 : 0.0));
            if (keep) {
                if (in_regs) {
#pragma unroll
                    for (int q = 0; q < HFXD_KREG; ++q) if (q == k) rr[q] = rate;
                } else {
                    W.s_rates[c] = rate;
                }
            }
            S = __dadd_rn(S, rate);
        }
    }
    __syncwarp();
    return S;
}

// The canonical pick (oracle: pick()): v[n_slots][32] in shared memory, S = this lane's sum, L = its scan.
// Returns the chosen index (slot * 32 + lane) and the running sum before it; warp-uniform.
__device__ __forceinline__ int hfxd_pick(const double *v, int n_slots, double S, double L, double target, int lane, double &base) {
    const unsigned over = __ballot_sync(HF_FULL, L > target);
    const unsigned nonzero = __ballot_sync(HF_FULL, S > 0.0);
    const int lstar = over ? __ffs(over) - 1 : (nonzero ? 31 - __clz(nonzero) : 0);
    const double below = __shfl_sync(HF_FULL, L, (lstar + 31) & 31);
    double acc = lstar > 0 ? below : 0.0, b_slot = acc, b_last = acc;
    int slot = -1, last = -1;
    for (int q = 0; q < n_slots; ++q) {
        const double r = v[(q << 5) + lstar], before = acc;
        acc = __dadd_rn(acc, r);
        if (r > 0.0) { last = q; b_last = before; }
        if (slot < 0 && acc > target) { slot = q; b_slot = before; }
    }
    if (slot < 0) { slot = last < 0 ? n_slots - 1 : last; b_slot = b_last; }
    base = b_slot;
    return (slot << 5) + lstar;
}

// The same pick over the K <= HFXD_KREG rates every lane holds in registers (rr[q] = its channel of slot q).
__device__ __forceinline__ int hfxd_pick_regs(const double (&rr)[HFXD_KREG], int n_slots, double S, double L, double target, int lane) {
    const unsigned over = __ballot_sync(HF_FULL, L > target);
    const unsigned nonzero = __ballot_sync(HF_FULL, S > 0.0);
    const int lstar = over ? __ffs(over) - 1 : (nonzero ? 31 - __clz(nonzero) : 0);
    const double below = __shfl_sync(HF_FULL, L, (lstar + 31) & 31);
    double acc = lstar > 0 ? below : 0.0;
    int slot = -1, last = -1;
#pragma unroll
    for (int q = 0; q < HFXD_KREG; ++q) {
        const double r = __shfl_sync(HF_FULL, rr[q], lstar);
        if (q < n_slots) {
            acc = __dadd_rn(acc, r);
            if (r > 0.0) last = q;
            if (slot < 0 && acc > target) slot = q;
        }
    }
    if (slot < 0) slot = last < 0 ? n_slots - 1 : last;
    return (slot << 5) + lstar;
}

// A new exciton for slot k: the first empty interior site along the device's stream (oracle: replace()).
__device__ __forceinline__ uint32_t hfxd_replace(const HfxdWarp &W, uint32_t ident, uint64_t &draws, int slot) {
    const HfxParams &X = W.P->X;
    const int64_t n_interior = (int64_t)X.X * X.Y * (X.Z - 2);
    for (;;) {
        const HfPhiloxBlock blk = hf_philox_block(X.k0, X.k1, HF_STREAM_XDENSE, ident, draws >> 1);
        draws += 2;
        const double ua = hf_block_uniform(blk, 0), ub = hf_block_uniform(blk, 1);
        int64_t q = (int64_t)floor(ua * (double)n_interior);
        if (q >= n_interior) q = n_interior - 1;
        const int x = (int)(q % X.X), y = (int)((q / X.X) % X.Y), z = 1 + (int)(q / ((int64_t)X.X * X.Y));
        const uint32_t site = (uint32_t)hf_pack(x, y, z);
        if (hfxd_hfind(W, site) >= 0) continue;                                  // same probe for every lane: uniform
        const int spin = ub < X.singlet_fraction ? HFX_SINGLET : HFX_TRIPLET;
        const uint32_t ps = site | ((uint32_t)spin << 30);
        __syncwarp();
        if (W.lane == 0) hfxd_hinsert(W, ps, slot);
        __syncwarp();
        return ps;
    }
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, HFXD_MIN_BLOCKS) hfxd_run_kernel(const HfxdParams P) {
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const HfxParams &X = P.X;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = X.M, padded = (int)hfx_padded_channels(M), K = padded >> 5, n_pad = P.n_pad, n_slots = n_pad >> 5;
    const bool in_regs = hfxd_rates_in_registers(M);
    const int H = P.hash_size;
    double *s_g6 = reinterpret_cast<double *>(hf_smem);
    double *s_gd = s_g6 + M;
    double *s_k = s_gd + M;
    double *s_ann = s_k + HFX_N_K;
    int32_t *s_off = reinterpret_cast<int32_t *>(s_ann + 4);
    int32_t *s_lin = s_off + ((M + 3) & ~3), *s_pk = s_lin + ((M + 3) & ~3), *s_pkc = s_pk + ((M + 3) & ~3), *s_pkc2 = s_pkc + ((M + 3) & ~3);
    unsigned char *wsm = reinterpret_cast<unsigned char *>(s_pkc2 + ((M + 3) & ~3)) + (size_t)warp * hfxd_warp_bytes(M, n_pad);
    HfxdWarp W;
    W.P = &P; W.s_g6 = s_g6; W.s_gd = s_gd; W.s_k = s_k; W.s_ann = s_ann; W.s_off = s_off; W.s_lin = s_lin; W.s_pk = s_pk; W.s_pkc = s_pkc; W.s_pkc2 = s_pkc2;
    W.s_T = reinterpret_cast<double *>(wsm);
    W.s_rates = W.s_T + n_pad;
    W.s_pos = reinterpret_cast<uint32_t *>(W.s_rates + (in_regs ? 0 : padded));
    W.s_keys = W.s_pos + n_pad; W.s_cnt = W.s_keys + H; W.s_dirty = W.s_cnt + 32; W.s_bits = W.s_dirty + 32;
    W.s_vals = reinterpret_cast<uint16_t *>(W.s_bits + H / 2);
    W.hmask = (uint32_t)H - 1u; W.hshift = 32 - (31 - __clz(H));
    W.lane = lane; W.K = K; W.M = M;
    for (int m = threadIdx.x; m < M; m += kWarps * 32) {
        const int o = X.offsets[m];
        const int dx = (o & 255) - 128, dy = ((o >> 8) & 255) - 128, dz = ((o >> 16) & 255) - 128;
        s_g6[m] = X.g6[m]; s_gd[m] = X.gd[m]; s_off[m] = o; s_lin[m] = X.lin[m];
        s_pk[m] = dx + (dy << 10) + (dz << 20);
        s_pkc[m] = (int32_t)((uint32_t)s_pk[m] * HFXD_C1);                       // (site + pk) * C == site * C + pk * C  (mod 2^32)
        s_pkc2[m] = (int32_t)((uint32_t)s_pk[m] * HFXD_C2);
    }
    if (threadIdx.x < 9) { s_k[threadIdx.x] = X.kf[threadIdx.x]; s_k[9 + threadIdx.x] = X.kd[threadIdx.x]; }
    if (threadIdx.x < 3) { s_k[18 + threadIdx.x] = X.k_isc[threadIdx.x]; s_k[21 + threadIdx.x] = X.k_risc[threadIdx.x]; }
    if (threadIdx.x < 6) { s_k[24 + threadIdx.x] = X.k_rad[threadIdx.x]; s_k[30 + threadIdx.x] = X.k_nonrad[threadIdx.x]; }
    if (threadIdx.x < 4) s_ann[threadIdx.x] = P.ann[threadIdx.x];
    __syncthreads();
    double rr[HFXD_KREG];

    for (int64_t dev = (int64_t)blockIdx.x * kWarps + warp; dev < X.B; dev += (int64_t)gridDim.x * kWarps) {
        const uint32_t ident = X.first_traj + (uint32_t)dev;
        const int64_t r = dev / X.traj_per_real;
        W.ws1 = X.ws1 + r * X.w_stride + X.w_halo; W.wt1 = X.wt1 + r * X.w_stride + X.w_halo;
        uint32_t *g_pos = P.pos + dev * n_pad;
        double *g_T = P.T + dev * n_pad, *g_birth = P.birth + dev * n_pad;
        double time, life;
        uint64_t draws;
        W.s_cnt[lane] = 0; W.s_dirty[lane] = 0;
        for (int h = lane; h < H; h += 32) W.s_keys[h] = 0;
        for (int h = lane; h < H / 2; h += 32) W.s_bits[h] = 0;
        __syncwarp();
        if (P.fill) {
            // hf_xd_spawn: an empty device, slots filled in order; their totals are computed by the refresh pass below
            time = 0.0; life = 0.0; draws = 0;
            for (int q = lane; q < n_pad; q += 32) { W.s_pos[q] = 0; W.s_T[q] = 0.0; g_birth[q] = 0.0; }
            __syncwarp();
            for (int k = 0; k < P.n; ++k) {
                const uint32_t ps = hfxd_replace(W, ident, draws, k);
                if (lane == 0) { W.s_pos[k] = ps; W.s_cnt[HFX_T_GENERATED] += 1; }
            }
            __syncwarp();
            if (lane < n_slots) W.s_dirty[lane] = (lane << 5) + 32 <= P.n ? 0xffffffffu : ((lane << 5) < P.n ? (1u << (P.n & 31)) - 1u : 0u);
        } else {
            time = P.time[dev]; life = P.life[dev]; draws = P.draws[dev];
            for (int q = lane; q < n_pad; q += 32) { W.s_pos[q] = g_pos[q]; W.s_T[q] = g_T[q]; }
            __syncwarp();
            for (int q = lane; q < P.n; q += 32) {                               // the occupancy set of this device, from pos[]
                const uint32_t ps = W.s_pos[q];
                uint32_t h = hfxd_hslot(W, ps & 0x3fffffffu);
                while (atomicCAS(&W.s_keys[h], 0u, ps) != 0u) h = (h + 1) & W.hmask;
                W.s_vals[h] = (uint16_t)q;
            }
            hfxd_bits_rebuild(W, P.n, H);
        }
        __syncwarp();
        const bool traced = X.trace && dev >= X.trace_first && dev < X.trace_first + X.trace_n;
        uint64_t trace_at = traced ? X.trace_len[dev - X.trace_first] : 0;

        // ev = -1: the refresh pass that computes the totals of a freshly filled device (every slot marked dirty)
        for (int64_t ev = P.fill ? -1 : 0; ev < X.n_events; ++ev) {
            // stage 0: evaluate the chosen exciton, pick its channel, apply the event
            //       1: refresh a slot whose own site / spin changed (its evaluation marks the excitons it meets)
            //       2: refresh the slots marked dirty
            if ((ev & 127) == 127) hfxd_bits_rebuild(W, P.n, H);                 // shed the stale bits of erased entries
            int stage = 2, tgt = -1, kx = -1, promoted = -1;
            uint32_t job = 0;
            double target = 0.0, base = 0.0, R = 0.0, u_time = 0.0;
            if (ev >= 0) {
                // (1) the exciton: canonical scan over the cached totals
                double A = 0.0;
                for (int q = 0; q < n_slots; ++q) A = __dadd_rn(A, W.s_T[(q << 5) + lane]);
                const double G = hfx_scan(A, lane);
                R = __shfl_sync(HF_FULL, G, 31);
                const HfPhiloxBlock rnd = hf_philox_block(X.k0, X.k1, HF_STREAM_XDENSE, ident, draws >> 1);
                draws += 2;
                const double u_sel = hf_block_uniform(rnd, 0);
                u_time = hf_block_uniform(rnd, 1);
                target = __dmul_rn(u_sel, R);
                kx = hfxd_pick(W.s_T, n_slots, A, G, target, lane, base);
                stage = 0; tgt = kx; job = W.s_pos[kx];
            }
            for (;;) {
                if (stage == 2) {                                                // next slot marked dirty, ascending
                    int q = -1;
                    for (int wq = 0; wq < n_slots && q < 0; ++wq) {
                        const unsigned todo = W.s_dirty[wq];
                        if (todo) q = (wq << 5) + __ffs(todo) - 1;
                    }
                    if (q < 0) break;
                    tgt = q; job = W.s_pos[q];
                }
                const double S = hfxd_lane_sum(W, job, rr, stage == 0, stage < 2);
                const double L = hfx_scan(S, lane);
                if (stage != 0) {
                    const double tot = __shfl_sync(HF_FULL, L, 31);
                    if (lane == 0) { W.s_T[tgt] = tot; W.s_dirty[tgt >> 5] &= ~(1u << (tgt & 31)); }
                    __syncwarp();
                    if (stage == 1 && promoted >= 0 && tgt != promoted) { tgt = promoted; job = W.s_pos[promoted]; continue; }
                    stage = 2;
                    continue;
                }
                // (2) the channel of the chosen exciton; the excitons its evaluation met see the site that is about to change
                const uint32_t ps = job;
                int chosen;
                if (in_regs) {
                    chosen = hfxd_pick_regs(rr, K, S, L, __dsub_rn(target, base), lane);
                } else {
                    double base2;
                    chosen = hfxd_pick(W.s_rates, K, S, L, __dsub_rn(target, base), lane, base2);
                }
                const double dt = __ddiv_rn(-log(__dsub_rn(1.0, u_time)), R);
                time = __dadd_rn(time, dt);
                // (3) apply
                const uint32_t site = ps & 0x3fffffffu;
                const int spin = (int)(ps >> 30);
                const bool singlet = spin == HFX_SINGLET;
                const int sidx = singlet ? 0 : 1;
                const int px = hf_px((int)site), py = hf_py((int)site), pz = hf_pz((int)site);
                const int64_t i = hf_site(X.X, X.Y, (int)site);
                const int a = hfx_tag_of(__ldg((singlet ? W.ws1 : W.wt1) + i));
                int kind, tally, fin = (int)site;
                bool lost = false;
                uint32_t ps_after = ps;
                if (chosen < M) {
                    const int o = W.s_off[chosen];
                    int xx = px + (o & 255) - 128, yy = py + ((o >> 8) & 255) - 128;
                    if (xx < 0) xx += X.X; else if (xx >= X.X) xx -= X.X;
                    if (yy < 0) yy += X.Y; else if (yy >= X.Y) yy -= X.Y;
                    fin = hf_pack(xx, yy, pz + ((o >> 16) & 255) - 128);
                    const int hj = hfxd_hfind(W, (uint32_t)fin);
                    const int occ_j = hj < 0 ? 0 : (int)(W.s_keys[hj] >> 30);
                    __syncwarp();
                    if (occ_j == 0) {
                        ps_after = (uint32_t)fin | ((uint32_t)spin << 30);
                        if (lane == 0) { hfxd_herase(W, site); hfxd_hinsert(W, ps_after, kx); }
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
                                promoted = (int)W.s_vals[hj];
                                __syncwarp();
                                if (lane == 0) {
                                    W.s_pos[promoted] = (uint32_t)fin | ((uint32_t)HFX_SINGLET << 30);
                                    W.s_keys[hj] = (uint32_t)fin | ((uint32_t)HFX_SINGLET << 30);
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
                    ps_after = site | ((uint32_t)flipped << 30);
                    __syncwarp();
                    if (lane == 0) hfxd_hset_spin(W, site, flipped);
                } else {
                    kind = HFX_K_DECAY;
                    tally = (chosen == M + 1 ? HFX_T_RAD : (singlet ? HFX_T_NONRAD_S : HFX_T_NONRAD_T)) + a;
                    lost = true;
                }
                if (lost) {
                    __syncwarp();
                    if (lane == 0) hfxd_herase(W, site);
                    life = __dadd_rn(life, __dsub_rn(time, g_birth[kx]));        // every lane reads the same word
                    __syncwarp();
                    ps_after = hfxd_replace(W, ident, draws, kx);
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
                // (4) refresh the cached totals: the slot itself in its new state, a promoted acceptor (whose neighbours
                //     see a different spin), then everybody marked
                stage = 1; tgt = kx; job = ps_after;
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
