// Exciton trajectories ("Path X"): warp-per-exciton rate evaluation, rejection-free
// (BKL) selection by a warp-shuffle inclusive prefix scan over the transfer / spin-flip / decay rates.
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
// (dy, dz) row of the cutoff sphere as one contiguous run.  Each realisation's weight arrays carry
// a halo of R zero planes below z = 0 and above z = Z - 1: a transfer through an open face reads a
// zero weight and gets rate 0.0 without a bounds test.
//
// Canonical summation order (specification): channel c -> lane c % 32, slot c / 32; each lane
// sums its slots in order (S_l), one Kogge-Stone scan over the 32 lane sums gives L_l and the total
// R; the chosen lane is the first with L_l > u_sel * R, and inside it the prefix restarts from
// L_{l-1} and walks the lane's slots.  One scan per event instead of one per 32 channels.
//
// Execution shape (v3).  A warp owns 32 trajectories at a time and alternates two phases per event:
//   A  warp per exciton: for e = 0..31 the 32 lanes evaluate exciton e's M + 3 rates (lane = channel
//      % 32, coalesced gathers, distance factors and linear offsets of the lane's channels held in
//      registers, the next exciton's gathers in flight while this one's are consumed), reduce them
//      to lane sums, scan them with shuffles and park the 32 prefix sums in shared memory;
//   B  thread per exciton: lane e draws exciton e's Philox block, finds the lane and the slot of
//      the chosen channel (re-evaluating only that lane's K rates, bit-identically), advances the
//      clock and applies the event.
// Everything that is per-event overhead (Philox, log, the divisions, bookkeeping, tallies) thereby
// costs one warp instruction per 32 events instead of one per event: 1033 -> 297 warp
// instructions per event at M = 256 (profiles/r01e_* vs r01f_*).
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
    int32_t block;                    // trajectories a warp owns at a time (1..32)
    const int32_t *offsets;           // [M] packed (dx+128) | (dy+128) << 8 | (dz+128) << 16
    const int32_t *lin;               // [M] dz*X*Y + dy*X + dx; brick layout: hfx_brick_lin()
    int32_t brick;                    // the weights are the brick-layout copy: planes padded by R periodic columns / rows on
    int32_t bX, bY;                   // every side (then up to multiples of 4): bX x bY, in 4 x 4 x 1 bricks (hfx_brick2)
    const double *g6, *gd;            // [M]
    double kf[9], kd[9];              // [donor][acceptor]
    double k_isc[3], k_risc[3], k_rad[6], k_nonrad[6];
    double singlet_fraction;
    const double *ws1, *wt1;          // [n_real][w_stride] tagged weights; site 0 of a realisation at + w_halo
    int64_t w_stride, w_halo;         // w_stride = N + 2 * w_halo, w_halo = R planes of zeros on either side in z
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

// Brick layout of a z-plane (used when the weights do not fit the L2): 4 x 4 bricks of 16 doubles = one 128-byte
// line, bricks row-major.  A cutoff-4 sphere then touches 39.5 lines on average instead of the 61.9 of the row-major
// plane, and the DRAM-bound gather runs 1.63x faster (scripts/micro/brick_layout.cu, profiles/r02_micro_brick_layout.txt).
__host__ __device__ __forceinline__ int hfx_brick2(int x, int y, int dim_x) {
    return (((y >> 2) * (dim_x >> 2) + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3);
}
// The brick copy pads every plane with R periodic columns and rows on each side (site (x, y) sits at (x + R, y + R)),
// so that NO site needs the periodic wrap: the kernel has one gather path and the offsets below hold everywhere.
// Offset of the neighbour (dx, dy, dz) of a site, brick layout.  With x + R = 4 bx + rx:
//   delta = dz * plane + 4 dy + dx + 12 * ((rx + dx) >> 2) + (4 X - 16) * ((ry + dy) >> 2)
// and (rx + dx) >> 2 = (dx >> 2) + ((rx + (dx & 3)) >> 2), the last term being 0 or 1.  Packed per channel:
// bits 8.. the part that does not depend on (rx, ry); bit rx: add 12; bit 4 + ry: add 4 X - 16.
__host__ __device__ inline int32_t hfx_brick_lin(int dx, int dy, int dz, int dim_x, int dim_y) {
    const int c16 = 4 * dim_x - 16;
    const int base = dz * dim_x * dim_y + 4 * dy + dx + 12 * (dx >> 2) + c16 * (dy >> 2);
    unsigned flags = 0;
    for (int r = 0; r < 4; ++r) flags |= (unsigned)((r + (dx & 3)) >> 2) << r | (unsigned)((r + (dy & 3)) >> 2) << (4 + r);
    return (int32_t)(((uint32_t)base << 8) | flags);
}

// ws1 / wt1 must be zero-filled beforehand (the halo planes stay +0.0, species tag 0)
__global__ void hfx_weights_kernel(const uint8_t *species, const double *s1, const double *t1, double *ws1, double *wt1,
                                   int64_t n_sites, int64_t n_real, int64_t w_stride, int64_t w_halo) {
    const int64_t n = n_sites * n_real;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = (i / n_sites) * w_stride + w_halo + i % n_sites;
        ws1[o] = hfx_tag(exp(-(s1[i] / HF_KT)), species[i]);
        wt1[o] = hfx_tag(exp(-(t1[i] / HF_KT)), species[i]);
    }
}
// The brick-layout copy: bX x bY padded planes (periodic images of the X x Y plane shifted by R), same z halo.
__global__ void hfx_brick_weights_kernel(const uint8_t *species, const double *s1, const double *t1, double *ws1, double *wt1,
                                         int X, int Y, int Z, int R, int bX, int bY, int64_t n_real, int64_t w_stride, int64_t w_halo) {
    const int64_t bplane = (int64_t)bX * bY, per_real = bplane * Z, n = per_real * n_real, n_sites = (int64_t)X * Y * Z;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / per_real, in = i % per_real;
        const int z = (int)(in / bplane), yp = (int)((in % bplane) / bX), xp = (int)(in % bX);
        const int x = ((xp - R) % X + X) % X, y = ((yp - R) % Y + Y) % Y;
        const int64_t src = r * n_sites + ((int64_t)z * Y + y) * X + x;
        const int64_t o = r * w_stride + w_halo + z * bplane + hfx_brick2(xp, yp, bX);
        ws1[o] = hfx_tag(exp(-(s1[src] / HF_KT)), species[src]);
        wt1[o] = hfx_tag(exp(-(t1[src] / HF_KT)), species[src]);
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

// Per-warp shared memory (one block of 32 trajectories), struct-of-arrays so that lane e writes and the
// warp broadcasts without bank conflicts:
//   L[32][33] f64 prefix sums | wp[32] u64 | inv[32] f64 | onsite[3][32] f64 | sm[32] int2 (site, meta) |
//   tallies[HFX_T_N][32] u32 | nz[32] u32
#define HFX_L_STRIDE 33               // doubles per exciton row of prefix sums (odd: conflict-free column reads)
#define HFX_N_K 36                    // kf[9] kd[9] k_isc[3] k_risc[3] k_rad[6] k_nonrad[6]
#define HFX_W_L 0
#define HFX_W_WP (8 * 32 * HFX_L_STRIDE)
#define HFX_W_INV (HFX_W_WP + 8 * 32)
#define HFX_W_ON (HFX_W_INV + 8 * 32)
#define HFX_W_SM (HFX_W_ON + 8 * 96)
#define HFX_W_CNT (HFX_W_SM + 8 * 32)
#define HFX_W_NZ (HFX_W_CNT + 4 * 32 * HFX_T_N)
#define HFX_W_BYTES (HFX_W_NZ + 4 * 32)

__host__ __device__ inline size_t hfx_smem_bytes(int M, int warps) {
    // CTA-wide tables: g6, gd, rate constants (f64), lin, off (i32, M rounded up to even) | per-warp blocks
    return 8 * (2 * (size_t)M + HFX_N_K) + 4 * 2 * (size_t)((M + 1) & ~1) + (size_t)warps * HFX_W_BYTES;
}

// Keep a value in a register: under register pressure nvcc otherwise rematerialises shared-memory
// bases (S2R + LEA chains) inside the event loop, which costs more issue slots than the register.
#if defined(__CUDA_ARCH__)
#define HFX_PIN32(x) asm volatile("" : "+r"(x))
#define HFX_PIN64(x) asm volatile("" : "+l"(x))
#else
#define HFX_PIN32(x) (void)0
#define HFX_PIN64(x) (void)0
#endif

// min(v, 1) as `v < 1.0 ? v : 1.0`, decided on the high word: for every double (negative, NaN and infinity included)
// v < 1.0 <=> (int)hi(v) < 0x3ff00000, and the integer compare is 3 instructions where DSETP.MIN + fix-ups are 6.
// Read-only shared-memory tables through 32-bit shared addresses held in registers (the hot loop of the dense
// kernel): a generic pointer pinned in a register makes nvcc emit generic loads, an unpinned one makes it rebuild the
// shared-window base at every use.  Only for data that is not written after the kernel's first barrier.
#if defined(__CUDA_ARCH__)
typedef unsigned hfx_sptr;
__device__ __forceinline__ hfx_sptr hfx_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double hfx_lds_f64(hfx_sptr a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int hfx_lds_s32(hfx_sptr a) { int v; asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#else
typedef uintptr_t hfx_sptr;
inline hfx_sptr hfx_saddr(const void *p) { return (uintptr_t)p; }
inline double hfx_lds_f64(hfx_sptr a) { return *(const double *)a; }
inline int hfx_lds_s32(hfx_sptr a) { return *(const int *)a; }
#endif

__device__ __forceinline__ double hfx_min1(double v) { return __double2hiint(v) < 0x3ff00000 ? v : 1.0; }

// Offset (in sites) from an exciton on (px, py, .) to the neighbour of packed offset o: periodic in
// x and y, straight through in z (the halo absorbs z outside [0, Z)).
template <bool kBrick = false>
__device__ __forceinline__ int hfx_delta(int o, int px, int py, int dim_x, int dim_y, int plane) {
    const int dx = (o & 255) - 128, dy = ((o >> 8) & 255) - 128, dz = ((o >> 16) & 255) - 128;
    int xx = px + dx, yy = py + dy;
    if (kBrick) return dz * plane + hfx_brick2(xx, yy, dim_x) - hfx_brick2(px, py, dim_x);   // padded planes (px, py include the pad): no wrap
    if (xx < 0) xx += dim_x; else if (xx >= dim_x) xx -= dim_x;
    if (yy < 0) yy += dim_y; else if (yy >= dim_y) yy -= dim_y;
    return dz * plane + (yy - py) * dim_x + (xx - px);
}

// rate of one transfer channel, the arithmetic of the specification: (pref[b] * g) * min(1, wj / wi)
__device__ __forceinline__ double hfx_rate(const double *pref, double g, double wj, double inv_wi) {
    return __dmul_rn(__dmul_rn(pref[hfx_tag_of(wj)], g), hfx_min1(__dmul_rn(wj, inv_wi)));
}

// inclusive Kogge-Stone scan over the 32 lane sums
__device__ __forceinline__ double hfx_scan(double S, int lane) {
    double L = S;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double up = __shfl_up_sync(HF_FULL, L, off);
        if (lane >= off) L = __dadd_rn(L, up);
    }
    return L;
}

template <int KF>
struct HfxLaneTables {                // this lane's channels of the KF full slots, in registers
    int32_t lin[KF > 0 ? KF : 1];     // linear site offsets (valid when no periodic wrap is needed)
    double gd[KF > 0 ? KF : 1];       // Dexter distance factors: triplets are the bulk of the events (they live 10^3 x longer)
};

// Phase A, loads: the weights of the neighbours on this lane's channels of the full slots.
// (slots kBase .. kBase + CH - 1: a kernel whose KF slots do not fit two register buffers gathers them in chunks)
template <int KF, int CH = KF, int kBase = 0, bool kBrick = false>
__device__ __forceinline__ void hfx_gather(const double *wp, int site, int meta, const HfxLaneTables<KF> &T,
                                           const int32_t *s_off_lane, int dim_x, int dim_y, int plane,
                                           double (&buf)[CH > 0 ? CH : 1], int pad = 0) {
    if (kBrick || (meta & 8)) {
        if (kBrick) {
            const int mx = 1 << ((hf_px(site) + pad) & 3), my = 16 << ((hf_py(site) + pad) & 3), c16 = 4 * dim_x - 16;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int t = T.lin[kBase + k];
                int d = t >> 8;
                if (t & mx) d += 12;
                if (t & my) d += c16;
                buf[k] = __ldg(wp + d);
            }
        } else {
#pragma unroll
            for (int k = 0; k < CH; ++k) buf[k] = __ldg(wp + T.lin[kBase + k]);
        }
    } else {
        const int px = hf_px(site), py = hf_py(site);
#pragma unroll
        for (int k = 0; k < CH; ++k) buf[k] = __ldg(wp + hfx_delta<kBrick>(s_off_lane[32 * (kBase + k)], px, py, dim_x, dim_y, plane));
    }
}

// KF: the first KF slots of 32 channels (KF <= M / 32) have register-resident tables and prefetched gathers
// (kRegG: the Dexter factors too, else they are read from shared memory); the slots after them — the rest of
// the transfer channels and the three on-site ones — go through the shared-memory tables.  KF == 0: everything
// does.  kTail >= 0: M = 32 * KF + kTail is known at compile time (the default cutoffs 3, 4, 5); kTail < 0: M = P.M.
// kChunk: slots gathered at a time (KF, or KF / 2 above 8 slots: two register buffers of 16 would spill).
// kBrick: the weights are stored in 4 x 4 x 1 bricks (hfx_brick2) and P.lin holds hfx_brick_lin() words.
template <int kWarps, int KF, int kTail, int kMinBlocks, bool kRegG, int kChunk = KF, bool kBrick = false>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) hfx_run_kernel(const HfxParams P) {
    static_assert(kChunk == KF || 2 * kChunk == KF, "one or two chunks");
    extern __shared__ __align__(16) unsigned char hf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = (KF > 0 && kTail >= 0) ? 32 * KF + kTail : P.M;
    const int K = (M + 3 + 31) >> 5;
    constexpr int KB = kChunk > 0 ? kChunk : 1;
    double *s_g6 = reinterpret_cast<double *>(hf_smem);
    double *s_gd = s_g6 + M;
    double *s_k = s_gd + M;
    int32_t *s_lin = reinterpret_cast<int32_t *>(s_k + HFX_N_K);
    int32_t *s_off = s_lin + ((M + 1) & ~1);
    unsigned wofs = (unsigned)(8 * (2 * M + HFX_N_K) + 8 * ((M + 1) & ~1)) + (unsigned)warp * HFX_W_BYTES;
    HFX_PIN32(wofs);
    unsigned char *const wsm = hf_smem + wofs;                                  // this warp's block
    double *const s_L = reinterpret_cast<double *>(wsm + HFX_W_L);
    const double **const s_wp = reinterpret_cast<const double **>(wsm + HFX_W_WP);
    double *const s_inv = reinterpret_cast<double *>(wsm + HFX_W_INV);
    double *const s_on = reinterpret_cast<double *>(wsm + HFX_W_ON);
    int2 *const s_sm = reinterpret_cast<int2 *>(wsm + HFX_W_SM);
    unsigned *const s_cnt = reinterpret_cast<unsigned *>(wsm + HFX_W_CNT);
    unsigned *const s_nz = reinterpret_cast<unsigned *>(wsm + HFX_W_NZ);
    for (int m = threadIdx.x; m < M; m += kWarps * 32) { s_g6[m] = P.g6[m]; s_gd[m] = P.gd[m]; s_lin[m] = P.lin[m]; s_off[m] = P.offsets[m]; }
    if (threadIdx.x < 9) { s_k[threadIdx.x] = P.kf[threadIdx.x]; s_k[9 + threadIdx.x] = P.kd[threadIdx.x]; }
    if (threadIdx.x < 3) { s_k[18 + threadIdx.x] = P.k_isc[threadIdx.x]; s_k[21 + threadIdx.x] = P.k_risc[threadIdx.x]; }
    if (threadIdx.x < 6) { s_k[24 + threadIdx.x] = P.k_rad[threadIdx.x]; s_k[30 + threadIdx.x] = P.k_nonrad[threadIdx.x]; }
    __syncthreads();
    const int pad = kBrick ? P.R : 0;                                           // brick copy: site (x, y) sits at (x + R, y + R) of a bX x bY plane
    const int dim_x = kBrick ? P.bX : P.X, dim_y = kBrick ? P.bY : P.Y, plane = dim_x * dim_y;
    HfxLaneTables<KF> T;
#pragma unroll
    for (int k = 0; k < KF; ++k) { T.lin[k] = s_lin[32 * k + lane]; T.gd[k] = kRegG ? s_gd[32 * k + lane] : 0.0; }
    const int32_t *s_off_lane = s_off + lane;
    const double *s_g6_lane = s_g6 + lane, *s_gd_lane = s_gd + lane;

    const int bs = P.block;
    const int64_t n_blocks = (P.B + bs - 1) / bs;
    for (int64_t blk = (int64_t)blockIdx.x * kWarps + warp; blk < n_blocks; blk += (int64_t)gridDim.x * kWarps) {
        const int64_t local = blk * bs + lane;
        const int count = (int)(P.B - blk * bs < bs ? P.B - blk * bs : bs);
        const bool active = lane < count;
        const uint32_t ident = P.first_traj + (uint32_t)local;
        const int64_t r = (active ? local : blk * bs) / P.traj_per_real;
        const double *ws1 = P.ws1 + r * P.w_stride + P.w_halo, *wt1 = P.wt1 + r * P.w_stride + P.w_halo;
        int site = 0, spin = HFX_SINGLET;
        double time = 0.0, lifetime = 0.0;
        uint64_t draws = 0;
        if (active) { site = P.site[local]; spin = P.spin[local]; time = P.time[local]; lifetime = P.lifetime[local]; draws = P.draws[local]; }
        const bool traced = active && P.trace && local >= P.trace_first && local < P.trace_first + P.trace_n;
#pragma unroll
        for (int t = 0; t < HFX_T_N; ++t) s_cnt[t * 32 + lane] = 0;

        for (int64_t ev = 0; ev < P.n_events; ++ev) {
            // ---- this lane's exciton: what phase A needs to know about it ----
            if (active) {
                const bool singlet = spin == HFX_SINGLET;
                const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
                const int64_t i = (int64_t)pz * plane + (kBrick ? hfx_brick2(px + pad, py + pad, dim_x) : py * dim_x + px);
                const double *wp = (singlet ? ws1 : wt1) + i;
                const double wi = __ldg(wp);
                const int a = hfx_tag_of(wi);
                const double inv_wi = __ddiv_rn(1.0, wi);
                const bool nowrap = px >= P.R && px < P.X - P.R && py >= P.R && py < P.Y - P.R;
                s_wp[lane] = wp;
                s_inv[lane] = inv_wi;
                s_sm[lane] = make_int2(site, (singlet ? 1 : 0) | (a << 1) | (nowrap ? 8 : 0));
                s_on[lane] = singlet ? s_k[18 + a] : __dmul_rn(s_k[21 + a], hfx_min1(__dmul_rn(__ldg(ws1 + i), inv_wi)));   // inv_wi = 1/wt1[i] for a triplet
                s_on[32 + lane] = s_k[24 + 2 * a + (singlet ? 0 : 1)];
                s_on[64 + lane] = s_k[30 + 2 * a + (singlet ? 0 : 1)];
            }
            __syncwarp();

            // ---- phase A: the warp evaluates one exciton at a time; the scan of exciton e - 1 is issued
            //      together with the rates of exciton e (independent chains), the gathers of e + 1 before both ----
            double buf0[KB], buf1[KB];
            double S_prev = 0.0;
            if (KF > 0) { const int2 sm0 = s_sm[0]; hfx_gather<KF, kChunk, 0, kBrick>(s_wp[0], sm0.x, sm0.y, T, s_off_lane, dim_x, dim_y, plane, buf0, pad); }
            // one exciton: `cur` holds its gathered weights, the next exciton's go into `nxt`
#define HFX_PHASE_A(e, cur, nxt)                                                                                              \
            {                                                                                                                 \
                const int2 sm = s_sm[e];                                                                                      \
                const int meta = sm.y;                                                                                        \
                const double e_inv = s_inv[e];                                                                                \
                const double *e_pref = s_k + ((meta & 1) ? 0 : 9) + 3 * ((meta >> 1) & 3);                                    \
                if (KF > 0 && (e) + 1 < count) {                                                                              \
                    const int2 smn = s_sm[(e) + 1];                                                                           \
                    hfx_gather<KF, KF, 0, kBrick>(s_wp[(e) + 1], smn.x, smn.y, T, s_off_lane, dim_x, dim_y, plane, nxt, pad);                     \
                }                                                                                                             \
                double S = 0.0, L_prev;                                                                                       \
                if (meta & 1) {                                                                                               \
                    _Pragma("unroll") for (int k = 0; k < KF; ++k) S = __dadd_rn(S, hfx_rate(e_pref, s_g6_lane[32 * k], cur[k], e_inv)); \
                    L_prev = hfx_scan(S_prev, lane);                                                                          \
                } else {                                                                                                      \
                    _Pragma("unroll") for (int k = 0; k < KF; ++k) S = __dadd_rn(S, hfx_rate(e_pref, kRegG ? T.gd[k] : s_gd_lane[32 * k], cur[k], e_inv)); \
                    L_prev = hfx_scan(S_prev, lane);                                                                          \
                }                                                                                                             \
                if ((e) > 0) {                                                                                                \
                    s_L[((e) - 1) * HFX_L_STRIDE + lane] = L_prev;                                                            \
                    const unsigned nonzero = __ballot_sync(HF_FULL, S_prev > 0.0);                                            \
                    if (lane == 0) s_nz[(e) - 1] = nonzero;                                                                   \
                }                                                                                                             \
                if (K > KF) {   /* slots not covered above: the remaining transfer channels and the three on-site ones */    \
                    const double *g = (meta & 1) ? s_g6 : s_gd;                                                               \
                    const double *e_wp = s_wp[e];                                                                             \
                    const int ex = hf_px(sm.x) + pad, ey = hf_py(sm.x) + pad;                                                             \
                    for (int k = KF; k < K; ++k) {                                                                            \
                        const int c = (k << 5) + lane;                                                                        \
                        double rate = 0.0;                                                                                    \
                        if (c < M) rate = hfx_rate(e_pref, g[c], __ldg(e_wp + hfx_delta<kBrick>(s_off[c], ex, ey, dim_x, dim_y, plane)), e_inv); \
                        else if (c < M + 3) rate = s_on[(c - M) * 32 + (e)];                                                  \
                        S = __dadd_rn(S, rate);                                                                               \
                    }                                                                                                         \
                }                                                                                                             \
                S_prev = S;                                                                                                   \
            }
            // the 16-slot variant: two halves per exciton; the first half's sums overlap the scan of exciton e - 1, the
            // second half of e (then the first half of e + 1) is in flight while a half is consumed
#define HFX_HALF(e, kBase, cur, nxt)                                                                                          \
            {                                                                                                                 \
                const int2 sm = s_sm[e];                                                                                      \
                const int meta = sm.y;                                                                                        \
                const double e_inv = s_inv[e];                                                                                \
                const double *e_pref = s_k + ((meta & 1) ? 0 : 9) + 3 * ((meta >> 1) & 3);                                    \
                if (kBase == 0) {                                                                                             \
                    hfx_gather<KF, kChunk, kChunk, kBrick>(s_wp[e], sm.x, sm.y, T, s_off_lane, dim_x, dim_y, plane, nxt, pad);             \
                } else if ((e) + 1 < count) {                                                                                 \
                    const int2 smn = s_sm[(e) + 1];                                                                           \
                    hfx_gather<KF, kChunk, 0, kBrick>(s_wp[(e) + 1], smn.x, smn.y, T, s_off_lane, dim_x, dim_y, plane, nxt, pad);          \
                }                                                                                                             \
                double S = kBase == 0 ? 0.0 : S_half;                                                                         \
                if (meta & 1) {                                                                                               \
                    _Pragma("unroll") for (int k = 0; k < kChunk; ++k) S = __dadd_rn(S, hfx_rate(e_pref, s_g6_lane[32 * (kBase + k)], cur[k], e_inv)); \
                } else {                                                                                                      \
                    _Pragma("unroll") for (int k = 0; k < kChunk; ++k) S = __dadd_rn(S, hfx_rate(e_pref, kRegG ? T.gd[kBase + k] : s_gd_lane[32 * (kBase + k)], cur[k], e_inv)); \
                }                                                                                                             \
                if (kBase == 0) {                                                                                             \
                    S_half = S;                                                                                               \
                    const double L_prev = hfx_scan(S_prev, lane);                                                             \
                    if ((e) > 0) {                                                                                            \
                        s_L[((e) - 1) * HFX_L_STRIDE + lane] = L_prev;                                                        \
                        const unsigned nonzero = __ballot_sync(HF_FULL, S_prev > 0.0);                                        \
                        if (lane == 0) s_nz[(e) - 1] = nonzero;                                                               \
                    }                                                                                                         \
                } else {                                                                                                      \
                    const double *g = (meta & 1) ? s_g6 : s_gd;                                                               \
                    const double *e_wp = s_wp[e];                                                                             \
                    const int ex = hf_px(sm.x) + pad, ey = hf_py(sm.x) + pad;                                                             \
                    for (int k = KF; k < K; ++k) {                                                                            \
                        const int c = (k << 5) + lane;                                                                        \
                        double rate = 0.0;                                                                                    \
                        if (c < M) rate = hfx_rate(e_pref, g[c], __ldg(e_wp + hfx_delta<kBrick>(s_off[c], ex, ey, dim_x, dim_y, plane)), e_inv); \
                        else if (c < M + 3) rate = s_on[(c - M) * 32 + (e)];                                                  \
                        S = __dadd_rn(S, rate);                                                                               \
                    }                                                                                                         \
                    S_prev = S;                                                                                               \
                }                                                                                                             \
            }
            if constexpr (kChunk == KF) {
                for (int e = 0; e < count; e += 2) {
                    HFX_PHASE_A(e, buf0, buf1)
                    if (e + 1 < count) HFX_PHASE_A(e + 1, buf1, buf0)
                }
            } else {
                double S_half = 0.0;
                for (int e = 0; e < count; ++e) {
                    HFX_HALF(e, 0, buf0, buf1)
                    HFX_HALF(e, kChunk, buf1, buf0)
                }
            }
#undef HFX_PHASE_A
#undef HFX_HALF
            {
                const double L_last = hfx_scan(S_prev, lane);
                s_L[(count - 1) * HFX_L_STRIDE + lane] = L_last;
                const unsigned nonzero = __ballot_sync(HF_FULL, S_prev > 0.0);
                if (lane == 0) s_nz[count - 1] = nonzero;
            }
            __syncwarp();

            // ---- phase B: every lane selects and applies its own exciton's event ----
            if (active) {
                const bool singlet = spin == HFX_SINGLET;
                const int px = hf_px(site), py = hf_py(site), pz = hf_pz(site);
                const double *wp = s_wp[lane];
                const double inv_wi = s_inv[lane];
                const int a = (s_sm[lane].y >> 1) & 3;
                const double *pref = s_k + (singlet ? 0 : 9) + 3 * a;
                const HfPhiloxBlock rnd = hf_philox_block(P.k0, P.k1, HF_STREAM_XTRAJ, ident, draws >> 1);
                draws += 2;
                const double u_sel = hf_block_uniform(rnd, 0), u_time = hf_block_uniform(rnd, 1);
                const double *row = s_L + lane * HFX_L_STRIDE;
                const double R = row[31];
                const double target = __dmul_rn(u_sel, R);
                int lstar = -1;
#pragma unroll 8
                for (int l = 31; l >= 0; --l) if (row[l] > target) lstar = l;     // first lane whose prefix passes the target
                if (lstar < 0) { const unsigned nz = s_nz[lane]; lstar = nz ? 31 - __clz(nz) : 0; }
                double acc = lstar > 0 ? row[lstar - 1] : 0.0;
                int slot = -1, last_live = -1;
                const double *g = singlet ? s_g6 : s_gd;
                for (int k = 0; k < K; ++k) {                                     // the chosen lane's rates again, same inputs, same arithmetic
                    const int c = (k << 5) + lstar;
                    double rate = 0.0;
                    if (c < M) rate = hfx_rate(pref, g[c], __ldg(wp + hfx_delta<kBrick>(s_off[c], px + pad, py + pad, dim_x, dim_y, plane)), inv_wi);
                    else if (c < M + 3) rate = s_on[(c - M) * 32 + lane];
                    acc = __dadd_rn(acc, rate);
                    if (rate > 0.0) last_live = k;
                    if (slot < 0 && acc > target) slot = k;
                }
                if (slot < 0) slot = last_live < 0 ? K - 1 : last_live;
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
                    kind = HFX_K_ISC;                                             // molecule.py:299-307
                    tally = singlet ? HFX_T_ISC : HFX_T_RISC;
                    spin = singlet ? HFX_TRIPLET : HFX_SINGLET;
                } else {
                    kind = HFX_K_DECAY;
                    tally = (chosen == M + 1 ? HFX_T_RAD : (singlet ? HFX_T_NONRAD_S : HFX_T_NONRAD_T)) + a;
                }
                int site_after = fin;
                s_cnt[HFX_T_EVENTS * 32 + lane] += 1;
                s_cnt[tally * 32 + lane] += 1;
                if (kind == HFX_K_DECAY) {
                    lifetime = __dadd_rn(lifetime, time);
                    hfx_spawn(P, ident, draws, site_after, spin, time);
                    s_cnt[HFX_T_GENERATED * 32 + lane] += 1;
                }
                if (traced) {
                    const uint64_t at = P.trace_len[local - P.trace_first];
                    if (at < (uint64_t)P.trace_cap) {
                        HfTraceRec rec;
                        rec.initial = (int32_t)hf_site(P.X, P.Y, site); rec.final_ = (int32_t)hf_site(P.X, P.Y, fin);
                        rec.tau = dt; rec.kind = (uint8_t)kind; rec.particule = HF_PART_EXCITON;
                        rec.pad[0] = (uint8_t)spin; rec.pad[1] = 0;     // spin of the trajectory's exciton after the step
                        rec.spins = (uint32_t)chosen; rec.draws = draws;
                        P.trace[(local - P.trace_first) * P.trace_cap + at] = rec;
                    }
                    P.trace_len[local - P.trace_first] = at + 1;
                }
                site = site_after;
            }
        }
        if (active) { P.site[local] = site; P.spin[local] = spin; P.time[local] = time; P.lifetime[local] = lifetime; P.draws[local] = draws; }
        // tallies: one atomic per counter and warp when the 32 trajectories share a realisation
        const int64_t r_first = __shfl_sync(HF_FULL, r, 0), r_last = __shfl_sync(HF_FULL, r, count - 1);
        const bool uniform = r_first == r_last;
        for (int t = 0; t < HFX_T_N; ++t) {
            unsigned v = s_cnt[t * 32 + lane];
            if (uniform) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(HF_FULL, v, off);
                if (lane == 0 && v) atomicAdd(P.totals + r * HFX_T_N + t, (unsigned long long)v);
            } else if (v) {
                atomicAdd(P.totals + r * HFX_T_N + t, (unsigned long long)v);
            }
        }
        __syncwarp();
    }
}
