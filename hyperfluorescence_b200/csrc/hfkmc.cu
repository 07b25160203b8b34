// libhfkmc.so — host side of the C-ABI declared in include/hfkmc.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared ...
// (see hyperfluorescence_b200/build.py).  No CPU path exists: without an sm_100 device
// hf_create fails with HF_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/hfkmc.h"
#include <stdlib.h>
#include "hf_frm.cuh"
#include "hf_frm_c1.cuh"
#include "hf_frm_small.cuh"
#include "hf_exciton.cuh"
#include "hf_exciton_dense.cuh"

static_assert(sizeof(HfTraceRec) == sizeof(hf_event), "trace record must match hf_event");
static_assert(sizeof(hf_event) == 32, "hf_event is 32 bytes");

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define HF_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(HF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kMaxDim = 1023;          // packed coordinates: 10 bits per axis

}  // namespace

struct hf_sim {
    hf_config cfg;
    HfParams P;
    int device = 0;
    int n_real = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool have_lattice = false, injected = false;
    std::vector<char> have_st;         // per realisation: S1 / T1 energies were generated or uploaded
    // lattice
    uint8_t *d_species = nullptr;
    double *d_homo = nullptr, *d_lumo = nullptr, *d_s1 = nullptr, *d_t1 = nullptr;
    // state
    int32_t *d_pos = nullptr, *d_flag = nullptr, *d_ev_init = nullptr, *d_ev_final = nullptr;
    double *d_ev_tau = nullptr;
    HfScalars *d_sc = nullptr;
    int32_t *d_cache_i = nullptr;      // init | final | kp, each [B][10]
    double *d_cache_tau = nullptr;
    unsigned long long *d_totals = nullptr, *d_sum = nullptr;
    double *d_replay = nullptr;
    HfTraceRec *d_trace = nullptr;
    uint64_t *d_trace_base = nullptr;
    std::vector<uint64_t> trace_base;  // host copy of d_trace_base
    // launch shape of the step kernel
    int warps = 4, grid = 1, resident_warps = 0;
    size_t smem = 0;
    // exciton path (hf_x_*)
    bool x_configured = false, x_spawned = false;
    HfxParams X;
    hf_x_params x_params;
    std::vector<int8_t> x_offsets;
    std::vector<double> x_g6, x_gd;
    int32_t *d_x_off = nullptr, *d_x_lin = nullptr, *d_x_site = nullptr, *d_x_spin = nullptr;
    double *d_x_g6 = nullptr, *d_x_gd = nullptr, *d_ws1 = nullptr, *d_wt1 = nullptr, *d_x_time = nullptr, *d_x_life = nullptr;
    uint64_t *d_x_draws = nullptr, *d_x_trace_len = nullptr;
    unsigned long long *d_x_totals = nullptr;
    int x_warps = 8, x_grid = 1, x_per_sm = 1;
    bool x_tune = false;               // pick the warp batch by measurement at the next hf_x_run
    std::unordered_map<void *, size_t> alloc_bytes;   // buffers managed by dev_reuse
    uint8_t *d_gen_layer = nullptr;    // scratch of lattice generation and of the per-realisation reduction, kept between
    int32_t *d_gen_pool = nullptr;     // calls: cudaMalloc + cudaFree per GA sweep cost 0.1-0.9 s every few calls
    int64_t *d_gen_counts = nullptr;
    unsigned long long *d_real_out = nullptr;
    bool x_brick = false;              // hfx_run_kernel reads brick-layout copies of the weights
    int32_t x_bX = 0, x_bY = 0;        // their padded plane
    double *d_bs1 = nullptr, *d_bt1 = nullptr;
    int32_t *d_x_blin = nullptr;
    size_t x_smem = 0;
    // high-density exciton devices (hf_xd_*)
    bool xd_configured = false, xd_spawned = false;
    HfxdParams XD;
    uint32_t *d_xd_rec_seen = nullptr;
    uint32_t *d_xd_pos = nullptr, *d_xd_occ = nullptr;
    double *d_xd_T = nullptr, *d_xd_birth = nullptr, *d_xd_rec_L = nullptr, *d_xd_time = nullptr, *d_xd_life = nullptr;
    uint64_t *d_xd_draws = nullptr;
    unsigned long long *d_xd_totals = nullptr;
    int xd_grid = 1;
    size_t xd_smem = 0;
    // integrated excitons (hf_p_excitons)
    int32_t *d_pi_site = nullptr, *d_pi_meta = nullptr, *d_pi_final = nullptr, *d_pi_n = nullptr;
    double *d_pi_tau = nullptr, *d_pi_clock = nullptr;
    uint64_t *d_pi_draws = nullptr;
    HfiTables *d_pi_tab = nullptr;
    double *d_inv_dist = nullptr;      // 1 / sqrt(d2) table of the warp kernel
    // timing of hf_run launches
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;
    std::vector<cudaEvent_t> event_pool;
    double timed_ms = 0.0;
    int64_t timed_launches = 0;
};

namespace {

template <int W>
int configure_launch(hf_sim *s) {
    const size_t per_warp = hf_warp_smem_bytes(s->P.cap_c, s->P.cap_f, hf_xcap(s->P));
    const size_t smem = per_warp * W;
    if (smem > 48 * 1024) {
        HF_CUDA(cudaFuncSetAttribute(hf_run_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HF_CUDA(cudaFuncSetAttribute(hf_run_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HF_CUDA(cudaFuncSetAttribute(hf_inject_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int per_sm = 0;
    if (s->P.xi.mode) HF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hf_run_kernel<W, true>, W * 32, smem));
    else HF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hf_run_kernel<W, false>, W * 32, smem));
    if (per_sm < 1) return fail(HF_ERR_UNSUPPORTED, "step kernel does not fit an SM (charges=%d)", s->P.C);
    const int64_t need = (s->P.B + W - 1) / W;
    const int64_t cap = (int64_t)s->sm_count * per_sm;
    s->warps = W;
    s->smem = smem;
    s->resident_warps = (int)(((int64_t)per_sm * W * s->sm_count < s->P.B ? (int64_t)per_sm * W * s->sm_count : s->P.B));
    s->grid = (int)(need < cap ? need : cap);
    if (s->grid < 1) s->grid = 1;
    return HF_OK;
}

// Warps per CTA: the choice that keeps the most trajectories resident (a trajectory is one warp), so that a batch is
// one wave whenever it can be; ties go to the larger CTA.  With many charges the lists of one trajectory are ~15 KB of
// shared memory and the choice is worth up to 1.7x (2048 devices of 164 + 164 charges: 12 resident warps per SM with
// 4-warp CTAs = two rounds, 15 with 1-warp CTAs = one).
int pick_launch(hf_sim *s) {
    const size_t per_warp = hf_warp_smem_bytes(s->P.cap_c, s->P.cap_f, hf_xcap(s->P));
    const size_t limit = 227 * 1024;
    if (per_warp > limit)
        return fail(HF_ERR_UNSUPPORTED, "charges=%d needs %zu B of shared memory per trajectory (limit %zu)", s->P.C, per_warp, limit);
    int best_w = 0;
    int64_t best_resident = -1;
    int rc;
    if (per_warp * 8 <= limit) { if ((rc = configure_launch<8>(s))) return rc; best_w = 8; best_resident = (int64_t)s->resident_warps; }
    if (per_warp * 4 <= limit) { if ((rc = configure_launch<4>(s))) return rc; if ((int64_t)s->resident_warps > best_resident) { best_w = 4; best_resident = s->resident_warps; } }
    { if ((rc = configure_launch<1>(s))) return rc; if ((int64_t)s->resident_warps > best_resident) { best_w = 1; best_resident = s->resident_warps; } }
    if (best_w == 8) return configure_launch<8>(s);
    if (best_w == 4) return configure_launch<4>(s);
    return configure_launch<1>(s);
}

template <typename T>
int dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    HF_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    return HF_OK;
}

// (Re)configuration paths: keep a buffer that is already large enough.  cudaFree + cudaMalloc of the multi-GB occupancy
// masks and records of the dense devices cost 15-100 ms per call and dominated the end-to-end step.
template <typename T>
int dev_reuse(hf_sim *s, T **p, size_t n) {
    if (n == 0) n = 1;
    const size_t bytes = n * sizeof(T);
    if (*p) {
        auto it = s->alloc_bytes.find((void *)*p);
        if (it != s->alloc_bytes.end() && it->second >= bytes) return HF_OK;
        if (it != s->alloc_bytes.end()) s->alloc_bytes.erase(it);
        cudaFree(*p);
        *p = nullptr;
    }
    HF_CUDA(cudaMalloc((void **)p, bytes));
    s->alloc_bytes[(void *)*p] = bytes;
    return HF_OK;
}

void fill_params(hf_sim *s) {
    HfParams &P = s->P;
    const int64_t B = P.B;
    P.species = s->d_species; P.homo = s->d_homo; P.lumo = s->d_lumo;
    for (int t = 0; t < 2; ++t) {
        P.pos[t] = s->d_pos + (int64_t)t * B * P.cap_c;
        P.flag[t] = s->d_flag + (int64_t)t * B * P.cap_f;
        P.ev_init[t] = s->d_ev_init + (int64_t)t * B * P.cap_c;
        P.ev_final[t] = s->d_ev_final + (int64_t)t * B * P.cap_c;
        P.ev_tau[t] = s->d_ev_tau + (int64_t)t * B * P.cap_c;
    }
    P.sc = s->d_sc;
    P.cache_init = s->d_cache_i;
    P.cache_final = s->d_cache_i + B * HF_CACHE;
    P.cache_kp = s->d_cache_i + 2 * B * HF_CACHE;
    P.cache_tau = s->d_cache_tau;
    P.totals = s->d_totals;
}

int check_traj(hf_sim *s, int64_t t) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (t < 0 || t >= s->P.B) return fail(HF_ERR_BAD_ARG, "trajectory %lld out of range [0, %lld)", (long long)t, (long long)s->P.B);
    return HF_OK;
}

int32_t unpack_site(const hf_sim *s, int32_t p) {
    const int x = p & 1023, y = (p >> 10) & 1023, z = p >> 20;
    return (int32_t)(((int64_t)z * s->P.Y + y) * s->P.X + x);
}

int use_device(hf_sim *s) {
    HF_CUDA(cudaSetDevice(s->device));
    return HF_OK;
}

cudaEvent_t get_event(hf_sim *s) {
    if (!s->event_pool.empty()) { cudaEvent_t e = s->event_pool.back(); s->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void fold_timing(hf_sim *s) {
    for (auto &pr : s->timing) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) { s->timed_ms += ms; s->timed_launches += 1; }
        s->event_pool.push_back(pr.first);
        s->event_pool.push_back(pr.second);
    }
    s->timing.clear();
}

template <int W>
void launch_run(hf_sim *s) {
    if (s->P.xi.mode) hf_run_kernel<W, true><<<s->grid, W * 32, s->smem, s->stream>>>(s->P);
    else hf_run_kernel<W, false><<<s->grid, W * 32, s->smem, s->stream>>>(s->P);
}
template <int W>
void launch_inject(hf_sim *s, int32_t *pool) {
    hf_inject_kernel<W><<<s->grid, W * 32, s->smem, s->stream>>>(s->P, pool);
}

}  // namespace

// The exciton step kernel for a neighbour count M: register-resident tables when M / 32 is one of the
// instantiated slot counts (cutoff 3, 4, 5 lattice constants), the generic shared-memory variant otherwise.
typedef void (*hfx_kernel_fn)(const HfxParams);
#ifndef HF_X_WARPS
#define HF_X_WARPS 8
#endif
constexpr int kXWarps = HF_X_WARPS;
static hfx_kernel_fn x_pick_brick_kernel(int M) {
    switch (M) {
        case 122: return hfx_run_kernel<kXWarps, 3, 26, 2, true, 3, true>;
        case 256: return hfx_run_kernel<kXWarps, 8, 0, 2, true, 8, true>;
        case 514: return hfx_run_kernel<kXWarps, 16, 2, 2, false, 8, true>;
        default: return nullptr;                                                          // other cutoffs keep the row-major layout
    }
}
static hfx_kernel_fn x_pick_kernel(int M, bool brick = false) {
    if (brick) return x_pick_brick_kernel(M);
    switch (M) {                                                                          // the default cutoffs: M at compile time
        case 122: return hfx_run_kernel<kXWarps, 3, 26, 2, true>;                         // cutoff 3
        case 256: return hfx_run_kernel<kXWarps, 8, 0, 2, true>;                          // cutoff 4
        case 514: return hfx_run_kernel<kXWarps, 16, 2, 2, false, 8>;                     // cutoff 5: two halves of 8 slots, Dexter factors from shared memory
        default: break;
    }
    switch (M >> 5) {                                                                     // any other cutoff: full slots in registers, M at run time
        case 0: return hfx_run_kernel<kXWarps, 0, -1, 2, false>;
        case 1: return hfx_run_kernel<kXWarps, 1, -1, 2, true>;
        case 2: return hfx_run_kernel<kXWarps, 2, -1, 2, true>;
        case 3: return hfx_run_kernel<kXWarps, 3, -1, 2, true>;
        case 4: return hfx_run_kernel<kXWarps, 4, -1, 2, true>;
        case 5: return hfx_run_kernel<kXWarps, 5, -1, 2, true>;
        case 6: return hfx_run_kernel<kXWarps, 6, -1, 2, true>;
        case 7: return hfx_run_kernel<kXWarps, 7, -1, 2, true>;
        case 8: case 9: return hfx_run_kernel<kXWarps, 8, -1, 2, true>;
        case 10: case 11: return hfx_run_kernel<kXWarps, 10, -1, 2, false, 5>;
        case 12: case 13: return hfx_run_kernel<kXWarps, 12, -1, 2, false, 6>;
        case 14: case 15: return hfx_run_kernel<kXWarps, 14, -1, 2, false, 7>;
        default: return hfx_run_kernel<kXWarps, 16, -1, 2, false, 8>;                     // 16 or more slots: the first 16 prefetched
    }
}

// what hfx_run_kernel is launched with: the brick-layout copies when they exist
static HfxParams x_launch_params(const hf_sim *s) {
    HfxParams X = s->X;
    if (s->x_brick) {
        X.ws1 = s->d_bs1; X.wt1 = s->d_bt1; X.lin = s->d_x_blin; X.brick = 1; X.bX = s->x_bX; X.bY = s->x_bY;
        X.w_halo = (int64_t)X.R * X.bX * X.bY;
        X.w_stride = (int64_t)X.Z * X.bX * X.bY + 2 * X.w_halo;
    }
    return X;
}

void x_set_block(hf_sim *s, int block) {
    s->X.block = block;
    const int64_t blocks = (s->X.B + block - 1) / block;
    const int64_t need = (blocks + kXWarps - 1) / kXWarps, cap = (int64_t)s->sm_count * s->x_per_sm;
    s->x_grid = (int)(need < cap ? need : cap);
    if (s->x_grid < 1) s->x_grid = 1;
}

// Pick the warp batch by measurement: a few events per candidate, each from the same saved state, which is
// restored afterwards (tallies included).  Only when the weights exceed half the L2 and nothing is traced.
int x_autotune(hf_sim *s) {
    const size_t B = (size_t)s->P.B, nt = (size_t)s->n_real * HFX_T_N;
    int32_t *site = nullptr, *spin = nullptr;
    double *time = nullptr, *life = nullptr;
    uint64_t *draws = nullptr;
    unsigned long long *totals = nullptr;
    int rc;
    if ((rc = dev_alloc(&site, B)) || (rc = dev_alloc(&spin, B)) || (rc = dev_alloc(&time, B)) || (rc = dev_alloc(&life, B)) ||
        (rc = dev_alloc(&draws, B)) || (rc = dev_alloc(&totals, nt))) {
        cudaFree(site); cudaFree(spin); cudaFree(time); cudaFree(life); cudaFree(draws); cudaFree(totals);
        return rc;
    }
    auto copy = [&](bool save) {
        const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
        cudaMemcpyAsync(save ? site : s->d_x_site, save ? s->d_x_site : site, 4 * B, k, s->stream);
        cudaMemcpyAsync(save ? spin : s->d_x_spin, save ? s->d_x_spin : spin, 4 * B, k, s->stream);
        cudaMemcpyAsync(save ? time : s->d_x_time, save ? s->d_x_time : time, 8 * B, k, s->stream);
        cudaMemcpyAsync(save ? life : s->d_x_life, save ? s->d_x_life : life, 8 * B, k, s->stream);
        cudaMemcpyAsync(save ? draws : s->d_x_draws, save ? s->d_x_draws : draws, 8 * B, k, s->stream);
        cudaMemcpyAsync(save ? totals : s->d_x_totals, save ? s->d_x_totals : totals, 8 * nt, k, s->stream);
    };
    copy(true);
    HfxParams X = x_launch_params(s);
    X.trace = nullptr; X.trace_n = 0; X.trace_len = nullptr;
    X.n_events = 12;
    const int candidates[] = {32, 16, 8, 6, 4, 3, 2};
    cudaEvent_t e0 = get_event(s), e1 = get_event(s);
    int best = 32;
    float best_ms = 0.f;
    for (int c : candidates) {
        x_set_block(s, c);
        X.block = c;
        float ms = 0.f;
        for (int rep = 0; rep < 2; ++rep) {                                      // the first repetition warms the caches for this batch size
            cudaEventRecord(e0, s->stream);
            x_pick_kernel(X.M, s->x_brick)<<<s->x_grid, kXWarps * 32, s->x_smem, s->stream>>>(X);
            cudaEventRecord(e1, s->stream);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        copy(false);
        if (best_ms == 0.f || ms < best_ms) { best_ms = ms; best = c; }
    }
    s->event_pool.push_back(e0); s->event_pool.push_back(e1);
    x_set_block(s, best);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    cudaFree(site); cudaFree(spin); cudaFree(time); cudaFree(life); cudaFree(draws); cudaFree(totals);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return fail(HF_ERR_CUDA, "exciton batch tuning failed: %s", cudaGetErrorString(e));
    s->x_tune = false;
    return HF_OK;
}

int x_configure_launch(hf_sim *s) {
    const size_t smem = hfx_smem_bytes(s->X.M, kXWarps);
    if (smem > 200 * 1024) return fail(HF_ERR_UNSUPPORTED, "cutoff too large: %zu B of shared memory", smem);
    int l2_bytes = 0;
    cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, s->device);
    const double footprint = 8.0 * (double)s->n_real * (double)s->X.w_stride;                // one spin manifold
    const bool beyond_l2 = l2_bytes > 0 && footprint > 0.5 * (double)l2_bytes;
    // hfx_run_kernel gathers from a copy of the weights in 4 x 4 x 1 bricks (one 128-byte line each): a cutoff-4 sphere
    // touches 36 % fewer lines (hf_exciton.cuh).  The copy pads every plane with R periodic columns / rows on each side
    // (then up to multiples of 4), so no site needs the periodic wrap.  Measured on one B200, 10^6 trajectories
    // (profiles/r02g_brick_product.txt): 256^3 x 64 realisations 8.6e8 -> 1.06e9 events/s, 256^3 x 16 9.2e8 -> 1.19e9, the
    // L2-resident 128^3 1.89e9 -> 1.97e9.  Needs one of the compiled cutoffs and offsets that fit the packed table;
    // HF_X_BRICK=0 keeps the row-major layout.  The dense-device and integrated kernels always read the row-major arrays.
    s->x_brick = false;
    {
        bool want = true;
        if (const char *e = getenv("HF_X_BRICK")) want = atoi(e) != 0;
        const int R = s->X.R, bX = (s->P.X + 2 * R + 3) & ~3, bY = (s->P.Y + 2 * R + 3) & ~3;
        const bool can = x_pick_brick_kernel(s->X.M) != nullptr && (int64_t)(R + 1) * bX * bY < (1ll << 22);
        if (want && can) {
            const int64_t b_halo = (int64_t)R * bX * bY, b_stride = (int64_t)s->P.Z * bX * bY + 2 * b_halo;
            const size_t M = (size_t)s->X.M, RN = (size_t)s->n_real * (size_t)b_stride;
            std::vector<int32_t> blin(M);
            for (size_t m = 0; m < M; ++m) blin[m] = hfx_brick_lin(s->x_offsets[3 * m], s->x_offsets[3 * m + 1], s->x_offsets[3 * m + 2], bX, bY);
            int rc;
            if ((rc = dev_reuse(s, &s->d_x_blin, M)) || (rc = dev_reuse(s, &s->d_bs1, RN)) || (rc = dev_reuse(s, &s->d_bt1, RN))) return rc;
            HF_CUDA(cudaMemcpyAsync(s->d_x_blin, blin.data(), 4 * M, cudaMemcpyHostToDevice, s->stream));
            HF_CUDA(cudaMemsetAsync(s->d_bs1, 0, 8 * RN, s->stream));
            HF_CUDA(cudaMemsetAsync(s->d_bt1, 0, 8 * RN, s->stream));
            hfx_brick_weights_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>(s->d_species, s->d_s1, s->d_t1, s->d_bs1, s->d_bt1, s->P.X, s->P.Y,
                                                                             s->P.Z, R, bX, bY, (int64_t)s->n_real, b_stride, b_halo);
            HF_CUDA(cudaStreamSynchronize(s->stream));
            HF_CUDA(cudaGetLastError());
            s->x_brick = true; s->x_bX = bX; s->x_bY = bY;
        }
    }
    hfx_kernel_fn fn = x_pick_kernel(s->X.M, s->x_brick);
    if (smem > 48 * 1024) HF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kXWarps * 32, smem));
    if (per_sm < 1) return fail(HF_ERR_UNSUPPORTED, "exciton kernel does not fit an SM");
    // Trajectories a warp owns at a time.  32 amortises the per-event selection work best.  When the Boltzmann
    // weights do not fit the L2, consecutive events of one exciton (whose spheres overlap by ~3/4) only hit the L2
    // if the lines all resident trajectories touch in between fit it; fewer trajectories per warp make them fit
    // at the price of less amortisation.  Where the optimum lies depends on the lattice size, the number of
    // realisations and the exciton density (128^3 x 16 realisations: 8; 256^3 x 64: 3), so it
    // is measured: the first hf_x_run after hf_x_spawn times a few events per candidate on a copy of the state
    // (x_autotune).  Results do not depend on the choice.
    s->x_per_sm = per_sm; s->x_warps = kXWarps; s->x_smem = smem;
    s->x_tune = beyond_l2;
    x_set_block(s, 32);
    if (const char *e = getenv("HF_X_BLOCK")) { const int v = atoi(e); if (v >= 1 && v <= 32) { x_set_block(s, v); s->x_tune = false; } }
    return HF_OK;
}

// HfParams.xi.tab <- the tables, rate constants and weights of hf_x_configure (integrated mode 2)
static int sync_integrated(hf_sim *s) {
    HfiTables T;
    memset(&T, 0, sizeof(T));
    const HfxParams &X = s->X;
    T.M = X.M; T.X = s->P.X; T.Y = s->P.Y; T.Z = s->P.Z; T.k0 = s->P.k0; T.k1 = s->P.k1;
    T.offsets = X.offsets; T.g6 = X.g6; T.gd = X.gd;
    memcpy(T.k, X.kf, sizeof(X.kf)); memcpy(T.k + 9, X.kd, sizeof(X.kd));
    memcpy(T.k + 18, X.k_isc, sizeof(X.k_isc)); memcpy(T.k + 21, X.k_risc, sizeof(X.k_risc));
    memcpy(T.k + 24, X.k_rad, sizeof(X.k_rad)); memcpy(T.k + 30, X.k_nonrad, sizeof(X.k_nonrad));
    T.ws1 = X.ws1; T.wt1 = X.wt1; T.w_stride = X.w_stride; T.w_halo = X.w_halo;
    int rc;
    if (!s->d_pi_tab && (rc = dev_alloc(&s->d_pi_tab, 1))) return rc;
    HF_CUDA(cudaMemcpyAsync(s->d_pi_tab, &T, sizeof(T), cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    s->P.xi.tab = s->d_pi_tab;
    return HF_OK;
}

extern "C" {

const char *hf_last_error(void) { return g_error; }
int hf_abi_version(void) { return HF_ABI_VERSION; }

int hf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int hf_create(const hf_config *c, hf_sim **out) {
    if (!c || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    const int X = c->dims[0], Y = c->dims[1], Z = c->dims[2];
    if (Z < 3) return fail(HF_ERR_BAD_ARG, "Expected the last component of dimension to be > 2, got %d.", Z);   // reseau.py:131
    if (X < 2 || Y < 2) return fail(HF_ERR_BAD_ARG, "x and y dimensions must be >= 2, got (%d, %d)", X, Y);
    if (X > kMaxDim || Y > kMaxDim || Z > kMaxDim) return fail(HF_ERR_UNSUPPORTED, "dimensions above %d are not supported", kMaxDim);
    const int64_t N = (int64_t)X * Y * Z;
    if (N > 0x7fffffffLL) return fail(HF_ERR_UNSUPPORTED, "more than 2^31-1 sites");
    double pmin = c->props[0] < c->props[1] ? c->props[0] : c->props[1];
    if (c->props[2] < pmin) pmin = c->props[2];
    if (pmin * (double)N < 1)                                                                    // reseau.py:137-139
        return fail(HF_ERR_BAD_ARG, "Not enough molecules for current proportions and dimension.");
    if (c->charges < 1 || c->charges > (int64_t)X * Y)                                           // reseau.py:212-213
        return fail(HF_ERR_BAD_ARG, "Required %d charges but only %lld molecules available.", c->charges, (long long)X * Y);
    // reseau.py:190-205 for any distance (quirk Q3 reproduced verbatim).  The table indexes [0 .. distance] on every
    // axis, so a lattice thinner than that is an IndexError in the reference.
    if (c->distance < 1 || c->distance > 8)
        return fail(HF_ERR_UNSUPPORTED, "charge_tranfer_distance=%d: supported range is 1..8", c->distance);
    if (X < c->distance + 1 || Y < c->distance + 1 || Z < c->distance + 1)
        return fail(HF_ERR_BAD_ARG, "charge_tranfer_distance=%d needs every dimension >= %d (list index out of range in the reference)",
                    c->distance, c->distance + 1);
    if (c->n_trajectories < 1 || c->n_realisations < 1 || c->n_trajectories % c->n_realisations != 0)
        return fail(HF_ERR_BAD_ARG, "n_realisations (%d) must divide n_trajectories (%lld)", c->n_realisations, (long long)c->n_trajectories);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(HF_ERR_NO_DEVICE, "no CUDA device visible: libhfkmc has no CPU path");
    }
    if (c->device < 0 || c->device >= ndev) return fail(HF_ERR_BAD_ARG, "device %d out of range (%d visible)", c->device, ndev);
    cudaDeviceProp prop;
    HF_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10)
        return fail(HF_ERR_NO_DEVICE, "device %d is sm_%d%d; libhfkmc is built for sm_100a only", c->device, prop.major, prop.minor);

    hf_sim *s = new (std::nothrow) hf_sim();
    if (!s) return fail(HF_ERR_BAD_ARG, "out of host memory");
    s->cfg = *c;
    s->device = c->device;
    s->n_real = c->n_realisations;
    s->sm_count = prop.multiProcessorCount;
    HfParams &P = s->P;
    memset(&P, 0, sizeof(P));
    P.X = X; P.Y = Y; P.Z = Z; P.N = N;
    P.field = c->electric_field;
    P.C = c->charges;
    P.cap_c = c->charges;
    P.dist = c->distance;
    // distance > 1 always runs the warp kernel, whose flag sets need their general capacity
    P.cap_f = (c->charges == 1 && c->distance == 1) ? 2 : 2 * c->charges + 8;
    P.B = c->n_trajectories;
    P.traj_per_real = c->n_trajectories / c->n_realisations;
    P.k0 = (uint32_t)c->seed; P.k1 = (uint32_t)(c->seed >> 32);
    hf_philox_round_keys(P.k0, P.k1, P.rk);
    P.first_traj = c->first_trajectory; P.first_real = c->first_realisation;
    {   // random.sample's setsize for k = charges (random.py:421-424)
        int64_t setsize = 21;
        const int64_t k = c->charges;
        if (k > 5) {
            const int64_t e = (int64_t)std::ceil(std::log((double)(k * 3)) / std::log(4.0));
            int64_t p4 = 1;
            for (int64_t i = 0; i < e; ++i) p4 *= 4;
            setsize += p4;
        }
        P.setsize = setsize;
    }
    int rc = HF_OK;
    auto bail = [&](int code) { hf_destroy(s); return code; };
    if (cudaSetDevice(s->device) != cudaSuccess) return bail(fail(HF_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(HF_ERR_CUDA, "cudaStreamCreate failed"));
    s->own_stream = true;
    const int64_t B = P.B, R = s->n_real;
    if ((rc = dev_alloc(&s->d_species, (size_t)(R * N)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_homo, (size_t)(R * N)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_lumo, (size_t)(R * N)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_s1, (size_t)(R * N)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_t1, (size_t)(R * N)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_pos, (size_t)(2 * B * P.cap_c)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_flag, (size_t)(2 * B * P.cap_f)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_ev_init, (size_t)(2 * B * P.cap_c)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_ev_final, (size_t)(2 * B * P.cap_c)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_ev_tau, (size_t)(2 * B * P.cap_c)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_sc, (size_t)B))) return bail(rc);
    if ((rc = dev_alloc(&s->d_cache_i, (size_t)(3 * B * HF_CACHE)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_cache_tau, (size_t)(B * HF_CACHE)))) return bail(rc);
    if ((rc = dev_alloc(&s->d_totals, (size_t)HF_N_TALLIES))) return bail(rc);
    if ((rc = dev_alloc(&s->d_sum, (size_t)HF_N_TALLIES))) return bail(rc);
    // no array is ever read uninitialised (a lattice uploaded without S1 / T1 reads zeros through the facade)
    cudaMemsetAsync(s->d_species, 0, (size_t)(R * N), s->stream);
    for (double *p : {s->d_homo, s->d_lumo, s->d_s1, s->d_t1}) cudaMemsetAsync(p, 0, 8 * (size_t)(R * N), s->stream);
    s->have_st.assign((size_t)R, 0);
    if (c->charges > HF_SMALL_MAX_CHARGES || (c->flags & HF_FLAG_WARP_KERNEL) || c->distance != 1) {
        const int m = (X > Y ? (X > Z ? X : Z) : (Y > Z ? Y : Z)) - 1;
        const size_t n = 3 * (size_t)m * (size_t)m + 1;
        if (n * 8 <= ((size_t)64 << 20)) {                                        // <= 64 MB; larger lattices compute on the fly
            std::vector<double> tab(n);
            tab[0] = 0.0;
            for (size_t d2 = 1; d2 < n; ++d2) tab[d2] = 1.0 / std::sqrt((double)d2);   // IEEE: the bits of __ddiv_rn(1, __dsqrt_rn(d2))
            if ((rc = dev_alloc(&s->d_inv_dist, n))) return bail(rc);
            if (cudaMemcpyAsync(s->d_inv_dist, tab.data(), n * 8, cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(HF_ERR_CUDA, "inverse-distance table upload failed"));
            P.inv_dist = s->d_inv_dist;
        }
    }
    fill_params(s);
    if ((rc = pick_launch(s))) return bail(rc);
    *out = s;
    return HF_OK;
}

int hf_destroy(hf_sim *s) {
    if (!s) return HF_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    fold_timing(s);
    for (cudaEvent_t e : s->event_pool) cudaEventDestroy(e);
    cudaFree(s->d_species); cudaFree(s->d_homo); cudaFree(s->d_lumo); cudaFree(s->d_s1); cudaFree(s->d_t1);
    cudaFree(s->d_pos); cudaFree(s->d_flag); cudaFree(s->d_ev_init); cudaFree(s->d_ev_final); cudaFree(s->d_ev_tau);
    cudaFree(s->d_sc); cudaFree(s->d_cache_i); cudaFree(s->d_cache_tau); cudaFree(s->d_totals); cudaFree(s->d_sum);
    cudaFree(s->d_replay); cudaFree(s->d_trace); cudaFree(s->d_trace_base);
    cudaFree(s->d_x_off); cudaFree(s->d_x_lin); cudaFree(s->d_x_site); cudaFree(s->d_x_spin); cudaFree(s->d_x_g6); cudaFree(s->d_x_gd);
    cudaFree(s->d_ws1); cudaFree(s->d_wt1); cudaFree(s->d_bs1); cudaFree(s->d_bt1); cudaFree(s->d_x_blin); cudaFree(s->d_x_time); cudaFree(s->d_x_life); cudaFree(s->d_x_draws);
    cudaFree(s->d_x_trace_len); cudaFree(s->d_x_totals);
    cudaFree(s->d_xd_pos); cudaFree(s->d_xd_occ); cudaFree(s->d_xd_T); cudaFree(s->d_xd_birth); cudaFree(s->d_xd_rec_L); cudaFree(s->d_xd_rec_seen); cudaFree(s->d_xd_time);
    cudaFree(s->d_xd_life); cudaFree(s->d_xd_draws); cudaFree(s->d_xd_totals);
    cudaFree(s->d_gen_layer); cudaFree(s->d_gen_pool); cudaFree(s->d_gen_counts); cudaFree(s->d_real_out);
    cudaFree(s->d_pi_site); cudaFree(s->d_pi_meta); cudaFree(s->d_pi_final); cudaFree(s->d_pi_n);
    cudaFree(s->d_pi_tau); cudaFree(s->d_pi_clock); cudaFree(s->d_pi_draws); cudaFree(s->d_pi_tab); cudaFree(s->d_inv_dist);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return HF_OK;
}

int hf_set_stream(hf_sim *s, void *cuda_stream) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    if (s->own_stream) { cudaStreamDestroy(s->stream); s->own_stream = false; }
    s->stream = (cudaStream_t)cuda_stream;
    return HF_OK;
}

int hf_sync(hf_sim *s) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    HF_CUDA(cudaGetLastError());
    return HF_OK;
}

static int generate_lattice_impl(hf_sim *s, const double *props /* [n_real][3] */) {
    int rc = use_device(s);
    if (rc) return rc;
    const HfParams &P = s->P;
    const int64_t xy = (int64_t)P.X * P.Y, grid_size = P.N;
    std::vector<int64_t> counts((size_t)(2 * s->n_real));
    int64_t total = 0;
    for (int r = 0; r < s->n_real; ++r) {
        const double *pr = props + 3 * r;
        const int64_t n_fluo = (int64_t)((double)grid_size * pr[2]);              // reseau.py:159
        const int64_t n_tadf = (int64_t)((double)grid_size * pr[1]);              // reseau.py:160
        const int64_t n_host = grid_size - 2 * xy - n_fluo - n_tadf;              // reseau.py:161
        total = n_host + n_tadf + n_fluo;
        if (total != xy * (P.Z - 2) || total <= 0 || n_fluo < 0 || n_tadf < 0)    // reseau.py:169
            return fail(HF_ERR_BAD_ARG, "realisation %d: Size of sub_grid (%lld) and x_max*y_max*sub_z_max (%lld) must match !",
                        r, (long long)total, (long long)(xy * (P.Z - 2)));
        counts[(size_t)r] = n_host;
        counts[(size_t)(s->n_real + r)] = n_tadf;
    }
    if ((rc = dev_reuse(s, &s->d_gen_layer, (size_t)(xy * s->n_real)))) return rc;
    if ((rc = dev_reuse(s, &s->d_gen_counts, counts.size()))) return rc;
    // realisations are shuffled in batches so that the pool scratch stays below ~1 GiB
    int batch = (int)((int64_t)(1 << 28) / total);
    if (batch < 1) batch = 1;
    if (batch > s->n_real) batch = s->n_real;
    if ((rc = dev_reuse(s, &s->d_gen_pool, (size_t)(total * batch)))) return rc;
    uint8_t *d_layer = s->d_gen_layer;
    int32_t *d_pool = s->d_gen_pool;
    int64_t *d_counts = s->d_gen_counts;
    cudaMemcpyAsync(d_counts, counts.data(), counts.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s->stream);
    for (int r0 = 0; r0 < s->n_real; r0 += batch) {
        const int nb = s->n_real - r0 < batch ? s->n_real - r0 : batch;
        hf_species_pool_init<<<s->sm_count * 8, 256, 0, s->stream>>>(d_pool, total, nb);
        hf_species_shuffle<<<nb, 32, 0, s->stream>>>(P, d_pool, total, d_counts, d_counts + s->n_real, d_layer, r0);
    }
    hf_fill_sites<<<s->sm_count * 8, 256, 0, s->stream>>>(P, d_layer, s->n_real, s->d_species, s->d_homo, s->d_lumo, s->d_s1, s->d_t1);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(HF_ERR_CUDA, "lattice generation failed: %s", cudaGetErrorString(e));
    HF_CUDA(cudaGetLastError());
    s->have_lattice = true;
    std::fill(s->have_st.begin(), s->have_st.end(), 1);
    return HF_OK;
}

int hf_generate_lattice(hf_sim *s) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    std::vector<double> props((size_t)(3 * s->n_real));
    for (int r = 0; r < s->n_real; ++r)
        for (int k = 0; k < 3; ++k) props[(size_t)(3 * r + k)] = s->cfg.props[k];
    return generate_lattice_impl(s, props.data());
}

int hf_generate_lattice_mixed(hf_sim *s, const double *props) {
    if (!s || !props) return fail(HF_ERR_BAD_ARG, "null argument");
    return generate_lattice_impl(s, props);
}

int hf_read_realisation_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    int rc = use_device(s);
    if (rc) return rc;
    if ((rc = dev_reuse(s, &s->d_real_out, (size_t)(5 * s->n_real)))) return rc;
    unsigned long long *d_out = s->d_real_out;
    hf_reduce_realisations_kernel<<<s->n_real, 256, 0, s->stream>>>(s->d_sc, s->P.traj_per_real, d_out);
    cudaMemcpyAsync(out, d_out, sizeof(uint64_t) * 5 * (size_t)s->n_real, cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return fail(HF_ERR_CUDA, "hf_read_realisation_tallies failed: %s", cudaGetErrorString(e));
    return HF_OK;
}

int hf_read_infos(hf_sim *s, int64_t first, int64_t n, hf_traj_info *out) {
    int rc = check_traj(s, first);
    if (rc) return rc;
    if (n < 1 || first + n > s->P.B || !out) return fail(HF_ERR_BAD_ARG, "bad range");
    if ((rc = use_device(s))) return rc;
    std::vector<HfScalars> tmp((size_t)n);
    HF_CUDA(cudaMemcpyAsync(tmp.data(), s->d_sc + first, sizeof(HfScalars) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < n; ++i) {
        const HfScalars &sc = tmp[(size_t)i];
        hf_traj_info &o = out[i];
        o.n_electrons = sc.n_pos[0]; o.n_holes = sc.n_pos[1]; o.n_eflags = sc.n_flag[0]; o.n_hflags = sc.n_flag[1];
        o.n_move_e = sc.n_ev[0]; o.n_move_h = sc.n_ev[1];
        o.special_kind = sc.special & 0xff; o.special_particule = sc.special >> 8;
        o.special_site = sc.special ? unpack_site(s, sc.special_site) : -1;
        o.exciton_state = sc.xstate; o.status = sc.status; o.cache_count = sc.cache_n;
        o.draws = sc.draws; o.spins = sc.spins; o.emission = sc.emission; o.recombination = sc.recombination;
        o.injection = sc.injection; o.steps = sc.steps; o.fired = sc.fired;
    }
    return HF_OK;
}

int hf_upload_lattice(hf_sim *s, int32_t r, const uint8_t *species, const double *homo, const double *lumo,
                      const double *s1, const double *t1) {
    if (!s || !species || !homo || !lumo) return fail(HF_ERR_BAD_ARG, "null argument");
    if (r < 0 || r >= s->n_real) return fail(HF_ERR_BAD_ARG, "realisation %d out of range", r);
    int rc = use_device(s);
    if (rc) return rc;
    const size_t N = (size_t)s->P.N, off = (size_t)r * N;
    HF_CUDA(cudaMemcpyAsync(s->d_species + off, species, N, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaMemcpyAsync(s->d_homo + off, homo, N * 8, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaMemcpyAsync(s->d_lumo + off, lumo, N * 8, cudaMemcpyHostToDevice, s->stream));
    if (s1) HF_CUDA(cudaMemcpyAsync(s->d_s1 + off, s1, N * 8, cudaMemcpyHostToDevice, s->stream));
    if (t1) HF_CUDA(cudaMemcpyAsync(s->d_t1 + off, t1, N * 8, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));      // the caller may free its buffers on return
    s->have_st[(size_t)r] = (s1 && t1) ? 1 : 0;
    s->have_lattice = true;
    return HF_OK;
}

int hf_download_lattice(hf_sim *s, int32_t r, uint8_t *species, double *homo, double *lumo, double *s1, double *t1) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (r < 0 || r >= s->n_real) return fail(HF_ERR_BAD_ARG, "realisation %d out of range", r);
    if (!s->have_lattice) return fail(HF_ERR_STATE, "no lattice yet");
    int rc = use_device(s);
    if (rc) return rc;
    const size_t N = (size_t)s->P.N, off = (size_t)r * N;
    if (species) HF_CUDA(cudaMemcpyAsync(species, s->d_species + off, N, cudaMemcpyDeviceToHost, s->stream));
    if (homo) HF_CUDA(cudaMemcpyAsync(homo, s->d_homo + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    if (lumo) HF_CUDA(cudaMemcpyAsync(lumo, s->d_lumo + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    if (s1) HF_CUDA(cudaMemcpyAsync(s1, s->d_s1 + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    if (t1) HF_CUDA(cudaMemcpyAsync(t1, s->d_t1 + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    return HF_OK;
}

int hf_set_replay(hf_sim *s, const double *u, int64_t n) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_replay);
    s->d_replay = nullptr;
    s->P.replay = nullptr;
    s->P.n_replay = 0;
    if (u) {
        if (n < 1) return fail(HF_ERR_BAD_ARG, "replay length must be positive");
        const size_t total = (size_t)n * (size_t)s->P.B;
        if ((rc = dev_alloc(&s->d_replay, total))) return rc;
        HF_CUDA(cudaMemcpyAsync(s->d_replay, u, total * 8, cudaMemcpyHostToDevice, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
        s->P.replay = s->d_replay;
        s->P.n_replay = n;
    }
    return HF_OK;
}

int hf_inject(hf_sim *s) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->have_lattice) return fail(HF_ERR_STATE, "hf_inject before a lattice exists (hf_generate_lattice / hf_upload_lattice)");
    int rc = use_device(s);
    if (rc) return rc;
    const HfParams &P = s->P;
    HF_CUDA(cudaMemsetAsync(s->d_sc, 0, sizeof(HfScalars) * (size_t)P.B, s->stream));
    HF_CUDA(cudaMemsetAsync(s->d_cache_i, 0, sizeof(int32_t) * 3 * (size_t)P.B * HF_CACHE, s->stream));
    HF_CUDA(cudaMemsetAsync(s->d_cache_tau, 0, sizeof(double) * (size_t)P.B * HF_CACHE, s->stream));
    HF_CUDA(cudaMemsetAsync(s->d_totals, 0, sizeof(unsigned long long) * HF_N_TALLIES, s->stream));
    if (s->P.xi.mode) {
        HF_CUDA(cudaMemsetAsync(s->d_pi_n, 0, sizeof(int32_t) * (size_t)P.B, s->stream));
        HF_CUDA(cudaMemsetAsync(s->d_pi_draws, 0, sizeof(uint64_t) * (size_t)P.B, s->stream));
        HF_CUDA(cudaMemsetAsync(s->d_pi_clock, 0, sizeof(double) * (size_t)P.B, s->stream));
    }
    if (s->d_trace_base) {                                                   // `fired` restarts at 0, so does the trace
        HF_CUDA(cudaMemsetAsync(s->d_trace_base, 0, sizeof(uint64_t) * s->trace_base.size(), s->stream));
        std::fill(s->trace_base.begin(), s->trace_base.end(), 0);
    }
    if (P.C == 1 && P.dist == 1 && P.xi.mode == 0 && P.cap_f == 2 && !(s->cfg.flags & HF_FLAG_WARP_KERNEL)) {
        // one thread per trajectory, like the step kernel (hf_frm_c1.cuh)
        const int64_t blocks = (P.B + HF_C1_THREADS - 1) / HF_C1_THREADS;
        hf_inject_c1_kernel<<<(unsigned)blocks, HF_C1_THREADS, 0, s->stream>>>(P);
        HF_CUDA(cudaGetLastError());
        s->injected = true;
        return HF_OK;
    }
    const int64_t molecules = (int64_t)P.X * P.Y;
    if (P.C <= HF_SMALL_MAX_CHARGES && P.dist == 1 && !(s->cfg.flags & HF_FLAG_WARP_KERNEL) && molecules > P.setsize && !getenv("HF_INJECT_WARP")) {
        // one thread per trajectory (hf_frm_small.cuh), set variant of random.sample
        const size_t smem = hf_small_thread_bytes(P.cap_c, P.cap_f, hf_xcap(P)) * HF_SMALL_THREADS;
        if (smem > 48 * 1024) HF_CUDA(cudaFuncSetAttribute(hf_inject_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t blocks = (P.B + HF_SMALL_THREADS - 1) / HF_SMALL_THREADS;
        hf_inject_small_kernel<<<(unsigned)blocks, HF_SMALL_THREADS, smem, s->stream>>>(P);
        HF_CUDA(cudaGetLastError());
        s->injected = true;
        return HF_OK;
    }
    int32_t *pool = nullptr;
    if (molecules <= P.setsize)
        if ((rc = dev_alloc(&pool, (size_t)((int64_t)s->grid * s->warps * molecules)))) return rc;
    switch (s->warps) {
        case 8: launch_inject<8>(s, pool); break;
        case 4: launch_inject<4>(s, pool); break;
        default: launch_inject<1>(s, pool); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && pool) e = cudaStreamSynchronize(s->stream);
    cudaFree(pool);
    if (e != cudaSuccess) return fail(HF_ERR_CUDA, "inject kernel failed: %s", cudaGetErrorString(e));
    s->injected = true;
    return HF_OK;
}

int hf_run(hf_sim *s, int64_t stop) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->injected) return fail(HF_ERR_STATE, "hf_run before hf_inject");
    if (stop < 0) return fail(HF_ERR_BAD_ARG, "stop must be >= 0");
    int rc = use_device(s);
    if (rc) return rc;
    s->P.n_steps = stop;
    if (s->timing.size() >= 1024) { HF_CUDA(cudaStreamSynchronize(s->stream)); fold_timing(s); }
    cudaEvent_t e0 = get_event(s), e1 = get_event(s);
    HF_CUDA(cudaEventRecord(e0, s->stream));
    const bool thread_kernels = !(s->cfg.flags & HF_FLAG_WARP_KERNEL) && s->P.dist == 1;   // both are built on the 3x3x3 table
    if (s->P.C == 1 && thread_kernels && s->P.xi.mode == 0) {
        // one thread per trajectory (hf_frm_c1.cuh)
        // waves x cost per wave of the two shapes (hf_frm_c1.cuh): the smaller-register shape only when it saves a wave
        // time of one full wave of shape B over one of shape A, measured on the final kernel: 127 872 trajectories x 1000
        // events in 18.65 ms against 113 664 in 15.50 ms (B is 12.5 % more trajectories and 6.5 % slower per event)
        constexpr double kWaveCostB = 1.20;
        const int64_t wave_a = (int64_t)s->sm_count * HF_C1_THREADS * HF_C1_MIN_BLOCKS;
        const int64_t wave_b = (int64_t)s->sm_count * HF_C1_THREADS_B * HF_C1_MIN_BLOCKS_B;
        const double cost_a = (double)((s->P.B + wave_a - 1) / wave_a), cost_b = kWaveCostB * (double)((s->P.B + wave_b - 1) / wave_b);
        // a partly filled last wave runs faster than a full one: credit it half
        const double frac_a = (double)(s->P.B % wave_a) / (double)wave_a, frac_b = (double)(s->P.B % wave_b) / (double)wave_b;
        const double est_a = cost_a - (frac_a > 0 ? 0.5 * (1.0 - frac_a) : 0.0), est_b = cost_b - (frac_b > 0 ? kWaveCostB * 0.5 * (1.0 - frac_b) : 0.0);
        if (est_b < est_a && !getenv("HF_C1_SHAPE_A")) {
            const int64_t blocks = (s->P.B + HF_C1_THREADS_B - 1) / HF_C1_THREADS_B;
            (s->P.replay ? hf_run_c1_kernel<HF_C1_THREADS_B, HF_C1_MIN_BLOCKS_B, true> : hf_run_c1_kernel<HF_C1_THREADS_B, HF_C1_MIN_BLOCKS_B, false>)<<<(unsigned)blocks, HF_C1_THREADS_B, 0, s->stream>>>(s->P);
        } else {
            const int64_t blocks = (s->P.B + HF_C1_THREADS - 1) / HF_C1_THREADS;
            (s->P.replay ? hf_run_c1_kernel<HF_C1_THREADS, HF_C1_MIN_BLOCKS, true> : hf_run_c1_kernel<HF_C1_THREADS, HF_C1_MIN_BLOCKS, false>)<<<(unsigned)blocks, HF_C1_THREADS, 0, s->stream>>>(s->P);
        }
    } else if (s->P.C <= HF_SMALL_MAX_CHARGES && thread_kernels) {
        // one thread per trajectory, lists in per-thread shared memory (hf_frm_small.cuh)
        const size_t smem = hf_small_thread_bytes(s->P.cap_c, s->P.cap_f, hf_xcap(s->P)) * HF_SMALL_THREADS;
        const int64_t blocks = (s->P.B + HF_SMALL_THREADS - 1) / HF_SMALL_THREADS;
        if (s->P.xi.mode) {
            if (smem > 48 * 1024) HF_CUDA(cudaFuncSetAttribute(hf_run_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hf_run_small_kernel<true><<<(unsigned)blocks, HF_SMALL_THREADS, smem, s->stream>>>(s->P);
        } else {
            if (smem > 48 * 1024) HF_CUDA(cudaFuncSetAttribute(hf_run_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hf_run_small_kernel<false><<<(unsigned)blocks, HF_SMALL_THREADS, smem, s->stream>>>(s->P);
        }
    } else {
        switch (s->warps) {
            case 8: launch_run<8>(s); break;
            case 4: launch_run<4>(s); break;
            default: launch_run<1>(s); break;
        }
    }
    HF_CUDA(cudaEventRecord(e1, s->stream));
    s->timing.emplace_back(e0, e1);
    HF_CUDA(cudaGetLastError());
    return HF_OK;
}

int hf_run_timing(hf_sim *s, double *total_ms, int64_t *launches) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    fold_timing(s);
    if (total_ms) *total_ms = s->timed_ms;
    if (launches) *launches = s->timed_launches;
    s->timed_ms = 0.0;
    s->timed_launches = 0;
    return HF_OK;
}

int hf_reduce_tallies_device(hf_sim *s, void *out_device) {
    if (!s || !out_device) return fail(HF_ERR_BAD_ARG, "null argument");
    int rc = use_device(s);
    if (rc) return rc;
    unsigned long long *out = (unsigned long long *)out_device;
    // channel counters (5..15) are accumulated by the step kernel; the rest is summed here
    HF_CUDA(cudaMemcpyAsync(out, s->d_totals, sizeof(unsigned long long) * HF_N_TALLIES, cudaMemcpyDeviceToDevice, s->stream));
    int blocks = (int)((s->P.B + 255) / 256);
    if (blocks > s->sm_count * 8) blocks = s->sm_count * 8;
    hf_reduce_kernel<<<blocks, 256, 0, s->stream>>>(s->d_sc, s->P.B, out);
    HF_CUDA(cudaGetLastError());
    return HF_OK;
}

int hf_read_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    int rc = hf_reduce_tallies_device(s, s->d_sum);
    if (rc) return rc;
    HF_CUDA(cudaMemcpyAsync(out, s->d_sum, sizeof(uint64_t) * HF_N_TALLIES, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    return HF_OK;
}

int hf_read_info(hf_sim *s, int64_t t, hf_traj_info *out) {
    int rc = check_traj(s, t);
    if (rc) return rc;
    if (!out) return fail(HF_ERR_BAD_ARG, "null argument");
    if ((rc = use_device(s))) return rc;
    HfScalars sc;
    HF_CUDA(cudaMemcpyAsync(&sc, s->d_sc + t, sizeof(sc), cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    out->n_electrons = sc.n_pos[0]; out->n_holes = sc.n_pos[1];
    out->n_eflags = sc.n_flag[0]; out->n_hflags = sc.n_flag[1];
    out->n_move_e = sc.n_ev[0]; out->n_move_h = sc.n_ev[1];
    out->special_kind = sc.special & 0xff; out->special_particule = sc.special >> 8;
    out->special_site = sc.special ? unpack_site(s, sc.special_site) : -1;
    out->exciton_state = sc.xstate; out->status = sc.status; out->cache_count = sc.cache_n;
    out->draws = sc.draws; out->spins = sc.spins;
    out->emission = sc.emission; out->recombination = sc.recombination; out->injection = sc.injection;
    out->steps = sc.steps; out->fired = sc.fired;
    return HF_OK;
}

static int read_list(hf_sim *s, int64_t t, const int32_t *base, int cap, int count, int32_t *sites, int32_t *n) {
    if (!sites || !n) return fail(HF_ERR_BAD_ARG, "null argument");
    if (*n < count) return fail(HF_ERR_BAD_ARG, "buffer too small: need %d, have %d", count, *n);
    std::vector<int32_t> tmp((size_t)(count > 0 ? count : 1));
    if (count > 0) {
        HF_CUDA(cudaMemcpyAsync(tmp.data(), base + t * cap, sizeof(int32_t) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
    }
    for (int i = 0; i < count; ++i) sites[i] = unpack_site(s, tmp[(size_t)i]);
    *n = count;
    return HF_OK;
}

int hf_read_positions(hf_sim *s, int64_t t, int32_t which, int32_t *sites, int32_t *n) {
    hf_traj_info info;
    int rc = hf_read_info(s, t, &info);
    if (rc) return rc;
    if (which == 0) return read_list(s, t, s->P.pos[0], s->P.cap_c, info.n_electrons, sites, n);
    if (which == 1) return read_list(s, t, s->P.pos[1], s->P.cap_c, info.n_holes, sites, n);
    if (which == 2 && s->P.xi.mode) {   // integrated: the exciton slots, in creation order
        int32_t cnt = 0;
        HF_CUDA(cudaMemcpyAsync(&cnt, s->d_pi_n + t, 4, cudaMemcpyDeviceToHost, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
        return read_list(s, t, s->d_pi_site, s->P.cap_c, cnt, sites, n);
    }
    if (which == 2) {   // _excitons_locations: the exciton lives between its bound and decay steps
        if (!sites || !n) return fail(HF_ERR_BAD_ARG, "null argument");
        if (info.special_kind == HF_EV_DECAY) {
            if (*n < 1) return fail(HF_ERR_BAD_ARG, "buffer too small");
            sites[0] = info.special_site;
            *n = 1;
        } else {
            *n = 0;
        }
        return HF_OK;
    }
    return fail(HF_ERR_BAD_ARG, "which must be 0, 1 or 2");
}

int hf_p_excitons(hf_sim *s, int32_t mode) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (mode < 0 || mode > 2) return fail(HF_ERR_BAD_ARG, "mode must be 0 (off), 1 (decay at once) or 2 (dynamics)");
    if (mode == 2 && !s->x_configured) return fail(HF_ERR_STATE, "hf_p_excitons(2) before hf_x_configure: the dynamics use its tables and weights");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    const size_t B = (size_t)s->P.B, n = B * (size_t)s->P.cap_c;
    if (mode && !s->d_pi_site) {
        if ((rc = dev_alloc(&s->d_pi_site, n)) || (rc = dev_alloc(&s->d_pi_meta, n)) || (rc = dev_alloc(&s->d_pi_final, n)) ||
            (rc = dev_alloc(&s->d_pi_tau, n)) || (rc = dev_alloc(&s->d_pi_n, B)) || (rc = dev_alloc(&s->d_pi_draws, B)) ||
            (rc = dev_alloc(&s->d_pi_clock, B))) return rc;
    }
    HfIntegrated &I = s->P.xi;
    memset(&I, 0, sizeof(I));
    I.mode = mode;
    I.x_site = s->d_pi_site; I.x_meta = s->d_pi_meta; I.x_final = s->d_pi_final; I.x_tau = s->d_pi_tau;
    I.x_n = s->d_pi_n; I.x_draws = s->d_pi_draws; I.x_clock = s->d_pi_clock;
    if (mode == 2 && (rc = sync_integrated(s))) return rc;
    s->injected = false;                               // the event lists change shape: hf_inject starts the trajectories
    return pick_launch(s);
}

int hf_read_excitons(hf_sim *s, int64_t t, int32_t *site, int32_t *spin, int32_t *kind, int32_t *final_, double *tau,
                     int32_t *radiative, int32_t *n, uint64_t *draws, double *clock) {
    int rc = check_traj(s, t);
    if (rc) return rc;
    if (!n) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->P.xi.mode) { *n = 0; if (draws) *draws = 0; if (clock) *clock = 0.0; return HF_OK; }
    if ((rc = use_device(s))) return rc;
    int32_t cnt = 0;
    uint64_t d = 0;
    double c = 0.0;
    HF_CUDA(cudaMemcpyAsync(&cnt, s->d_pi_n + t, 4, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(&d, s->d_pi_draws + t, 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(&c, s->d_pi_clock + t, 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    if (draws) *draws = d;
    if (clock) *clock = c;
    if (*n < cnt) return fail(HF_ERR_BAD_ARG, "buffer too small: need %d", cnt);
    *n = cnt;
    if (cnt == 0) return HF_OK;
    std::vector<int32_t> a((size_t)cnt), b((size_t)cnt), f((size_t)cnt);
    const size_t off = (size_t)t * (size_t)s->P.cap_c;
    HF_CUDA(cudaMemcpyAsync(a.data(), s->d_pi_site + off, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(b.data(), s->d_pi_meta + off, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(f.data(), s->d_pi_final + off, 4 * (size_t)cnt, cudaMemcpyDeviceToHost, s->stream));
    if (tau) HF_CUDA(cudaMemcpyAsync(tau, s->d_pi_tau + off, 8 * (size_t)cnt, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < cnt; ++i) {
        if (site) site[i] = unpack_site(s, a[(size_t)i]);
        if (final_) final_[i] = unpack_site(s, f[(size_t)i]);
        if (spin) spin[i] = HFI_SPIN(b[(size_t)i]);
        if (kind) kind[i] = HFI_KIND(b[(size_t)i]);
        if (radiative) radiative[i] = (b[(size_t)i] & HFI_RAD) ? 1 : 0;
    }
    return HF_OK;
}

int hf_read_flags(hf_sim *s, int64_t t, int32_t which, int32_t *sites, int32_t *n) {
    hf_traj_info info;
    int rc = hf_read_info(s, t, &info);
    if (rc) return rc;
    if (which != 0 && which != 1) return fail(HF_ERR_BAD_ARG, "which must be 0 or 1");
    return read_list(s, t, s->P.flag[which], s->P.cap_f, which ? info.n_hflags : info.n_eflags, sites, n);
}

int hf_read_move_events(hf_sim *s, int64_t t, int32_t which, int32_t *initial, int32_t *final_, double *tau, int32_t *n) {
    hf_traj_info info;
    int rc = hf_read_info(s, t, &info);
    if (rc) return rc;
    if (which != 0 && which != 1) return fail(HF_ERR_BAD_ARG, "which must be 0 or 1");
    if (!initial || !final_ || !tau || !n) return fail(HF_ERR_BAD_ARG, "null argument");
    const int count = which ? info.n_move_h : info.n_move_e;
    int32_t cap = *n;
    if ((rc = read_list(s, t, s->P.ev_init[which], s->P.cap_c, count, initial, &cap))) return rc;
    cap = *n;
    if ((rc = read_list(s, t, s->P.ev_final[which], s->P.cap_c, count, final_, &cap))) return rc;
    if (count > 0) {
        HF_CUDA(cudaMemcpyAsync(tau, s->P.ev_tau[which] + t * s->P.cap_c, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
    }
    *n = count;
    return HF_OK;
}

int hf_read_last_events(hf_sim *s, int64_t t, hf_event *out, int32_t *n) {
    hf_traj_info info;
    int rc = hf_read_info(s, t, &info);
    if (rc) return rc;
    if (!out || !n) return fail(HF_ERR_BAD_ARG, "null argument");
    HfScalars sc;
    HF_CUDA(cudaMemcpyAsync(&sc, s->d_sc + t, sizeof(sc), cudaMemcpyDeviceToHost, s->stream));
    int32_t ci[HF_CACHE], cf[HF_CACHE], ck[HF_CACHE];
    double ct[HF_CACHE];
    const size_t pitch4 = sizeof(int32_t) * (size_t)s->P.B, pitch8 = sizeof(double) * (size_t)s->P.B;   // ring is [slot][trajectory]
    HF_CUDA(cudaMemcpy2DAsync(ci, sizeof(int32_t), s->P.cache_init + t, pitch4, sizeof(int32_t), HF_CACHE, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpy2DAsync(cf, sizeof(int32_t), s->P.cache_final + t, pitch4, sizeof(int32_t), HF_CACHE, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpy2DAsync(ck, sizeof(int32_t), s->P.cache_kp + t, pitch4, sizeof(int32_t), HF_CACHE, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpy2DAsync(ct, sizeof(double), s->P.cache_tau + t, pitch8, sizeof(double), HF_CACHE, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    const int count = sc.cache_n;
    for (int i = 0; i < count; ++i) {          // oldest first
        const int slot = ((sc.cache_head - count + i) % HF_CACHE + HF_CACHE) % HF_CACHE;
        memset(&out[i], 0, sizeof(hf_event));
        out[i].initial = unpack_site(s, ci[slot]);
        out[i].final_ = unpack_site(s, cf[slot]);
        out[i].tau = ct[slot];
        out[i].kind = (uint8_t)(ck[slot] & 0xff);
        out[i].particule = (uint8_t)(ck[slot] >> 8);
    }
    *n = count;
    return HF_OK;
}

int hf_set_trace(hf_sim *s, int64_t first, int64_t n, int64_t capacity) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_trace); cudaFree(s->d_trace_base);
    s->d_trace = nullptr; s->d_trace_base = nullptr;
    s->P.trace = nullptr; s->P.trace_base = nullptr;
    s->trace_base.clear();
    s->P.trace_first = s->P.trace_n = s->P.trace_cap = 0;
    if (n > 0) {
        if (first < 0 || first + n > s->P.B || capacity < 1) return fail(HF_ERR_BAD_ARG, "bad trace window");
        if ((rc = dev_alloc(&s->d_trace, (size_t)(n * capacity)))) return rc;
        if ((rc = dev_alloc(&s->d_trace_base, (size_t)n))) return rc;
        HF_CUDA(cudaMemsetAsync(s->d_trace, 0, sizeof(HfTraceRec) * (size_t)(n * capacity), s->stream));
        // records are indexed from the events each trajectory has fired so far: arming the trace after earlier
        // hf_run calls loses no slot (hf_inject restarts the count)
        s->trace_base.assign((size_t)n, 0);
        if (s->injected) {
            std::vector<HfScalars> sc((size_t)n);
            HF_CUDA(cudaMemcpyAsync(sc.data(), s->d_sc + first, sizeof(HfScalars) * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
            HF_CUDA(cudaStreamSynchronize(s->stream));
            for (int64_t i = 0; i < n; ++i) s->trace_base[(size_t)i] = sc[(size_t)i].fired;
        }
        HF_CUDA(cudaMemcpyAsync(s->d_trace_base, s->trace_base.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
        s->P.trace = s->d_trace; s->P.trace_base = s->d_trace_base;
        s->P.trace_first = first; s->P.trace_n = n; s->P.trace_cap = capacity;
    }
    cudaFree(s->d_x_trace_len);
    s->d_x_trace_len = nullptr;
    if (n > 0) {
        if ((rc = dev_alloc(&s->d_x_trace_len, (size_t)n))) return rc;
        HF_CUDA(cudaMemsetAsync(s->d_x_trace_len, 0, sizeof(uint64_t) * (size_t)n, s->stream));
    }
    return HF_OK;
}

int hf_read_trace(hf_sim *s, int64_t t, hf_event *out, int64_t *n) {
    hf_traj_info info;
    int rc = hf_read_info(s, t, &info);
    if (rc) return rc;
    if (!out || !n) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->d_trace || t < s->P.trace_first || t >= s->P.trace_first + s->P.trace_n)
        return fail(HF_ERR_STATE, "trajectory %lld is not traced", (long long)t);
    const int64_t since = (int64_t)(info.fired - s->trace_base[(size_t)(t - s->P.trace_first)]);
    int64_t count = since < s->P.trace_cap ? since : s->P.trace_cap;
    if (*n < count) return fail(HF_ERR_BAD_ARG, "buffer too small: need %lld", (long long)count);
    if (count > 0) {
        HF_CUDA(cudaMemcpyAsync(out, s->d_trace + (t - s->P.trace_first) * s->P.trace_cap, sizeof(hf_event) * (size_t)count,
                                cudaMemcpyDeviceToHost, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
    }
    *n = count;
    return HF_OK;
}

int hf_eval_hops(hf_sim *s, int64_t t, int32_t particule, int32_t n, const int32_t *initial, const int32_t *final_,
                 const double *u, double *tau, double *rate, int32_t *zero_division) {
    int rc = check_traj(s, t);
    if (rc) return rc;
    if (!s->injected) return fail(HF_ERR_STATE, "hf_eval_hops before hf_inject");
    if (particule != HF_P_ELECTRON && particule != HF_P_HOLE) return fail(HF_ERR_BAD_ARG, "particule must be 1 or 2");
    if (n < 1 || !initial || !final_ || !u || !tau || !rate || !zero_division) return fail(HF_ERR_BAD_ARG, "bad argument");
    for (int i = 0; i < n; ++i)
        if (initial[i] < 0 || initial[i] >= s->P.N || final_[i] < 0 || final_[i] >= s->P.N) return fail(HF_ERR_BAD_ARG, "site out of range");
    if ((rc = use_device(s))) return rc;
    int32_t *d_i = nullptr, *d_zd = nullptr;
    double *d_d = nullptr;
    if ((rc = dev_alloc(&d_i, (size_t)(2 * n)))) return rc;
    if ((rc = dev_alloc(&d_zd, (size_t)n))) { cudaFree(d_i); return rc; }
    if ((rc = dev_alloc(&d_d, (size_t)(3 * n)))) { cudaFree(d_i); cudaFree(d_zd); return rc; }
    cudaMemcpyAsync(d_i, initial, 4 * (size_t)n, cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(d_i + n, final_, 4 * (size_t)n, cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(d_d, u, 8 * (size_t)n, cudaMemcpyHostToDevice, s->stream);
    const size_t smem = hf_warp_smem_bytes(s->P.cap_c, s->P.cap_f, hf_xcap(s->P));
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(hf_eval_hops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        hf_eval_hops_kernel<<<1, 32, smem, s->stream>>>(s->P, t, particule - 1, n, d_i, d_i + n, d_d, d_d + n, d_d + 2 * n, d_zd);
        cudaMemcpyAsync(tau, d_d + n, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
        cudaMemcpyAsync(rate, d_d + 2 * n, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
        cudaMemcpyAsync(zero_division, d_zd, 4 * (size_t)n, cudaMemcpyDeviceToHost, s->stream);
        e = cudaStreamSynchronize(s->stream);
    }
    cudaFree(d_i); cudaFree(d_zd); cudaFree(d_d);
    if (e != cudaSuccess) return fail(HF_ERR_CUDA, "hf_eval_hops failed: %s", cudaGetErrorString(e));
    return HF_OK;
}

// ---- exciton trajectories (Path X; model specified by oracle/exciton_oracle.c) ----------------------

int hf_x_default_params(hf_x_params *p) {
    if (!p) return fail(HF_ERR_BAD_ARG, "null argument");
    memset(p, 0, sizeof(*p));
    // None of these values comes from the reference (it has no exciton dynamics).  They are
    // order-of-magnitude literature values for a DPEPO / ACRSA / TBPe hyperfluorescent stack.
    p->cutoff = 4.0;
    const double r0[3][3] = {{1.0, 3.0, 3.5},      // host  -> host, TADF, fluorescent
                             {0.5, 2.0, 4.0},      // TADF  -> ...   (TADF -> emitter is the hyperfluorescence channel)
                             {0.5, 1.0, 2.5}};     // fluo  -> ...
    const double kd0[3][3] = {{1e10, 1e10, 1e10}, {1e10, 1e10, 1e10}, {1e10, 1e10, 1e10}};
    memcpy(p->forster_radius, r0, sizeof(r0));
    memcpy(p->dexter_prefactor, kd0, sizeof(kd0));
    p->dexter_length = 0.5;
    const double isc[3] = {1e5, 1e7, 1e5}, risc[3] = {0.0, 1e7, 0.0};
    memcpy(p->k_isc, isc, sizeof(isc));
    memcpy(p->k_risc, risc, sizeof(risc));
    const double rad[3][2] = {{0.0, 0.0}, {1e7, 0.0}, {2e8, 0.0}};          // hosts do not emit in the visible (molecule.py:396-400)
    const double nonrad[3][2] = {{2e8, 1e3}, {1e6, 1e4}, {1e7, 1e5}};
    memcpy(p->k_rad, rad, sizeof(rad));
    memcpy(p->k_nonrad, nonrad, sizeof(nonrad));
    p->singlet_fraction = 0.25;                                              // molecule.py:93
    p->outcoupling = 0.2;
    return HF_OK;
}

int hf_x_configure(hf_sim *s, const hf_x_params *p) {
    if (!s || !p) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->have_lattice) return fail(HF_ERR_STATE, "hf_x_configure before a lattice exists");
    for (int r = 0; r < s->n_real; ++r)
        if (!s->have_st[(size_t)r])
            return fail(HF_ERR_STATE, "realisation %d was uploaded without S1 / T1 energies: the exciton rates need them", r);
    if (!(p->cutoff >= 1.0) || p->cutoff > 6.0)      // <= 32 slots per lane: (M + 3) <= 1024 channels
        return fail(HF_ERR_BAD_ARG, "cutoff must be in [1, 6] lattice constants");
    const int R = (int)std::floor(p->cutoff);
    if (s->P.X < 2 * R + 1 || s->P.Y < 2 * R + 1)
        return fail(HF_ERR_BAD_ARG, "x and y must be >= 2*cutoff+1 = %d for unique periodic images", 2 * R + 1);
    if (!(p->dexter_length > 0.0)) return fail(HF_ERR_BAD_ARG, "dexter_length must be positive");
    if (!(p->singlet_fraction >= 0.0 && p->singlet_fraction <= 1.0)) return fail(HF_ERR_BAD_ARG, "singlet_fraction must be a probability");
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b)
            if (!(p->forster_radius[a][b] >= 0.0) || !(p->dexter_prefactor[a][b] >= 0.0))
                return fail(HF_ERR_BAD_ARG, "transfer parameters must be >= 0");
        if (!(p->k_isc[a] >= 0.0) || !(p->k_risc[a] >= 0.0)) return fail(HF_ERR_BAD_ARG, "spin-flip rates must be >= 0");
        for (int sp = 0; sp < 2; ++sp) {
            if (!(p->k_rad[a][sp] >= 0.0) || !(p->k_nonrad[a][sp] >= 0.0)) return fail(HF_ERR_BAD_ARG, "decay rates must be >= 0");
            // every exciton must be able to end: with no decay channel the total rate of an isolated site can be 0
            if (!(p->k_rad[a][sp] + p->k_nonrad[a][sp] > 0.0))
                return fail(HF_ERR_BAD_ARG, "species %d needs a non-zero decay rate for spin index %d", a, sp);
        }
    }
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    s->x_params = *p;
    // neighbour table: dz-major, dx fastest (consecutive lanes read consecutive x)
    s->x_offsets.clear(); s->x_g6.clear(); s->x_gd.clear();
    std::vector<int32_t> packed, lin;
    const double rc2 = p->cutoff * p->cutoff;
    for (int dz = -R; dz <= R; ++dz)
        for (int dy = -R; dy <= R; ++dy)
            for (int dx = -R; dx <= R; ++dx) {
                const int d2 = dx * dx + dy * dy + dz * dz;
                if (d2 == 0 || (double)d2 > rc2) continue;
                s->x_offsets.push_back((int8_t)dx); s->x_offsets.push_back((int8_t)dy); s->x_offsets.push_back((int8_t)dz);
                const double d2d = (double)d2;
                s->x_g6.push_back(1.0 / ((d2d * d2d) * d2d));                       // r^-6, lattice constant 1 nm
                s->x_gd.push_back(std::exp(-2.0 * std::sqrt(d2d) / p->dexter_length));
                packed.push_back((dx + 128) | ((dy + 128) << 8) | ((dz + 128) << 16));
                lin.push_back((int32_t)((int64_t)dz * s->P.X * s->P.Y + (int64_t)dy * s->P.X + dx));
            }
    HfxParams &X = s->X;
    memset(&X, 0, sizeof(X));
    X.X = s->P.X; X.Y = s->P.Y; X.Z = s->P.Z; X.N = s->P.N; X.B = s->P.B; X.traj_per_real = s->P.traj_per_real;
    X.k0 = s->P.k0; X.k1 = s->P.k1; X.first_traj = s->P.first_traj;
    X.M = (int32_t)packed.size();
    X.R = (int32_t)std::ceil(p->cutoff);
    for (int a = 0; a < 3; ++a) {
        const double ks = p->k_rad[a][0] + p->k_nonrad[a][0];                       // donor singlet decay rate
        for (int b = 0; b < 3; ++b) {
            const double r0 = p->forster_radius[a][b], r3 = (r0 * r0) * r0;
            X.kf[3 * a + b] = ks * (r3 * r3);
            X.kd[3 * a + b] = p->dexter_prefactor[a][b];
        }
        X.k_isc[a] = p->k_isc[a]; X.k_risc[a] = p->k_risc[a];
        for (int sp = 0; sp < 2; ++sp) { X.k_rad[2 * a + sp] = p->k_rad[a][sp]; X.k_nonrad[2 * a + sp] = p->k_nonrad[a][sp]; }
    }
    X.singlet_fraction = p->singlet_fraction;
    // weights: R zero planes below and above every realisation (an open-face transfer reads weight 0)
    X.w_halo = (int64_t)X.R * s->P.X * s->P.Y;
    X.w_stride = s->P.N + 2 * X.w_halo;
    const size_t M = packed.size(), RN = (size_t)s->n_real * (size_t)X.w_stride, B = (size_t)s->P.B;
    if ((rc = dev_reuse(s, &s->d_x_off, M))) return rc;
    if ((rc = dev_reuse(s, &s->d_x_lin, M))) return rc;
    if ((rc = dev_reuse(s, &s->d_x_g6, M))) return rc;
    if ((rc = dev_reuse(s, &s->d_x_gd, M))) return rc;
    HF_CUDA(cudaMemcpyAsync(s->d_x_off, packed.data(), 4 * M, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaMemcpyAsync(s->d_x_lin, lin.data(), 4 * M, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaMemcpyAsync(s->d_x_g6, s->x_g6.data(), 8 * M, cudaMemcpyHostToDevice, s->stream));
    HF_CUDA(cudaMemcpyAsync(s->d_x_gd, s->x_gd.data(), 8 * M, cudaMemcpyHostToDevice, s->stream));
    if ((rc = dev_reuse(s, &s->d_ws1, RN))) return rc;
    if ((rc = dev_reuse(s, &s->d_wt1, RN))) return rc;
    HF_CUDA(cudaMemsetAsync(s->d_ws1, 0, 8 * RN, s->stream));
    HF_CUDA(cudaMemsetAsync(s->d_wt1, 0, 8 * RN, s->stream));
    if (!s->d_x_site) {
        if ((rc = dev_alloc(&s->d_x_site, B))) return rc;
        if ((rc = dev_alloc(&s->d_x_spin, B))) return rc;
        if ((rc = dev_alloc(&s->d_x_time, B))) return rc;
        if ((rc = dev_alloc(&s->d_x_life, B))) return rc;
        if ((rc = dev_alloc(&s->d_x_draws, B))) return rc;
        if ((rc = dev_alloc(&s->d_x_totals, (size_t)s->n_real * HFX_T_N))) return rc;
    }
    hfx_weights_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>(s->d_species, s->d_s1, s->d_t1, s->d_ws1, s->d_wt1, (int64_t)s->P.N,
                                                               (int64_t)s->n_real, X.w_stride, X.w_halo);
    X.offsets = s->d_x_off; X.lin = s->d_x_lin; X.g6 = s->d_x_g6; X.gd = s->d_x_gd;
    X.ws1 = s->d_ws1; X.wt1 = s->d_wt1;
    X.site = s->d_x_site; X.spin = s->d_x_spin; X.time = s->d_x_time; X.lifetime = s->d_x_life; X.draws = s->d_x_draws;
    X.totals = s->d_x_totals;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    HF_CUDA(cudaGetLastError());
    if ((rc = x_configure_launch(s))) return rc;
    s->x_configured = true;
    s->x_spawned = false;
    if (s->P.xi.mode == 2 && (rc = sync_integrated(s))) return rc;   // the integrated mode reads these tables and weights
    s->xd_configured = false; s->xd_spawned = false;      // the dense devices share these tables and weights
    return HF_OK;
}

int hf_x_spawn(hf_sim *s) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->x_configured) return fail(HF_ERR_STATE, "hf_x_spawn before hf_x_configure");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaMemsetAsync(s->d_x_totals, 0, sizeof(unsigned long long) * (size_t)s->n_real * HFX_T_N, s->stream));
    if (s->d_x_trace_len) HF_CUDA(cudaMemsetAsync(s->d_x_trace_len, 0, sizeof(uint64_t) * (size_t)s->P.trace_n, s->stream));
    const int threads = 256;
    hfx_spawn_kernel<<<(unsigned)((s->X.B + threads - 1) / threads), threads, 0, s->stream>>>(s->X);
    HF_CUDA(cudaGetLastError());
    s->x_spawned = true;
    return HF_OK;
}

int hf_x_run(hf_sim *s, int64_t n_events) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->x_spawned) return fail(HF_ERR_STATE, "hf_x_run before hf_x_spawn");
    if (n_events < 0) return fail(HF_ERR_BAD_ARG, "n_events must be >= 0");
    int rc = use_device(s);
    if (rc) return rc;
    if (s->x_tune && !s->d_trace && n_events > 0 && (rc = x_autotune(s))) return rc;
    s->X.n_events = n_events;
    s->X.trace = s->d_trace; s->X.trace_first = s->P.trace_first; s->X.trace_n = s->P.trace_n; s->X.trace_cap = s->P.trace_cap;
    s->X.trace_len = s->d_x_trace_len;
    if (s->timing.size() >= 1024) { HF_CUDA(cudaStreamSynchronize(s->stream)); fold_timing(s); }
    cudaEvent_t e0 = get_event(s), e1 = get_event(s);
    HF_CUDA(cudaEventRecord(e0, s->stream));
    x_pick_kernel(s->X.M, s->x_brick)<<<s->x_grid, kXWarps * 32, s->x_smem, s->stream>>>(x_launch_params(s));
    HF_CUDA(cudaEventRecord(e1, s->stream));
    s->timing.emplace_back(e0, e1);
    HF_CUDA(cudaGetLastError());
    return HF_OK;
}

int hf_x_read_realisation_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->x_configured) return fail(HF_ERR_STATE, "exciton path not configured");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaMemcpyAsync(out, s->d_x_totals, sizeof(uint64_t) * (size_t)s->n_real * HFX_T_N, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    return HF_OK;
}

int hf_x_read_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    std::vector<uint64_t> all((size_t)(s ? s->n_real : 0) * HFX_T_N);
    int rc = hf_x_read_realisation_tallies(s, all.data());
    if (rc) return rc;
    for (int k = 0; k < HFX_T_N; ++k) out[k] = 0;
    for (int r = 0; r < s->n_real; ++r)
        for (int k = 0; k < HFX_T_N; ++k) out[k] += all[(size_t)r * HFX_T_N + k];
    return HF_OK;
}

int hf_x_read_tables(hf_sim *s, int32_t *n_offsets, int8_t *offsets, double *g6, double *gd, double *kf, double *kd) {
    if (!s || !n_offsets) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->x_configured) return fail(HF_ERR_STATE, "exciton path not configured");
    const int32_t M = s->X.M;
    if (offsets || g6 || gd) {
        if (*n_offsets < M) return fail(HF_ERR_BAD_ARG, "buffer too small: need %d offsets", M);
        if (offsets) memcpy(offsets, s->x_offsets.data(), (size_t)M * 3);
        if (g6) memcpy(g6, s->x_g6.data(), sizeof(double) * (size_t)M);
        if (gd) memcpy(gd, s->x_gd.data(), sizeof(double) * (size_t)M);
    }
    if (kf) memcpy(kf, s->X.kf, sizeof(double) * 9);
    if (kd) memcpy(kd, s->X.kd, sizeof(double) * 9);
    *n_offsets = M;
    return HF_OK;
}

int hf_x_launch_info(hf_sim *s, int32_t *block, int32_t *brick, int32_t *grid) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->x_configured) return fail(HF_ERR_STATE, "exciton path not configured");
    if (block) *block = s->X.block;
    if (brick) *brick = s->x_brick ? 1 : 0;
    if (grid) *grid = s->x_grid;
    return HF_OK;
}

int hf_x_download_weights(hf_sim *s, int32_t r, double *ws1, double *wt1) {
    if (!s || !ws1 || !wt1) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->x_configured) return fail(HF_ERR_STATE, "exciton path not configured");
    if (r < 0 || r >= s->n_real) return fail(HF_ERR_BAD_ARG, "realisation %d out of range", r);
    int rc = use_device(s);
    if (rc) return rc;
    const size_t N = (size_t)s->P.N, off = (size_t)r * (size_t)s->X.w_stride + (size_t)s->X.w_halo;
    HF_CUDA(cudaMemcpyAsync(ws1, s->d_ws1 + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(wt1, s->d_wt1 + off, N * 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    return HF_OK;
}

int hf_x_read_state(hf_sim *s, int64_t first, int64_t n, int32_t *site, int32_t *spin, double *time, uint64_t *draws,
                    double *lifetime_sum) {
    int rc = check_traj(s, first);
    if (rc) return rc;
    if (!s->x_spawned) return fail(HF_ERR_STATE, "no excitons yet");
    if (n < 1 || first + n > s->P.B) return fail(HF_ERR_BAD_ARG, "bad range");
    if ((rc = use_device(s))) return rc;
    std::vector<int32_t> packed((size_t)n);
    HF_CUDA(cudaMemcpyAsync(packed.data(), s->d_x_site + first, 4 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (spin) HF_CUDA(cudaMemcpyAsync(spin, s->d_x_spin + first, 4 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (time) HF_CUDA(cudaMemcpyAsync(time, s->d_x_time + first, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (draws) HF_CUDA(cudaMemcpyAsync(draws, s->d_x_draws + first, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    if (lifetime_sum) HF_CUDA(cudaMemcpyAsync(lifetime_sum, s->d_x_life + first, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    if (site) for (int64_t i = 0; i < n; ++i) site[i] = unpack_site(s, packed[(size_t)i]);
    return HF_OK;
}

// ---- high-density exciton devices (Path XD; model specified by oracle/exciton_dense_oracle.c) ------------

#ifndef HF_XD_WARPS
#define HF_XD_WARPS 4
#endif
constexpr int kXdWarps = HF_XD_WARPS;

int hf_xd_configure(hf_sim *s, int32_t n_excitons, const double *annihilation, double p_tta) {
    if (!s || !annihilation) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->x_configured) return fail(HF_ERR_STATE, "hf_xd_configure before hf_x_configure");
    if (n_excitons < 1 || n_excitons > HFXD_MAX_EXCITONS)
        return fail(HF_ERR_BAD_ARG, "excitons per device must be in [1, %d]", HFXD_MAX_EXCITONS);
    const int64_t n_interior = (int64_t)s->P.X * s->P.Y * (s->P.Z - 2);
    if ((int64_t)n_excitons * 2 > n_interior) return fail(HF_ERR_BAD_ARG, "more excitons than half the interior sites");
    if (!(p_tta >= 0.0 && p_tta <= 1.0)) return fail(HF_ERR_BAD_ARG, "p_tta must be a probability");
    if (hfx_padded_channels(s->X.M) / 32 > 31) return fail(HF_ERR_UNSUPPORTED, "cutoff too large for the dense devices: %d channels (at most 989)", s->X.M + 3);
    for (int k = 0; k < 4; ++k)
        if (!(annihilation[k] >= 0.0)) return fail(HF_ERR_BAD_ARG, "annihilation factors must be >= 0");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaStreamSynchronize(s->stream));
    HfxdParams &D = s->XD;
    memset(&D, 0, sizeof(D));
    D.X = s->X;
    D.n = n_excitons; D.n_pad = (n_excitons + 31) & ~31;
    for (int k = 0; k < 4; ++k) D.ann[k] = annihilation[k];
    D.p_tta = p_tta;
    D.occ_words = (s->X.w_stride + 15) / 16;
    const size_t B = (size_t)s->P.B, np = (size_t)D.n_pad;
    if ((rc = dev_reuse(s, &s->d_xd_pos, B * np))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_T, B * np))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_birth, B * np))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_rec_L, B * np * 32))) return rc;             // 384 B per exciton: the record of its last evaluation
    if ((rc = dev_reuse(s, &s->d_xd_rec_seen, B * np * 32))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_occ, B * (size_t)D.occ_words))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_time, B))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_life, B))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_draws, B))) return rc;
    if ((rc = dev_reuse(s, &s->d_xd_totals, (size_t)s->n_real * HFXD_T_N))) return rc;
    D.pos = s->d_xd_pos; D.T = s->d_xd_T; D.birth = s->d_xd_birth; D.occ = s->d_xd_occ;
    D.rec_L = s->d_xd_rec_L; D.rec_seen = s->d_xd_rec_seen;
    D.time = s->d_xd_time; D.life = s->d_xd_life; D.draws = s->d_xd_draws; D.totals = s->d_xd_totals;
    const size_t smem = hfxd_smem_bytes(s->X.M, D.n_pad, kXdWarps);
    if (smem > 227 * 1024) return fail(HF_ERR_UNSUPPORTED, "%d excitons per device need %zu B of shared memory", n_excitons, smem);
    if (smem > 48 * 1024) HF_CUDA(cudaFuncSetAttribute(hfxd_run_kernel<kXdWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hfxd_run_kernel<kXdWarps>, kXdWarps * 32, smem));
    if (per_sm < 1) return fail(HF_ERR_UNSUPPORTED, "dense exciton kernel does not fit an SM");
    const int64_t need = ((int64_t)B + kXdWarps - 1) / kXdWarps, cap = (int64_t)s->sm_count * per_sm;
    s->xd_smem = smem; s->xd_grid = (int)(need < cap ? need : cap);
    if (s->xd_grid < 1) s->xd_grid = 1;
    s->xd_configured = true; s->xd_spawned = false;
    return HF_OK;
}

static int xd_launch(hf_sim *s, int64_t n_events, int fill) {
    HfxdParams &D = s->XD;
    D.fill = fill;
    D.X.n_events = n_events;
    D.X.trace = s->d_trace; D.X.trace_first = s->P.trace_first; D.X.trace_n = s->P.trace_n; D.X.trace_cap = s->P.trace_cap;
    D.X.trace_len = s->d_x_trace_len;
    if (s->timing.size() >= 1024) { HF_CUDA(cudaStreamSynchronize(s->stream)); fold_timing(s); }
    cudaEvent_t e0 = get_event(s), e1 = get_event(s);
    HF_CUDA(cudaEventRecord(e0, s->stream));
    hfxd_run_kernel<kXdWarps><<<s->xd_grid, kXdWarps * 32, s->xd_smem, s->stream>>>(D);
    HF_CUDA(cudaEventRecord(e1, s->stream));
    if (!fill) s->timing.emplace_back(e0, e1);
    else { s->event_pool.push_back(e0); s->event_pool.push_back(e1); }
    HF_CUDA(cudaGetLastError());
    return HF_OK;
}

int hf_xd_spawn(hf_sim *s) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->xd_configured) return fail(HF_ERR_STATE, "hf_xd_spawn before hf_xd_configure");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaMemsetAsync(s->d_xd_occ, 0, 4 * (size_t)s->P.B * (size_t)s->XD.occ_words, s->stream));
    HF_CUDA(cudaMemsetAsync(s->d_xd_totals, 0, sizeof(unsigned long long) * (size_t)s->n_real * HFXD_T_N, s->stream));
    if (s->d_x_trace_len) HF_CUDA(cudaMemsetAsync(s->d_x_trace_len, 0, sizeof(uint64_t) * (size_t)s->P.trace_n, s->stream));
    if ((rc = xd_launch(s, 0, 1))) return rc;
    s->xd_spawned = true;
    return HF_OK;
}

int hf_xd_run(hf_sim *s, int64_t n_events) {
    if (!s) return fail(HF_ERR_BAD_ARG, "null handle");
    if (!s->xd_spawned) return fail(HF_ERR_STATE, "hf_xd_run before hf_xd_spawn");
    if (n_events < 0) return fail(HF_ERR_BAD_ARG, "n_events must be >= 0");
    int rc = use_device(s);
    if (rc) return rc;
    return xd_launch(s, n_events, 0);
}

int hf_xd_read_realisation_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->xd_spawned) return fail(HF_ERR_STATE, "no dense exciton devices yet");
    int rc = use_device(s);
    if (rc) return rc;
    HF_CUDA(cudaMemcpyAsync(out, s->d_xd_totals, sizeof(uint64_t) * (size_t)s->n_real * HFXD_T_N, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    return HF_OK;
}

int hf_xd_read_tallies(hf_sim *s, uint64_t *out) {
    if (!s || !out) return fail(HF_ERR_BAD_ARG, "null argument");
    std::vector<uint64_t> all((size_t)s->n_real * HFXD_T_N);
    int rc = hf_xd_read_realisation_tallies(s, all.data());
    if (rc) return rc;
    for (int k = 0; k < HFXD_T_N; ++k) out[k] = 0;
    for (int r = 0; r < s->n_real; ++r)
        for (int k = 0; k < HFXD_T_N; ++k) out[k] += all[(size_t)r * HFXD_T_N + k];
    return HF_OK;
}

int hf_xd_read_device(hf_sim *s, int64_t device, int32_t *site, int32_t *spin, double *total_rate, double *birth, double *scalars) {
    int rc = check_traj(s, device);
    if (rc) return rc;
    if (!s->xd_spawned) return fail(HF_ERR_STATE, "no dense exciton devices yet");
    if ((rc = use_device(s))) return rc;
    const size_t n = (size_t)s->XD.n, np = (size_t)s->XD.n_pad;
    std::vector<uint32_t> pos(n);
    HF_CUDA(cudaMemcpyAsync(pos.data(), s->d_xd_pos + (size_t)device * np, 4 * n, cudaMemcpyDeviceToHost, s->stream));
    if (total_rate) HF_CUDA(cudaMemcpyAsync(total_rate, s->d_xd_T + (size_t)device * np, 8 * n, cudaMemcpyDeviceToHost, s->stream));
    if (birth) HF_CUDA(cudaMemcpyAsync(birth, s->d_xd_birth + (size_t)device * np, 8 * n, cudaMemcpyDeviceToHost, s->stream));
    uint64_t draws = 0;
    double tl[2] = {0.0, 0.0};
    HF_CUDA(cudaMemcpyAsync(&tl[0], s->d_xd_time + device, 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(&tl[1], s->d_xd_life + device, 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaMemcpyAsync(&draws, s->d_xd_draws + device, 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t k = 0; k < n; ++k) {
        if (site) site[k] = unpack_site(s, (int32_t)(pos[k] & 0x3fffffffu));
        if (spin) spin[k] = (int32_t)(pos[k] >> 30);
    }
    if (scalars) { scalars[0] = tl[0]; scalars[1] = tl[1]; scalars[2] = (double)draws; }
    return HF_OK;
}

int hf_xd_read_occupancy(hf_sim *s, int64_t device, uint8_t *codes) {
    int rc = check_traj(s, device);
    if (rc) return rc;
    if (!codes) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->xd_spawned) return fail(HF_ERR_STATE, "no dense exciton devices yet");
    if ((rc = use_device(s))) return rc;
    std::vector<uint32_t> words((size_t)s->XD.occ_words);
    HF_CUDA(cudaMemcpyAsync(words.data(), s->d_xd_occ + (size_t)device * (size_t)s->XD.occ_words, 4 * words.size(), cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < s->P.N; ++i) {
        const int64_t idx = i + s->X.w_halo;
        codes[i] = (uint8_t)((words[(size_t)(idx >> 4)] >> (2 * (int)(idx & 15))) & 3u);
    }
    return HF_OK;
}

int hf_x_read_trace(hf_sim *s, int64_t t, hf_event *out, int64_t *n) {
    int rc = check_traj(s, t);
    if (rc) return rc;
    if (!out || !n) return fail(HF_ERR_BAD_ARG, "null argument");
    if (!s->d_trace || !s->d_x_trace_len || t < s->P.trace_first || t >= s->P.trace_first + s->P.trace_n)
        return fail(HF_ERR_STATE, "trajectory %lld is not traced", (long long)t);
    if ((rc = use_device(s))) return rc;
    uint64_t len = 0;
    HF_CUDA(cudaMemcpyAsync(&len, s->d_x_trace_len + (t - s->P.trace_first), 8, cudaMemcpyDeviceToHost, s->stream));
    HF_CUDA(cudaStreamSynchronize(s->stream));
    int64_t count = (int64_t)len < s->P.trace_cap ? (int64_t)len : s->P.trace_cap;
    if (*n < count) return fail(HF_ERR_BAD_ARG, "buffer too small: need %lld", (long long)count);
    if (count > 0) {
        HF_CUDA(cudaMemcpyAsync(out, s->d_trace + (t - s->P.trace_first) * s->P.trace_cap, sizeof(hf_event) * (size_t)count,
                                cudaMemcpyDeviceToHost, s->stream));
        HF_CUDA(cudaStreamSynchronize(s->stream));
    }
    *n = count;
    return HF_OK;
}

}  // extern "C"
