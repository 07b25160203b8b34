// Integrated exciton dynamics: the exciton formed by a `bound` event lives on in the same First Reaction Method
// as the carriers, in the slots the reference reserves for it (_move_exciton_events / _isc_events / _decay_events
// reseau.py:233-235, the `...` branches of the dispatcher 526-550, the routing comment 533,
// TADF.intersystem_crossing molecule.py:299-307).  Included by hf_frm.cuh (before hf_step).
//
// OWN MODEL, PARITY UNPINNED beyond mode 1: the specification is the "integrated excitons" block of
// the Path P restatement under oracle/ (its x_new_event and the xsel branch of first_reaction_method; DESIGN.md
// section 11); every statement below follows it operation for operation, so event sequences and taus are
// bit-identical to it (log() aside).  HfParams.xi.mode:
//   0  off — the reference as it is: the decay event is the `special` slot with tau = 0
//   1  "decay at once" THROUGH the exciton slots and lists: must reproduce every golden trajectory of the
//      unmodified reference bit for bit (the gate of tests/test_integrated_excitons.py)
//   2  dynamics: transfer (singlet Forster / triplet Dexter), ISC / RISC, radiative / non-radiative decay with the
//      rate laws, tables and tagged Boltzmann weights of the exciton path (hf_exciton.cuh, hf_x_configure)
//
// One exciton = one slot (site, meta, pending event's final site, tau); slots are kept in creation order.
// meta = spin (bits 0-1: 0 none, 1 singlet, 3 triplet) | on-formation-site << 2 | radiative << 3 | kind << 4.
#pragma once

__device__ __forceinline__ int hfi_tag_of(double w) { return (int)(__double_as_longlong(w) & 3ll); }
__device__ __forceinline__ double hfi_min1(double v) { return v < 1.0 ? v : 1.0; }
__device__ __forceinline__ int hfi_class(int kind) { return kind == HF_K_DECAY ? 2 : (kind == HF_K_ISC ? 1 : 0); }

// neighbour of packed site p along packed offset o: periodic x, y; z straight through (the halo absorbs z outside
// [0, Z): weight +0.0, rate +0.0).  Returns its index relative to site 0 (may lie in the halo) and its packed form.
__device__ __forceinline__ int64_t hfi_neighbour(int X, int Y, int p, int o, int *packed) {
    const int dx = (o & 255) - 128, dy = ((o >> 8) & 255) - 128, dz = ((o >> 16) & 255) - 128;
    int xx = hf_px(p) + dx, yy = hf_py(p) + dy;
    const int zz = hf_pz(p) + dz;
    if (xx < 0) xx += X; else if (xx >= X) xx -= X;
    if (yy < 0) yy += Y; else if (yy >= Y) yy -= Y;
    *packed = hf_pack(xx, yy, zz < 0 ? 0 : zz);
    return ((int64_t)zz * Y + yy) * X + xx;
}

struct HfiSite {                      // what every channel of one exciton shares
    const double *w;                  // the manifold's weights of this realisation, halo skipped
    const double *g;                  // g6 (singlet) or gd (triplet)
    const int32_t *off;
    double pref0, pref1, pref2;       // kf[a][.] or kd[a][.]
    double inv_wi, on0, on1, on2;     // 1 / w[i]; ISC-or-RISC, radiative, non-radiative rates
    int site, M, X, Y;
};

__device__ __forceinline__ double hfi_rate(const HfiSite &s, int c) {
    if (c < s.M) {
        int q;
        const double wj = __ldg(s.w + hfi_neighbour(s.X, s.Y, s.site, __ldg(s.off + c), &q));
        const int b = hfi_tag_of(wj);
        const double pref = b == 0 ? s.pref0 : (b == 1 ? s.pref1 : s.pref2);
        return __dmul_rn(__dmul_rn(pref, __ldg(s.g + c)), hfi_min1(__dmul_rn(wj, s.inv_wi)));
    }
    return c == s.M ? s.on0 : (c == s.M + 1 ? s.on1 : (c == s.M + 2 ? s.on2 : 0.0));   // 0.0: padding up to a multiple of 32
}

struct HfiEvent { double tau; int32_t fin, kindrad; };     // kindrad = kind << 4 | HFI_RAD

// x_new_event of the specification (mode 2) for an exciton of `spin` on packed `site` of realisation `real`, drawing
// block xdraws / 2 of the trajectory's exciton stream.  L = 32: the lanes own the specification's lanes (scratch: 32
// doubles); L = 1: one thread walks them, keeping every fourth prefix sum (scratch: 8 checkpoints) and re-adding at
// most four lane sums and one lane's slots for the selection — same additions in the same order, hence the same bits.
// NOT inlined: its registers are its own, the step kernels keep theirs.
template <int L>
__device__ __noinline__ HfiEvent hfi_draw_event(const HfiTables *T, int64_t real, uint32_t ident, uint64_t xdraws,
                                                int site, int spin, double *scratch, int lane) {
    const int sidx = spin == HF_XS_SINGLET ? 0 : 1;
    const int X = T->X, Y = T->Y;
    const int64_t i = hf_site(X, Y, site);
    const int64_t wbase = real * T->w_stride + T->w_halo;
    HfiSite s;
    s.w = (sidx == 0 ? T->ws1 : T->wt1) + wbase;
    s.g = sidx == 0 ? T->g6 : T->gd;
    s.off = T->offsets;
    const double wi = __ldg(s.w + i);
    const int a = hfi_tag_of(wi);
    const double *pref = T->k + (sidx == 0 ? 0 : 9) + 3 * a;
    s.pref0 = __ldg(pref); s.pref1 = __ldg(pref + 1); s.pref2 = __ldg(pref + 2);
    s.inv_wi = __ddiv_rn(1.0, wi);
    s.site = site; s.M = T->M; s.X = X; s.Y = Y;
    s.on0 = sidx == 0 ? __ldg(T->k + 18 + a)
                      : __dmul_rn(__ldg(T->k + 21 + a), hfi_min1(__dmul_rn(__ldg(T->ws1 + wbase + i), __ddiv_rn(1.0, __ldg(T->wt1 + wbase + i)))));
    s.on1 = __ldg(T->k + 24 + 2 * a + sidx);
    s.on2 = __ldg(T->k + 30 + 2 * a + sidx);
    const int K = (s.M + 3 + 31) >> 5;
    const HfPhiloxBlock blk = hf_philox_block(T->k0, T->k1, HF_STREAM_PX, ident, xdraws >> 1);
    const double u_time = hf_block_uniform(blk, 0), u_sel = hf_block_uniform(blk, 1);
    double R, target, before = 0.0;
    int ln = -1;
    if constexpr (L == 32) {
        double S = 0.0;
        for (int q = 0; q < K; ++q) S = __dadd_rn(S, hfi_rate(s, 32 * q + lane));
        hf_sync<L>();
        scratch[lane] = S;
        hf_sync<L>();
        R = 0.0;
        for (int l = 0; l < 32; ++l) R = __dadd_rn(R, scratch[l]);
        target = __dmul_rn(u_sel, R);
        double p = 0.0;
        int last_pos = -1;
        for (int l = 0; l < 32; ++l) {
            const double sl = scratch[l], pn = __dadd_rn(p, sl);
            if (sl > 0.0) last_pos = l;
            if (ln < 0 && pn > target) { ln = l; before = p; }
            p = pn;
        }
        if (ln < 0) { ln = last_pos; before = 0.0; for (int l = 0; l < ln; ++l) before = __dadd_rn(before, scratch[l]); }
        hf_sync<L>();
    } else {
        double *cp = scratch;                                                    // P_3, P_7, ... P_31
        double p = 0.0;
        int last_pos = -1;
#pragma unroll 1
        for (int l = 0; l < 32; ++l) {
            double S = 0.0;
#pragma unroll 1
            for (int q = 0; q < K; ++q) S = __dadd_rn(S, hfi_rate(s, 32 * q + l));
            if (S > 0.0) last_pos = l;
            p = __dadd_rn(p, S);
            if ((l & 3) == 3) cp[l >> 2] = p;
        }
        R = p;
        target = __dmul_rn(u_sel, R);
        int grp = 0;
        while (grp < 8 && !(cp[grp] > target)) ++grp;
        const int stop = grp < 8 ? 4 * grp + 4 : last_pos + 1;                   // none exceeds the target: the last lane with S_l > 0
        const int from = grp < 8 ? 4 * grp : (last_pos & ~3);
        p = from ? cp[(from >> 2) - 1] : 0.0;
#pragma unroll 1
        for (int l = from; l < stop; ++l) {
            double S = 0.0;
#pragma unroll 1
            for (int q = 0; q < K; ++q) S = __dadd_rn(S, hfi_rate(s, 32 * q + l));
            const double pn = __dadd_rn(p, S);
            if (pn > target || (grp == 8 && l == last_pos)) { ln = l; before = p; break; }
            p = pn;
        }
    }
    // inside the lane: the prefix restarts from P_(ln-1) and walks the slots
    int slot = -1, last_nonzero = -1;
    double acc = before;
#pragma unroll 1
    for (int q = 0; q < K; ++q) {
        const double r = hfi_rate(s, 32 * q + ln);
        acc = __dadd_rn(acc, r);
        if (r > 0.0) last_nonzero = q;
        if (slot < 0 && acc > target) slot = q;
    }
    if (slot < 0) slot = last_nonzero;
    const int chosen = 32 * slot + ln;
    HfiEvent e;
    e.tau = __ddiv_rn(-log(__dsub_rn(1.0, u_time)), R);
    e.fin = site;
    if (chosen < s.M) {
        hfi_neighbour(X, Y, site, __ldg(s.off + chosen), &e.fin);
        e.kindrad = (sidx == 0 ? HF_K_FORSTER : HF_K_MOVE) << 4;
    } else if (chosen == s.M) {
        e.kindrad = HF_K_ISC << 4;
    } else {
        e.kindrad = (HF_K_DECAY << 4) | (chosen == s.M + 1 ? HFI_RAD : 0);
    }
    return e;
}

// The pending event of exciton slot k (x_new_event of the specification).
template <int L>
__device__ __forceinline__ void hfi_new_event(const HfParams &P, HfWarp &w, int k, int lane) {
    const int site = w.x_site[k];
    const int spin = HFI_SPIN(w.x_meta[k]);
    const int keep = w.x_meta[k] & (3 | HFI_ONSITE);
    HfiEvent e;
    if (P.xi.mode == 1 || spin == 0) {                                           // decay at once (molecule.py:185-199 ...)
        const bool rad = P.xi.mode == 1 && spin == HF_XS_SINGLET && w.species[hf_site(P, site)] != HF_SPECIES_HOST;
        e.tau = 0.0; e.fin = site; e.kindrad = (HF_K_DECAY << 4) | (rad ? HFI_RAD : 0);
    } else {
        e = hfi_draw_event<L>(P.xi.tab, w.local / P.traj_per_real, w.ident, w.xdraws, site, spin, w.x_scratch, lane);
        w.xdraws += 2;
    }
    hf_sync<L>();
    if (lane == 0) { w.x_final[k] = e.fin; w.x_tau[k] = e.tau; w.x_meta[k] = keep | e.kindrad; }
    hf_sync<L>();
}

template <int L> __device__ __forceinline__ int hf_regen(const HfParams &P, HfWarp &w, int t, bool reinject, int lane, int *queue);

// The xsel branch of the specification's first_reaction_method: exciton slot k fires its pending event.
// *xnew = k when the slot needs a new event afterwards (drawn by hf_step's tail).
template <int L>
__device__ __forceinline__ int hfi_fire(const HfParams &P, HfWarp &w, int k, int kind, int fin, double tau, int lane, int *queue,
                                        int *xnew) {
    const int site = w.x_site[k], meta = w.x_meta[k];
    const int spin = HFI_SPIN(meta);
    hf_sync<L>();
    if (kind == HF_K_FORSTER || kind == HF_K_MOVE) {                             // transfer: reseau.py:549-550 / 528
        if (meta & HFI_ONSITE) { hf_clear_flag<L>(w, 0, site, lane); hf_clear_flag<L>(w, 1, site, lane); }
        if (spin == HF_XS_SINGLET) w.c_x_forster += 1; else w.c_x_dexter += 1;
        if (lane == 0) { w.x_site[k] = fin; w.x_meta[k] = meta & 3; }
        hf_sync<L>();
        hf_update_events<L, true>(w, tau, lane);
        *xnew = k;
        return HF_ST_OK;
    }
    if (kind == HF_K_ISC) {                                                      // 526-527, molecule.py:299-307
        if (spin == HF_XS_SINGLET) w.c_x_isc += 1; else w.c_x_risc += 1;
        if (lane == 0) w.x_meta[k] = (meta & HFI_ONSITE) | (spin == HF_XS_SINGLET ? HF_XS_TRIPLET : HF_XS_SINGLET);
        hf_sync<L>();
        hf_update_events<L, true>(w, tau, lane);
        *xnew = k;
        return HF_ST_OK;
    }
    // decay: 541-547
    const int sp = w.species[hf_site(P, site)];
    const bool photon = (meta & HFI_RAD) != 0;
    if (meta & HFI_ONSITE) { hf_clear_flag<L>(w, 0, site, lane); hf_clear_flag<L>(w, 1, site, lane); }   // exciton_decay
    {   // remove slot k keeping the order of the rest
        const int n = w.n_x;
#pragma unroll 1
        for (int base = k; base < n - 1; base += L) {
            const int i = base + lane;
            int a = 0, b = 0, c = 0;
            double d = 0.0;
            if (i < n - 1) { a = w.x_site[i + 1]; b = w.x_meta[i + 1]; c = w.x_final[i + 1]; d = w.x_tau[i + 1]; }
            hf_sync<L>();
            if (i < n - 1) { w.x_site[i] = a; w.x_meta[i] = b; w.x_final[i] = c; w.x_tau[i] = d; }
            hf_sync<L>();
        }
        w.n_x = n - 1;
    }
    w.recombination += 1;                                                        // 472-476
    w.c_recomb_host += sp == 0; w.c_recomb_tadf += sp == 1; w.c_recomb_fluo += sp == 2;
    if (photon) { w.emission += 1; w.c_emit_tadf += sp == 1; w.c_emit_fluo += sp == 2; }
    hf_update_events<L, true>(w, tau, lane);
    const int st = hf_regen<L>(P, w, 0, true, lane, queue);
    if (st != HF_ST_OK) return st;
    return hf_regen<L>(P, w, 1, true, lane, queue);
}
