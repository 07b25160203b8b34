"""Python driver of the host-side warp emulator (tests/emu/warp_emu.h).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libemu_frm.so")
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "hyperfluorescence_b200", "csrc")
DEPS = [os.path.join(HERE, "emu_frm.cpp"), os.path.join(HERE, "warp_emu.h")] + \
       [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]

EVENT_DTYPE = np.dtype([("initial", "<i4"), ("final", "<i4"), ("tau", "<f8"), ("kind", "u1"), ("particule", "u1"),
                        ("pad", "u1", (2,)), ("spins", "<u4"), ("draws", "<u8")])
assert EVENT_DTYPE.itemsize == 32


def build():
    # HF_EMU_LIB: a prebuilt variant, e.g. one compiled with -fsanitize=address (run python under LD_PRELOAD=libasan.so)
    if os.environ.get("HF_EMU_LIB"):
        return os.environ["HF_EMU_LIB"]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
                        "-shared", "-fPIC", "-o", LIB, os.path.join(HERE, "emu_frm.cpp")], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def unpack(packed, dims):
    packed = np.asarray(packed, dtype=np.int64)
    x, y, z = packed & 1023, (packed >> 10) & 1023, packed >> 20
    return ((z * dims[1] + y) * dims[0] + x).astype(np.int32)


def generate(dims, props, seed, realisation=0):
    N = dims[0] * dims[1] * dims[2]
    out = dict(species=np.zeros(N, np.uint8), homo=np.zeros(N), lumo=np.zeros(N), s1=np.zeros(N), t1=np.zeros(N))
    d = (C.c_int32 * 3)(*dims)
    p = (C.c_double * 3)(*props)
    rc = lib().emu_generate(d, p, C.c_uint64(seed), C.c_uint32(realisation),
                            *(_p(out[k]) for k in ("species", "homo", "lumo", "s1", "t1")))
    if rc:
        raise ValueError("emu_generate failed")
    return out


class EmuXModel(C.Structure):
    _fields_ = [("mode", C.c_int32), ("M", C.c_int32)] + [(k, C.c_void_p) for k in (
        "offsets", "g6", "gd", "kf", "kd", "k_isc", "k_risc", "k_rad", "k_nonrad", "ws1", "wt1")] + \
        [("w_stride", C.c_int64), ("w_halo", C.c_int64)]


def xmodel(mode, dims=None, params=None, tables=None, ws1=None, wt1=None):
    """The integrated exciton model as the device would hold it (P.xi): packed offsets, haloed tagged weights."""
    m = EmuXModel()
    m.mode = int(mode)
    keep = []
    if m.mode == 2:
        import math
        off = np.asarray(tables["offsets"], np.int32)
        packed = np.ascontiguousarray((off[:, 0] + 128) | ((off[:, 1] + 128) << 8) | ((off[:, 2] + 128) << 16), np.int32)
        halo = int(math.ceil(params["cutoff"])) * dims[0] * dims[1]
        N = dims[0] * dims[1] * dims[2]
        def haloed(w):
            out = np.zeros(N + 2 * halo)
            out[halo:halo + N] = w
            return out
        arrs = dict(offsets=packed, g6=tables["g6"], gd=tables["gd"], kf=tables["kf"], kd=tables["kd"], k_isc=params["k_isc"],
                    k_risc=params["k_risc"], k_rad=params["k_rad"], k_nonrad=params["k_nonrad"], ws1=haloed(ws1), wt1=haloed(wt1))
        for k, v in arrs.items():
            a = v if k == "offsets" else np.ascontiguousarray(v, np.float64)
            keep.append(a)
            setattr(m, k, a.ctypes.data)
        m.M, m.w_stride, m.w_halo = int(packed.size), N + 2 * halo, halo
    m._keep = keep
    return m


def run(dims, field, charges, seed, trajectory, species, homo, lumo, stop, n_calls=1, replay=None, small_kernel=False,
        distance=1, excitons=None):
    dims = tuple(int(v) for v in dims)
    cap, cap_f = charges, (2 if (charges == 1 and distance == 1) else 2 * charges + 8)
    trace_cap = stop * n_calls
    trace = np.zeros(max(trace_cap, 1), EVENT_DTYPE)
    ints = np.zeros(12, np.int32)
    u64 = np.zeros(8, np.uint64)
    lists = np.zeros(6 * cap, np.int32)
    flags = np.zeros(2 * cap_f, np.int32)
    taus = np.zeros(2 * cap, np.float64)
    totals = np.zeros(24, np.uint64)
    x_ints, x_tau, x_scal = np.zeros(3 * cap, np.int32), np.zeros(cap, np.float64), np.zeros(3, np.float64)
    cache_i = np.zeros(30, np.int32)
    cache_tau = np.zeros(10, np.float64)
    init_lists = np.zeros(6 * cap, np.int32)
    init_taus = np.zeros(2 * cap, np.float64)
    init_counts = np.zeros(5, np.int32)
    species = np.ascontiguousarray(species, np.uint8)
    homo = np.ascontiguousarray(homo, np.float64)
    lumo = np.ascontiguousarray(lumo, np.float64)
    rp = None if replay is None else np.ascontiguousarray(replay, np.float64)
    d = (C.c_int32 * 3)(*dims)
    rc = lib().emu_run(d, C.c_double(field), C.c_int32(charges), C.c_uint64(seed), C.c_uint32(trajectory),
                       _p(species), _p(homo), _p(lumo), None if rp is None else _p(rp),
                       C.c_int64(0 if rp is None else rp.size), C.c_int64(stop), C.c_int32(n_calls),
                       C.c_int32(int(small_kernel)), C.c_int32(int(distance)), _p(trace), C.c_int64(trace_cap), _p(ints), _p(u64), _p(lists), _p(flags), _p(taus),
                       _p(totals), _p(cache_i), _p(cache_tau), _p(init_lists), _p(init_taus), _p(init_counts),
                       None if excitons is None else C.byref(excitons), _p(x_ints), _p(x_tau), _p(x_scal))
    if rc:
        raise RuntimeError(f"emu_run failed ({rc})")
    n_e, n_h, n_fe, n_fh, n_me, n_mh = (int(v) for v in ints[:6])
    fired = int(u64[6])
    out = dict(
        trace=trace[:min(fired, trace_cap)],
        e_pos=unpack(lists[0:n_e], dims), h_pos=unpack(lists[cap:cap + n_h], dims),
        me_init=unpack(lists[2 * cap:2 * cap + n_me], dims), mh_init=unpack(lists[3 * cap:3 * cap + n_mh], dims),
        me_final=unpack(lists[4 * cap:4 * cap + n_me], dims), mh_final=unpack(lists[5 * cap:5 * cap + n_mh], dims),
        me_tau=taus[:n_me].copy(), mh_tau=taus[cap:cap + n_mh].copy(),
        e_flags=unpack(flags[:n_fe], dims), h_flags=unpack(flags[cap_f:cap_f + n_fh], dims),
        special=int(ints[6]), special_site=int(unpack([ints[7]], dims)[0]), xstate=int(ints[8]), status=int(ints[9]),
        draws=int(u64[0]), spins=int(u64[1]), emission=int(u64[2]), recombination=int(u64[3]),
        injection=int(u64[4]), steps=int(u64[5]), fired=fired, init_draws=int(u64[7]), totals=totals,
        cache_head=int(ints[10]), cache_n=int(ints[11]),
        cache_init=unpack(cache_i[0:10], dims), cache_final=unpack(cache_i[10:20], dims), cache_kp=cache_i[20:30].copy(),
        cache_tau=cache_tau,
    )
    if excitons is not None and excitons.mode:
        nx = int(x_scal[0])
        meta = x_ints[cap:cap + nx]
        out.update(x_site=unpack(x_ints[:nx], dims), x_spin=meta & 3, x_onsite=(meta >> 2) & 1, x_rad=(meta >> 3) & 1,
                   x_kind=(meta >> 4) & 15, x_final=unpack(x_ints[2 * cap:2 * cap + nx], dims), x_tau=x_tau[:nx].copy(),
                   xdraws=int(x_scal[1]), clock=float(x_scal[2]))
    ne0, nh0, nme0, nmh0, st0 = (int(v) for v in init_counts)
    out.update(
        init_status=st0,
        init_e_pos=unpack(init_lists[0:ne0], dims), init_h_pos=unpack(init_lists[cap:cap + nh0], dims),
        init_me_init=unpack(init_lists[2 * cap:2 * cap + nme0], dims), init_mh_init=unpack(init_lists[3 * cap:3 * cap + nmh0], dims),
        init_me_final=unpack(init_lists[4 * cap:4 * cap + nme0], dims), init_mh_final=unpack(init_lists[5 * cap:5 * cap + nmh0], dims),
        init_me_tau=init_taus[:nme0].copy(), init_mh_tau=init_taus[cap:cap + nmh0].copy(),
    )
    return out


SCALARS_DTYPE = np.dtype([("n_pos", "<i4", (2,)), ("n_flag", "<i4", (2,)), ("n_ev", "<i4", (2,)), ("special", "<i4"),
                          ("special_site", "<i4"), ("xstate", "<i4"), ("status", "<i4"), ("cache_head", "<i4"),
                          ("cache_n", "<i4"), ("draws", "<u8"), ("spins", "<u8"), ("emission", "<u8"),
                          ("recombination", "<u8"), ("injection", "<u8"), ("steps", "<u8"), ("fired", "<u8")])
assert SCALARS_DTYPE.itemsize == 104


def run_batch_c1(dims, field, seed, first_traj, B, use_c1, species, homo, lumo, stop, n_calls=1):
    """B <= 32 charges == 1 trajectories through hf_inject + hf_run with either step kernel."""
    dims = tuple(int(v) for v in dims)
    cap = stop * n_calls
    trace = np.zeros((B, max(cap, 1)), EVENT_DTYPE)
    sc = np.zeros(B, SCALARS_DTYPE)
    pos, ev_init, ev_final = (np.zeros((2, B), np.int32) for _ in range(3))
    ev_tau = np.zeros((2, B), np.float64)
    flag = np.zeros((2, B, 2), np.int32)
    cache_i = np.zeros((3, 10, B), np.int32)
    cache_tau = np.zeros((10, B), np.float64)
    totals = np.zeros(20, np.uint64)
    species = np.ascontiguousarray(species, np.uint8)
    homo = np.ascontiguousarray(homo, np.float64)
    lumo = np.ascontiguousarray(lumo, np.float64)
    d = (C.c_int32 * 3)(*dims)
    rc = lib().emu_run_batch_c1(d, C.c_double(field), C.c_uint64(seed), C.c_uint32(first_traj), C.c_int32(B),
                                C.c_int32(1 if use_c1 else 0), _p(species), _p(homo), _p(lumo), C.c_int64(stop),
                                C.c_int32(n_calls), _p(trace), C.c_int64(max(cap, 1)), _p(sc), _p(pos), _p(ev_init),
                                _p(ev_final), _p(ev_tau), _p(flag), _p(cache_i), _p(cache_tau), _p(totals))
    if rc:
        raise RuntimeError(f"emu_run_batch_c1 failed ({rc})")
    return dict(trace=trace, sc=sc, pos=pos, ev_init=ev_init, ev_final=ev_final, ev_tau=ev_tau, flag=flag,
                cache_i=cache_i, cache_tau=cache_tau, totals=totals)


def x_run(oracle, seed, first_traj, B, n_events, n_calls=1, generic=False, brick=False):
    """The exciton kernels under emulation, with the model / lattice held by an ExcitonOracle.
    brick: the variant that reads the weights from the 4 x 4 x 1 brick layout (X, Y multiples of 4, cutoff 3 / 4 / 5)."""
    o = oracle
    cap = n_events * n_calls
    trace = np.zeros((B, max(cap, 1)), EVENT_DTYPE)
    site, spin = np.zeros(B, np.int32), np.zeros(B, np.int32)
    time, life, draws = np.zeros(B), np.zeros(B), np.zeros(B, np.uint64)
    totals = np.zeros(16, np.uint64)
    d = (C.c_int32 * 3)(*o.dims)
    rc = lib().emu_x_run(d, _p(o.species), _p(o.ws1), _p(o.wt1), C.c_int32(len(o.g6)), _p(o.off), _p(o.g6), _p(o.gd),
                         _p(o.kf), _p(o.kd), _p(o.k_isc), _p(o.k_risc), _p(o.k_rad), _p(o.k_nonrad),
                         C.c_double(o.params["singlet_fraction"]), C.c_uint64(seed), C.c_uint32(first_traj), C.c_int32(B),
                         C.c_int64(n_events), C.c_int32(n_calls), C.c_int32(2 if brick else int(generic)), _p(trace), C.c_int64(max(cap, 1)), _p(site), _p(spin),
                         _p(time), _p(life), _p(draws), _p(totals))
    if rc:
        raise RuntimeError(f"emu_x_run failed ({rc})")
    return dict(trace=trace, site=site, spin=spin, time=time, lifetime_sum=life, draws=draws, totals=totals)


def brick_selftest(bX, bY, r_max):
    """Mismatches between the packed brick offsets of hfx_brick_lin and the brick addresses (0 = none)."""
    return int(lib().emu_brick_selftest(C.c_int32(bX), C.c_int32(bY), C.c_int32(r_max)))


def xd_run(oracle, seed, device, n_excitons, n_events, n_calls=1):
    """The high-density kernel (one device = one warp) under emulation, model held by a DenseExcitonOracle."""
    o = oracle
    cap = max(n_events * n_calls, 1)
    trace = np.zeros(cap, EVENT_DTYPE)
    n = int(n_excitons)
    site, spin, tot, scal = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n), np.zeros(3)
    totals = np.zeros(24, np.uint64)
    occ = np.zeros(int(np.prod(o.dims)), np.uint8)
    d = (C.c_int32 * 3)(*o.dims)
    rc = lib().emu_xd_run(d, _p(o.ws1), _p(o.wt1), C.c_int32(len(o.g6)), _p(o.off), _p(o.g6), _p(o.gd), _p(o.kf), _p(o.kd),
                          _p(o.k_isc), _p(o.k_risc), _p(o.k_rad), _p(o.k_nonrad), C.c_double(o.params["singlet_fraction"]),
                          C.c_double(o.params["cutoff"]), _p(o.ann), C.c_double(o.p_tta), C.c_uint64(seed), C.c_uint32(device),
                          C.c_int32(n), C.c_int64(n_events), C.c_int32(n_calls), _p(trace), C.c_int64(cap), _p(site), _p(spin),
                          _p(tot), _p(scal), _p(totals), _p(occ))
    if rc:
        raise RuntimeError(f"emu_xd_run failed ({rc})")
    return dict(trace=trace, site=site, spin=spin, total_rate=tot, time=scal[0], lifetime_sum=scal[1], draws=int(scal[2]),
                totals=totals, occupancy=occ)
